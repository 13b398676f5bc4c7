"""Deterministic synthetic weights / features / labels for parity tests, goldens and bench.

Every tensor is drawn from its own CPU generator seeded by (seed, crc32(name)), so a tensor's
values depend only on its name, shape and the seed — not on construction order, torch RNG
consumption of other code, or the device.  The same state_dict can therefore be loaded into
the reference class (golden generation, build container), the CPU oracle and the CUDA path
(GPU box) without shipping the weights.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Optional, Sequence, Tuple

import torch

from .specs import TranslatorSpec


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


def make_state_dict(spec: TranslatorSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic fp32 parameters keyed by the reference state_dict names.

    Scales follow the torch defaults the reference relies on: matrices
    ~U(+-1/sqrt(fan_in)), embeddings ~N(0,1).  Biases and LayerNorm gamma/beta are made
    non-trivial (not 0/1) so that parity tests exercise them."""
    out: Dict[str, torch.Tensor] = {}
    for name, shape in spec.param_shapes().items():
        g = _gen(seed, name)
        if name in ("task_embed", "pe", "embedding.weight"):
            t = torch.randn(shape, generator=g)
        elif name.endswith(("norm1.weight", "norm2.weight", "norm3.weight", ".0.norm.weight", ".1.net.0.weight")) or name in ("ln.weight", "linear_head.0.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(("norm1.bias", "norm2.bias", "norm3.bias", ".0.norm.bias", ".1.net.0.bias")) or name in ("ln.bias", "linear_head.0.bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            bound = 1.0 / math.sqrt(shape[-1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # Linear / in_proj biases
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        out[name] = t.contiguous()
    return out


def make_features(spec: TranslatorSpec, batch: int, seg_tokens: Optional[Sequence[int]] = None, seed: int = 0,
                  dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Per-task features (B, D_k, K_k) ~ N(0,1): the shapes the frozen backbones emit."""
    if seg_tokens is None:
        seg_tokens = [s.tokens for s in spec.segments]
    feats = {}
    for s, d in zip(spec.segments, seg_tokens):
        g = _gen(seed, "feat." + s.name)
        feats[s.name] = torch.randn((batch, d, s.in_dim), generator=g).to(dtype)
    return feats


def make_labels(spec: TranslatorSpec, batch: int, seg_tokens: Optional[Sequence[int]] = None, seed: int = 0):
    """Task labels: TTM (B,) in {0,1}; ASD (B*D,) in {0,1}; PNR keyframe (B,) in [0,16) or OSCC (B,) in {0,1};
    LTA (B,Z,2) verb/noun ids; EgoT2-g (rows,3) task-prompt target tokens."""
    g = _gen(seed, "labels")
    if spec.family == "hhi_ttm":
        return torch.randint(0, 2, (batch,), generator=g)
    if spec.family == "hhi_asd":
        d = seg_tokens[0]
        return torch.randint(0, 2, (batch * d,), generator=g)
    if spec.family == "hoi_pnr":
        return torch.randint(0, spec.n_out, (batch,), generator=g)
    if spec.family == "hoi_ar":        # (B, 2): verb id, noun id (RecognitionTask2Loader labels[:, 0] / labels[:, 1])
        v = torch.randint(0, spec.head_groups[0], (batch, 1), generator=g)
        n = torch.randint(0, spec.head_groups[1], (batch, 1), generator=g)
        return torch.cat([v, n], dim=-1)
    if spec.family == "hoi_lta":
        v = torch.randint(0, spec.head_groups[0], (batch, spec.n_heads_out, 1), generator=g)
        n = torch.randint(0, spec.head_groups[1], (batch, spec.n_heads_out, 1), generator=g)
        return torch.cat([v, n], dim=-1)
    if spec.family == "hhi_g":
        # task-prompt targets (rows, 3) = [task word, answer, answer] with the answers in {'0': 5, '1': 6}; the decoder
        # reads target[:, :-1] and is scored on target[:, 1:] (HHI/tasks/multitask/video_tasktranslation.py:48-61).
        # 'asd' scores every frame: rows = B * D (vocab: HHI/utils/utils.py:12-18)
        rows = batch * seg_tokens[0] if spec.g_mode == "asd" else batch
        task_tok = {"ttm": 2, "lam": 3, "asd": 4}[spec.g_mode]
        ans = torch.randint(5, 7, (rows, 2), generator=g)
        return torch.cat([torch.full((rows, 1), task_tok, dtype=torch.int64), ans], dim=1)
    if spec.family == "hoi_g":
        # (B, 3) = [task word, answer, answer] drawn from the whole vocabulary past the task words, e.g. ['action', verb,
        # noun] (HOI/tasks/multitask/video_task.py:182-199: decoder input target[:, :-1], CE on target[:, 1:])
        ans = torch.randint(5, spec.vocab, (batch, 2), generator=g)
        return torch.cat([torch.full((batch, 1), 4, dtype=torch.int64), ans], dim=1)
    raise ValueError(spec.family)

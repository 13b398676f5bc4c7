"""TEST / BENCH INFRASTRUCTURE ONLY - training steps of the UNMODIFIED reference translator classes on bench workloads.

Used by `bench.py --impl reference` (CPU, all host threads: the reference arm) and by `bench.py`'s `gpu_baseline` leg (the
same classes in stock torch eager on the same B200, fp32 and bf16 autocast: the "existing GPU path" comparator of SURVEY.md
section 8d).  The classes come from `oracle/ref_shims.py`: the reference source tree in the build container, its byte-compiled
modules under `oracle/_ref/` on the GPU box.  Never imported by the product package.

One step = what the reference's Lightning task does per batch: forward in train mode (dropout on), the task's own loss,
`loss.backward()`, `torch.optim.Adam.step()` (HHI/tasks/ttm/video_task.py:36,64-66; HOI/tasks/pnr/video_taskspecific_pnr.py:29-31;
HOI/tasks/lta/long_term_anticipation_taskspecfic.py:177-183; HHI/tasks/multitask/video_tasktranslation.py:39-66).
"""
from __future__ import annotations

import contextlib
import io
import warnings
from typing import Callable, List, Tuple

import torch

from egot2_b200 import specs, synth
from . import ref_shims as rs
from .cases import Case
from .make_golden import build_reference, hoi_g_reference_forward, reference_forward_loss

_LOADED = {}


def _ref():
    if "m" not in _LOADED:
        warnings.filterwarnings("ignore")
        _LOADED["m"] = (rs.load_hhi(), rs.load_hoi())
    return _LOADED["m"]


def available() -> bool:
    return rs.reference_available()


def _to(x, dev, dt=None):
    if isinstance(x, dict):
        return {k: _to(v, dev, dt) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_to(v, dev, dt) for v in x]
    if torch.is_floating_point(x) and dt is not None:
        return x.to(device=dev, dtype=dt)
    return x.to(dev)


def make_step(wl_name: str, wl: dict, device: str = "cpu", autocast_bf16: bool = False, adam: bool = True,
              training: bool = True, n_batches: int = 2, batch: int = 0, seed: int = 0) -> Tuple[Callable[[int], float], dict]:
    """Returns (step(i) -> loss tensor, info).  training=False: forward only under no_grad in eval mode."""
    hhi, hoi = _ref()
    spec = wl["spec"]()
    B = batch or wl["batch"]
    dev = torch.device(device)
    if wl.get("prompt") and wl.get("g_kind") != "hoi":
        modes = ("lam", "ttm", "asd")
        sub = {m_: (specs.hhi_g_spec(spec.hidden, spec.heads, spec.layers, spec.p_layer, m_), wl["g_batches"][m_]) for m_ in modes}
    case = Case(wl_name, spec, B, tuple(wl["seg_tokens"]), seed)
    with contextlib.redirect_stdout(io.StringIO()):      # the reference constructors print ("Freezing task-specific models")
        m = build_reference(case, hhi, hoi)
    m.load_state_dict(synth.make_state_dict(spec, seed), strict=False)
    m.to(dev)
    m.train(training)
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4) if (adam and training) else None

    batches: List = []
    for i in range(n_batches):
        if wl.get("prompt") and wl.get("g_kind") == "hoi":
            parts = []
            for j, task in enumerate(("pnr", "oscc", "action")):
                b = wl["g_batches"][task]
                parts.append((_to(synth.make_features(spec, b, seed=seed * 7 + 10 * i + j), dev),
                              synth.make_labels(spec, b, seed=seed * 7 + 10 * i + j).to(dev)))
            batches.append(parts)
        elif wl.get("prompt"):
            parts = []
            for m_ in ("lam", "ttm", "asd"):
                sp_m, (b, d) = sub[m_]
                seg = (d,) if m_ == "lam" else (d, d, d)
                parts.append((Case(wl_name + m_, sp_m, b, seg, seed), _to(synth.make_features(sp_m, b, seg, seed=seed * 7 + 10 * i), dev),
                              synth.make_labels(sp_m, b, seg, seed=seed * 7 + 10 * i).to(dev)))
            batches.append(parts)
        else:
            batches.append((_to(synth.make_features(spec, B, wl["seg_tokens"], seed=seed + 100 * i), dev),
                            synth.make_labels(spec, B, wl["seg_tokens"], seed=seed + 100 * i).to(dev)))

    def loss_of(bt):
        if wl.get("prompt") and wl.get("g_kind") == "hoi":
            loss = 0.0
            for feats, labels in bt:
                out = hoi_g_reference_forward(spec, m, feats, labels[:, :-1])
                loss = loss + torch.nn.functional.cross_entropy(out.float(), labels[:, 1:])
            return loss
        if wl.get("prompt"):
            loss = 0.0
            for c, feats, labels in bt:
                _, l = reference_forward_loss(c, m, hhi, feats, labels, {})
                loss = loss + l
            return loss
        feats, labels = bt
        _, l = reference_forward_loss(case, m, hhi, feats, labels, {}, keep_head_dropout=True)
        return l

    def step(i: int):
        bt = batches[i % len(batches)]
        if not training:
            with torch.no_grad(), torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast_bf16):
                return loss_of(bt).detach()
        if opt is not None:
            opt.zero_grad(set_to_none=True)
        else:
            for p in params:
                p.grad = None
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast_bf16):
            loss = loss_of(bt)
        loss.backward()
        if opt is not None:
            opt.step()
        return loss.detach()

    info = {"class": type(m).__name__, "source": rs.reference_kind(), "params": sum(p.numel() for p in params),
            "step": ("forward (eval, no_grad)" if not training else "forward (train mode, dropout on) + loss + backward" + (" + torch.optim.Adam" if opt else "")),
            "dtype": "bf16 autocast" if autocast_bf16 else "fp32", "device": str(dev), "clips": B}
    return step, info

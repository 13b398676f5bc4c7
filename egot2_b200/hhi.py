"""Drop-in HHI EgoT2-s translators (same class names, ctor args, forward signatures, state_dict keys).

Reference (relative to /root/reference):
  ttm.TaskFusionMFTransformer2Task   HHI/models/ttm/model_taskspecific.py:154-194
  ttm.TaskFusionMFTransformer3Task   HHI/models/ttm/model_taskspecific.py:197-245
  asd.TaskFusionMFTransformer3Task   HHI/models/asd/model_taskspecific.py:109-158
  lossAV                             HHI/tasks/asd/loss.py:11-30
Two different classes share the name TaskFusionMFTransformer3Task in the reference (one per task
registry); they live in the `ttm` and `asd` namespaces below, each with its MODEL_REGISTRY/build_model
like HHI/models/{ttm,asd}/build.py.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .engine import TranslatorEngine
from .modules import PrecomputedFeatures, TranslatorBase, device_softmax
from .functional import translator_apply
from .specs import hhi_asd_spec, hhi_g_spec, hhi_ttm_spec


class PositionalEncoding(nn.Module):
    """Parameter-free container of the sinusoid buffer `pe` (max_len,1,d_model) — same buffer name and
    values as HHI/models/ttm/model_taskspecific.py:131-151; the add + Dropout(0.1) run inside libegot2."""

    def __init__(self, d_model, dropout=0.1, max_len=1000):
        super().__init__()
        self.p = dropout
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(1))


def _reference_backbones(args, want_asd: bool) -> Dict[str, nn.Module]:
    """Build the frozen task-specific backbones with the reference's own classes (only possible when this
    package is used inside an EgoT2 checkout: `models.*` / `utils.*` importable)."""
    try:
        from models.lam.model import LAMBackbone            # type: ignore
        from models.ttm.model import TTMBackbone            # type: ignore
        from utils.utils import freeze_params, load_ckpt    # type: ignore
    except Exception as e:  # pragma: no cover - depends on the host checkout
        raise _lib.Egot2Error("the frozen LAM/TTM/ASD backbones are not part of egot2_b200: run inside an EgoT2 "
                              "checkout or pass backbones={'lam_model':..., 'ttm_model':..., 'asd_model':...}") from e
    out = {"lam_model": LAMBackbone(args.lam_checkpoint), "ttm_model": TTMBackbone(args.ttm_checkpoint)}
    freeze_params(out["lam_model"])
    if want_asd:
        from models.asd.talkNetModel import talkNetModel   # type: ignore
        out["asd_model"] = talkNetModel()
        load_ckpt(out["asd_model"], args.asd_checkpoint, load_asd=True)
        freeze_params(out["asd_model"])
    if not getattr(args, "nofreeze", False):
        freeze_params(out["ttm_model"])
    return out


class _HHITranslator(TranslatorBase):
    def _build(self, args, n_tasks: int, spec, backbones):
        self.n_tasks = n_tasks
        self.dim = args.hidden_dim
        self.n_heads = args.num_heads
        self.dp_rate = args.dropout
        self.num_layers = args.num_layers
        if backbones is None:
            backbones = _reference_backbones(args, n_tasks == 3)
        for k, m in backbones.items():
            setattr(self, k, m)
        # parameter containers in the reference's registration order (same RNG draws, same state_dict keys)
        self.proj_lam = nn.Linear(256, self.dim)
        self.proj_ttm = nn.Linear(256, self.dim)
        if n_tasks == 3:
            self.proj_asd = nn.Linear(256, self.dim)
        self.task_embed = nn.Parameter(torch.randn(1, n_tasks, self.dim), requires_grad=True)
        self.pos_embed = PositionalEncoding(self.dim, dropout=0.1)
        self.transformer_encoder = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.dim, nhead=self.n_heads, dropout=self.dp_rate),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.ln = nn.LayerNorm(self.dim)
        self.linear_head = nn.Sequential(nn.LayerNorm(self.dim), nn.Linear(self.dim, 2))
        self._poison_containers(self.proj_lam, self.proj_ttm, self.transformer_encoder, self.ln, self.linear_head,
                                *([self.proj_asd] if n_tasks == 3 else []))
        self._init_translator(spec)

    def _configure_engine(self, eng: TranslatorEngine):
        eng.set_sinusoid(self.pos_embed.pe)

    def _asd_features(self, video_asd, audio_asd):
        """The reference's TalkNet call sequence (model_taskspecific.py:229-234) -> (N, D, 256)."""
        N, D = video_asd.shape[0], video_asd.shape[1]
        a = self.asd_model.forward_audio_frontend(audio_asd)
        v = self.asd_model.forward_visual_frontend(video_asd)
        a, v = self.asd_model.forward_cross_attention(a, v)
        outs = self.asd_model.forward_audio_visual_backend(a, v)
        return outs.view(N, D, -1)


class _TTM2Task(_HHITranslator):
    """Task Translation for 2 tasks: LAM and TTM -> TTM logits (B,2)."""

    def __init__(self, args, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self._build(args, 2, hhi_ttm_spec(args.hidden_dim, args.num_heads, args.num_layers, args.dropout, False),
                    backbones)

    def forward(self, video, audio):
        lam_out = self.lam_model(video, middle=True)           # (bs, D, 256)
        ttm_out = self.ttm_model(video, audio, middle=True)
        return self._translate([ttm_out, lam_out])             # token order (ttm, lam)


class _TTM3Task(_HHITranslator):
    """Task Translation for 3 HHI tasks: LAM, TTM, ASD -> TTM logits (B,2)."""

    def __init__(self, args, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self._build(args, 3, hhi_ttm_spec(args.hidden_dim, args.num_heads, args.num_layers, args.dropout, True),
                    backbones)

    def forward(self, video, video_asd, audio, audio_asd):
        asd_out = self._asd_features(video_asd, audio_asd)
        lam_out = self.lam_model(video, middle=True)
        ttm_out = self.ttm_model(video, audio, middle=True)
        return self._translate([ttm_out, lam_out, asd_out])    # token order (ttm, lam, asd)


class _ASD3Task(_HHITranslator):
    """ASD-of-interest translator: returns the encoded ASD tokens (N*D, hidden); head lives in lossAV."""

    def __init__(self, args, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self._build(args, 3, hhi_asd_spec(args.hidden_dim, args.num_heads, args.num_layers, args.dropout), backbones)
        self.output_dim = self.dim

    def forward(self, video, video_asd, audio, audio_asd):
        with torch.no_grad():
            asd_out = self._asd_features(video_asd, audio_asd)
            lam_out = self.lam_model(video, middle=True)
            ttm_out = self.ttm_model(video, audio, middle=True)
        return self._translate([asd_out, ttm_out, lam_out])    # token order (asd, ttm, lam)


class TaskTranslationPromptTransformer(_HHITranslator):
    """EgoT2-g (HHI/models/multitask/task_prompt_model.py:174-293): one model, three forwards per step
    ('lam': LAM tokens only; 'ttm': lam+ttm+asd tokens; 'asd': same encoder, 3-token memory per frame), an
    nn.TransformerDecoder over the task prompt and a 7-word vocabulary head.  Imported directly by
    HHI/tasks/multitask/video_tasktranslation.py:18,35 (not through a registry)."""

    def __init__(self, args, vocab, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.args = args
        self.vocab = vocab
        self.n_tasks = 3
        self.dim = args.hidden_dim
        self.n_heads = args.num_heads
        self.num_layers = args.num_layers
        self.dp_rate = args.dropout
        self.max_output_length = 500
        n_vocab = len(vocab) if vocab is not None else 7
        # parameter containers in the reference's registration order
        self.transformer_encoder = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.dim, nhead=self.n_heads, dropout=self.dp_rate),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.transformer_decoder = nn.TransformerDecoder(
            decoder_layer=nn.TransformerDecoderLayer(d_model=self.dim, nhead=self.n_heads, dropout=self.dp_rate),
            num_layers=self.num_layers)
        self.ln = nn.LayerNorm(self.dim)
        self.task_embed = nn.Parameter(torch.randn(1, self.n_tasks, self.dim), requires_grad=True)
        self.pos_embed = PositionalEncoding(self.dim, dropout=0.1)
        self.embedding = nn.Embedding(n_vocab, self.dim)
        self.proj_lam = nn.Linear(256, self.dim)
        self.proj_ttm = nn.Linear(256, self.dim)
        self.proj_asd = nn.Linear(256, self.dim)
        self.fc = nn.Linear(self.dim, n_vocab)
        self.seq_len = 2
        for p in self.parameters():                      # _init_parameters(): xavier on every dim>1 parameter
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if backbones is None:
            backbones = _reference_backbones(args, True)
        for k, m in backbones.items():
            setattr(self, k, m)
        self._poison_containers(self.transformer_encoder, self.transformer_decoder, self.ln, self.proj_lam, self.proj_ttm,
                                self.proj_asd, self.fc)
        self.embedding.forward = None
        self._specs = {m: hhi_g_spec(self.dim, self.n_heads, self.num_layers, self.dp_rate, m, vocab=n_vocab)
                       for m in ("lam", "ttm", "asd")}
        self._init_translator(self._specs["ttm"])
        self._mode_engines: Dict[str, TranslatorEngine] = {}

    def _engine_for(self, mode: str, device: torch.device) -> TranslatorEngine:
        base = self._ensure_engine(device)               # 'ttm' spec: owns the arena every parameter lives in
        if mode == "ttm":
            return base
        eng = self._mode_engines.get(mode)
        if eng is None or eng.arena is not base.arena or eng.dtype != base.dtype:
            eng = TranslatorEngine(self._specs[mode], device, base.dtype, arena=base.arena)
            eng.set_sinusoid(self.pos_embed.pe)
            self._mode_engines[mode] = eng
        return eng

    def _features(self, video, video_asd, audio, audio_asd, task):
        with torch.no_grad():
            lam_feat = self.lam_model(video, middle=True)
            if task == "lam":
                return [lam_feat]
            ttm_feat = self.ttm_model(video, audio, middle=True)
            asd_feat = self._asd_features(video_asd, audio_asd)
        return [lam_feat, ttm_feat, asd_feat]                 # token order (lam id0, ttm id1, asd id2)

    def _decode(self, feats, tokens, task):
        eng = self._engine_for(task, feats[0].device)
        out = translator_apply(eng, list(feats), self._params(), self._param_names, self.training, self._next_seed(),
                               prompt=tokens)
        rows, S = tokens.shape
        return out.view(rows, S, -1)                          # (rows, seq, vocab)

    def forward(self, video, video_asd, audio, audio_asd, target, task):
        assert task in ["lam", "ttm", "asd"]
        feats = self._features(video, video_asd, audio, audio_asd, task)
        return self._decode(feats, target, task).permute(0, 2, 1)      # (bs, vocab_size, seq_y)

    def predict(self, video, video_asd, audio, audio_asd, task):
        assert task in ["lam", "ttm", "asd"]
        feats = self._features(video, video_asd, audio, audio_asd, task)
        rows = feats[0].shape[0] * (feats[0].shape[1] if task == "asd" else 1)
        y = torch.full((rows, 1), int(self.vocab[task]), dtype=torch.int64, device=feats[0].device)
        return self._decode(feats, y, task)[:, 0, -2:]       # logits of the words '0', '1'


def _registry(*classes):
    reg = {}
    for name, cls in classes:
        cls.__name__ = name
        cls.__qualname__ = name
        reg[name] = cls
    return reg


ttm = SimpleNamespace(TaskFusionMFTransformer2Task=_TTM2Task, TaskFusionMFTransformer3Task=_TTM3Task)
ttm.MODEL_REGISTRY = _registry(("TaskFusionMFTransformer2Task", _TTM2Task), ("TaskFusionMFTransformer3Task", _TTM3Task))
ttm.build_model = lambda args, **kw: ttm.MODEL_REGISTRY[args.model](args, **kw)

asd = SimpleNamespace(TaskFusionMFTransformer3Task=_ASD3Task)
asd.MODEL_REGISTRY = {"TaskFusionMFTransformer3Task": _ASD3Task}
asd.build_model = lambda args, **kw: asd.MODEL_REGISTRY[args.model](args, **kw)


class lossAV(nn.Module):
    """HHI/tasks/asd/loss.py:11-30 — FC(dim->2) + CrossEntropy(weight [1,4]) + softmax score / rounded label /
    correct count, with FC, CE and their gradients computed by libegot2 (egot2_head_loss_fwd/bwd)."""

    def __init__(self, dim=256):
        super().__init__()
        self.criterion = nn.CrossEntropyLoss(weight=torch.FloatTensor([1, 4]))   # container: keeps `criterion.weight`
        self.FC = nn.Linear(dim, 2)
        self.FC.forward = None
        self.dim = dim

    def forward(self, x, labels=None):
        from .losses import linear_ce
        x = x.squeeze(1)
        if labels is None:
            logits, _ = linear_ce(x, self.FC.weight, self.FC.bias, None, None)
            return logits[:, 1].t().reshape(-1).detach().cpu().numpy()
        logits, nloss = linear_ce(x, self.FC.weight, self.FC.bias, labels, self.criterion.weight)
        predScore = device_softmax(logits)
        predLabel = torch.round(predScore)[:, 1]
        correctNum = (predLabel == labels).sum().float()
        return nloss, predScore, predLabel, correctNum

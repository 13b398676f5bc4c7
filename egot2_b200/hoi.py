"""Drop-in HOI EgoT2-s translators (same class names, ctor `cfg`, forward signatures, state_dict keys).

Reference (relative to /root/reference):
  pnr.TaskFusionMFTransformer3TaskDropout   HOI/models/pnr/video_model_transfer_3task.py:212-258
  lta.TaskFusionMFTransformerLTA4Task       HOI/models/lta/lta_models_lta_transfer.py:257-377
  pnr.TaskFusionMFTransformerDropout        HOI/models/pnr/video_model_transfer.py:70-105   (2-task sibling)
  lta.TaskFusionMFTransformer3Task          HOI/models/lta/lta_models_transfer.py:96-137    (action-recognition sibling)
  lta.TaskFusionMFTransformer2TaskAR        HOI/models/lta/lta_models_transfer.py:169-235   (AR from recognition + LTA features)
  lta.TaskFusionMFTransformer2Task          HOI/models/lta/lta_models_lta_transfer.py:429-526 (LTA 2-task sibling, H <= 1024)
  MultiTaskHead (LTA head)                  HOI/models/lta/head_helper.py:218-291
The frozen PNR/OSCC/SlowFast/LTA backbones are not part of this package: inside an EgoT2 checkout
they are built by the reference's own loaders; otherwise pass `backbones={...}`.
"""
from __future__ import annotations

import ctypes as C
from functools import reduce
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical

from . import _lib as L
from .engine import _stream
from .modules import PrecomputedFeatures, TranslatorBase
from .specs import hoi_ar2_spec, hoi_ar_spec, hoi_lta2_spec, hoi_lta_spec, hoi_pnr2_spec, hoi_pnr_spec


def slowfast_pool(x5: torch.Tensor, t_out: int, out_dtype: torch.dtype) -> torch.Tensor:
    """AdaptiveAvgPool3d((t_out,1,1)) + squeeze + permute(0,2,1): (B,C,Tin,h,w) -> (B,t_out,C), one kernel
    (video_model_transfer_3task.py:226-227,245-247)."""
    if not x5.is_cuda:
        raise L.Egot2Error("egot2_b200 runs on CUDA only (no CPU fallback)")
    B, Cc, Tin, h, w = x5.shape
    x5 = x5.contiguous()
    if x5.dtype not in (torch.float32, torch.bfloat16):
        x5 = x5.float()
    if x5.dtype == torch.bfloat16:
        out_dtype = torch.bfloat16
    out = torch.empty((B, t_out, Cc), device=x5.device, dtype=out_dtype)
    code = {torch.float32: L.F32, torch.bfloat16: L.BF16}
    L.call("egot2_slowfast_pool_fwd", x5.data_ptr(), code[x5.dtype], B, Cc, Tin, h * w, t_out, out.data_ptr(),
           code[out_dtype], _stream())
    return out


class _PNR3TaskDropout(TranslatorBase):
    """mid fusion transformer: PNR + OSCC + AR(slow,fast) -> PNR keyframe logits (B,1,16) or OSCC (B,2,1)."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg_pnr = None
        self.cfg_recognition = None
        if backbones is None:
            backbones = _reference_pnr_backbones(self, cfg)
        for k, m in backbones.items():
            setattr(self, k, m)
        self.num_classes = 16 if "keyframe_localization" in cfg.DATA.TASK else 2
        self.unsqueeze_dim = 1 if "keyframe_localization" in cfg.DATA.TASK else 2
        self.sequence_len = 48
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.proj1 = nn.Linear(8192, self.feature_dim)
        self.proj2 = nn.Linear(8192, self.feature_dim)
        self.proj3_slow = nn.Linear(2048, self.feature_dim)
        self.proj3_fast = nn.Linear(256, self.feature_dim)
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.ln = nn.LayerNorm(self.feature_dim)
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=8,
                                                     dropout=cfg.MODEL.TRANSFORMER_DROPOUT_RATE,
                                                     dim_feedforward=self.feature_dim * 2, batch_first=True),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.linear_head = nn.Sequential(self.ln, nn.Linear(self.feature_dim, self.num_classes))   # shared ln (F5)
        self._poison_containers(self.proj1, self.proj2, self.proj3_slow, self.proj3_fast, self.transformer,
                                self.linear_head)
        self._init_translator(hoi_pnr_spec(self.feature_dim, self.num_layers, self.num_classes,
                                           cfg.MODEL.FEAT_DROPOUT_RATE, cfg.MODEL.TRANSFORMER_DROPOUT_RATE))

    def forward(self, x1, x2):
        x_pnr = x1
        x_oscc = x1.copy()
        x_action = x2
        pnr_feat = self.pnr_model(x_pnr, middle=True)                    # (bs, 16, 8192)
        oscc_feat = self.oscc_model(x_oscc, middle=True)                 # (bs, 16, 8192)
        slow5, fast5 = self.recognition_model(x_action, middle=True)     # (bs,2048,8,7,7), (bs,256,32,7,7)
        dt = torch.float32 if self.compute_dtype == "fp32" else torch.bfloat16
        slow = slowfast_pool(slow5, slow5.shape[2], dt) if slow5.dim() == 5 else slow5
        fast = slowfast_pool(fast5, 8, dt) if fast5.dim() == 5 else fast5
        out = self._translate([pnr_feat, oscc_feat, slow, fast])         # token order (pnr, oscc, slow, fast)
        return out.unsqueeze(self.unsqueeze_dim)


class _PNR2TaskDropout(TranslatorBase):
    """2-task sibling (HOI/models/pnr/video_model_transfer.py:70-105): PNR + OSCC -> keyframe logits (B,1,16) / OSCC
    (B,2,1); H=256, 3 layers, 32 tokens, head = bare Linear on the mean token."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg_recognition = None
        if backbones is None:
            backbones = _reference_pnr_backbones(self, cfg, with_recognition=False)
        for k, m in backbones.items():
            setattr(self, k, m)
        self.num_classes = 16 if cfg.DATA.TASK == "keyframe_localization" else 2
        self.unsqueeze_dim = 1 if cfg.DATA.TASK == "keyframe_localization" else 2
        self.sequence_len = 32
        self.feature_dim = 256
        self.proj1 = nn.Linear(8192, self.feature_dim)
        self.proj2 = nn.Linear(8192, self.feature_dim)
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.ln = nn.LayerNorm(self.feature_dim)
        self.dpmode = cfg.MODEL.FEAT_DROPOUT_MODE
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=8,
                                                     dropout=cfg.MODEL.TRANSFORMER_DROPOUT_RATE,
                                                     dim_feedforward=self.feature_dim * 2, batch_first=True),
            num_layers=3, enable_nested_tensor=False)
        self.linear_head = nn.Linear(self.feature_dim, self.num_classes)
        self._poison_containers(self.proj1, self.proj2, self.transformer, self.linear_head, self.ln)
        self._init_translator(hoi_pnr2_spec(self.num_classes, cfg.MODEL.TRANSFORMER_DROPOUT_RATE))

    def forward(self, x):
        if self.dpmode > 0 and self.training:
            # reference :95-98 then drops the PNR segment's projected features only; the embed stage has one
            # feature-dropout switch for all segments, so this (non-default) mode is refused rather than approximated
            raise L.Egot2Error("TaskFusionMFTransformerDropout: FEAT_DROPOUT_MODE > 0 (per-segment feature dropout) is "
                               "not built; the shipped default is 0")
        x2 = x.copy()
        pnr_feat = self.pnr_model(x, middle=True)                        # (bs, 16, 8192)
        oscc_feat = self.oscc_model(x2, middle=True)                     # (bs, 16, 8192)
        out = self._translate([pnr_feat, oscc_feat])                     # token order (pnr, oscc)
        return out.unsqueeze(self.unsqueeze_dim)


class _AR3Task(TranslatorBase):
    """Action-recognition sibling (HOI/models/lta/lta_models_transfer.py:96-137): AR(slow,fast) + PNR + OSCC ->
    [verb logits (B,115), noun logits (B,478)]; one LayerNorm shared by the token LN and both heads."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg_pnr = None
        self.cfg_recognition = None
        if backbones is None:
            backbones = _reference_pnr_backbones(self, cfg)
        for k, m in backbones.items():
            setattr(self, k, m)
        num_cls1, num_cls2 = cfg.MODEL.NUM_CLASSES
        self.num_classes = (num_cls1, num_cls2)
        self.sequence_len = 48
        self.num_heads = cfg.MODEL.TRANSLATION_HEADS
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        self.dp_rate = cfg.MODEL.TRANSLATION_DROPOUT
        self.proj1 = nn.Linear(8192, self.feature_dim)
        self.proj2 = nn.Linear(8192, self.feature_dim)
        self.proj3_slow = nn.Linear(2048, self.feature_dim)
        self.proj3_fast = nn.Linear(256, self.feature_dim)
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=self.num_heads,
                                                     dropout=self.dp_rate, batch_first=True),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.ln = nn.LayerNorm(self.feature_dim)
        self.linear_head1 = nn.Sequential(self.ln, nn.Linear(self.feature_dim, num_cls1))     # shared ln
        self.linear_head2 = nn.Sequential(self.ln, nn.Linear(self.feature_dim, num_cls2))
        self._poison_containers(self.proj1, self.proj2, self.proj3_slow, self.proj3_fast, self.transformer,
                                self.linear_head1, self.linear_head2)
        self._init_translator(hoi_ar_spec(self.feature_dim, self.num_layers, self.num_heads, self.dp_rate,
                                          (num_cls1, num_cls2)))

    def forward(self, x_action, x_pnr):
        x_oscc = x_pnr.copy()
        pnr_feat = self.pnr_model(x_pnr, middle=True)                    # (bs, 16, 8192)
        oscc_feat = self.oscc_model(x_oscc, middle=True)                 # (bs, 16, 8192)
        slow5, fast5 = self.recognition_model(x_action, middle=True)     # (bs,2048,8,7,7), (bs,256,32,7,7)
        dt = torch.float32 if self.compute_dtype == "fp32" else torch.bfloat16
        slow = slowfast_pool(slow5, slow5.shape[2], dt) if slow5.dim() == 5 else slow5
        fast = slowfast_pool(fast5, 8, dt) if fast5.dim() == 5 else fast5
        out = self._translate([slow, fast, pnr_feat, oscc_feat])         # token order (slow, fast, pnr, oscc)
        return list(torch.split(out, list(self.num_classes), dim=-1))


class _AR2Task(TranslatorBase):
    """`TaskFusionMFTransformer2TaskAR` (HOI/models/lta/lta_models_transfer.py:169-235): the last input clip through the
    recognition backbone (slow/fast maps) + the first NUM_INPUT_CLIPS clips through the LTA backbone -> [verbs, nouns]."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg = cfg
        self.num_input = cfg.FORECASTING.NUM_INPUT_CLIPS
        self.input_offset = cfg.FORECASTING.INPUT_OFFSET
        num_cls1, num_cls2 = cfg.MODEL.NUM_CLASSES
        self.num_classes = (num_cls1, num_cls2)
        self.sequence_len = 18
        self.num_heads = cfg.MODEL.TRANSLATION_HEADS
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        self.dp_rate = cfg.MODEL.TRANSLATION_DROPOUT
        self.proj_lta = nn.Linear(2048, self.feature_dim)
        self.proj_slow = nn.Linear(2048, self.feature_dim)
        self.proj_fast = nn.Linear(256, self.feature_dim)
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=self.num_heads,
                                                     dropout=self.dp_rate, batch_first=True),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.ln = nn.LayerNorm(self.feature_dim)
        self.linear_head1 = nn.Sequential(self.ln, nn.Linear(self.feature_dim, num_cls1))
        self.linear_head2 = nn.Sequential(self.ln, nn.Linear(self.feature_dim, num_cls2))
        for p in self.parameters():                       # reference _init_parameters (:211-214), before the backbones exist
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._poison_containers(self.proj_lta, self.proj_slow, self.proj_fast, self.transformer, self.linear_head1,
                                self.linear_head2)
        if backbones is None:
            backbones = _reference_ar2_backbones(cfg)
        for k, m in backbones.items():
            setattr(self, k, m)
        self._init_translator(hoi_ar2_spec(self.feature_dim, self.num_layers, self.num_heads, self.dp_rate,
                                           (num_cls1, num_cls2)))

    def forward(self, x):
        x1 = x.copy()
        x_action = [x1[0][:, -1, ...], x1[1][:, -1, ...]]
        x_lta = [x[0][:, 0:self.num_input, ...], x[1][:, 0:self.num_input, ...]]
        with torch.no_grad():
            slow5, fast5 = self.action_model(x_action, middle=True)      # (bs,2048,8,7,7), (bs,256,32,7,7)
            feat_lta = self.lta_model(x_lta, middle=True).transpose(0, 1)   # (bs, num_input, 2048)
        return self.translate(slow5, fast5, feat_lta)

    def translate(self, slow5, fast5, feat_lta):
        dt = torch.float32 if self.compute_dtype == "fp32" else torch.bfloat16
        slow = slowfast_pool(slow5, slow5.shape[2], dt) if slow5.dim() == 5 else slow5
        fast = slowfast_pool(fast5, 8, dt) if fast5.dim() == 5 else fast5
        out = self._translate([slow, fast, feat_lta.contiguous()])       # token order (slow, fast, lta)
        return list(torch.split(out, list(self.num_classes), dim=-1))


def _reference_ar2_backbones(cfg):  # pragma: no cover - needs an EgoT2 checkout + checkpoints
    """lta_models_transfer.py:216-227: recognition SlowFast trunk (no head) + LTA encoder (no decoder), both frozen."""
    try:
        from models.lta.video_model_builder import SlowFast                                        # type: ignore
        from models.lta.lta_models import ForecastingEncoderDecoder                                # type: ignore
        from utils.lta.parser import load_config_from_file as load_lta_config                      # type: ignore
        from utils.multitask.load_model import load_lta_backbone, freeze_backbone_params, freeze_params  # type: ignore
    except Exception as e:
        raise L.Egot2Error("the frozen SlowFast/LTA backbones are not part of egot2_b200: run inside an EgoT2 checkout "
                           "or pass backbones={'action_model':..., 'lta_model':...}") from e
    out = {}
    cfg_rec = load_lta_config(cfg.PRETRAIN.ACTION_CFG)
    cfg_rec.MODEL.NUM_CLASSES = [cfg.MODEL.TRANSLATION_INPUT_FEATURES]
    cfg_rec.MODEL.HEAD_ACT = None
    out["action_model"] = SlowFast(cfg_rec, with_head=False)
    load_lta_backbone(out["action_model"], cfg_rec.CHECKPOINT_FILE_PATH, True, True)
    freeze_backbone_params(out["action_model"])
    cfg_lta = load_lta_config(cfg.PRETRAIN.LTA_CFG)
    out["lta_model"] = ForecastingEncoderDecoder(cfg_lta, build_decoder=False)
    load_lta_backbone(out["lta_model"], cfg_lta.CHECKPOINT_FILE_PATH)
    freeze_params(out["lta_model"])
    return out


def _reference_pnr_backbones(self, cfg, with_recognition=True):  # pragma: no cover - needs an EgoT2 checkout + checkpoints
    try:
        from models.pnr.video_model_builder import KeyframeLocalizationResNet, StateChangeClsResNet  # type: ignore
        from models.lta.video_model_builder import SlowFast                                        # type: ignore
        from utils.pnr.parser import load_config_file                                              # type: ignore
        from utils.lta.parser import load_config_from_file as load_lta_config                      # type: ignore
        from utils.multitask.load_model import (load_checkpoint, freeze_params,                   # type: ignore
                                                load_recognition_backbone, freeze_backbone_params)
    except Exception as e:
        raise L.Egot2Error("the frozen PNR/OSCC/SlowFast backbones are not part of egot2_b200: run inside an EgoT2 "
                           "checkout or pass backbones={'pnr_model':..., 'oscc_model':..., 'recognition_model':...}") from e
    out = {}
    cfg_pnr = load_config_file(cfg.PRETRAIN.PNR_CFG)
    out["pnr_model"] = KeyframeLocalizationResNet(cfg_pnr)
    load_checkpoint(out["pnr_model"], cfg_pnr.MISC.CHECKPOINT_FILE_PATH)
    if cfg.PRETRAIN.PNR_FT:
        out["pnr_model"].eval(); freeze_params(out["pnr_model"])
    cfg_oscc = load_config_file(cfg.PRETRAIN.OSCC_CFG)
    self.cfg_pnr = cfg_oscc
    cfg_oscc.MODEL.NO_TEMP_POOL = True
    out["oscc_model"] = StateChangeClsResNet(cfg_oscc)
    load_checkpoint(out["oscc_model"], cfg_oscc.MISC.CHECKPOINT_FILE_PATH)
    if cfg.PRETRAIN.OSCC_FT:
        out["oscc_model"].eval(); freeze_params(out["oscc_model"])
    if not with_recognition:
        return out
    cfg_rec = load_lta_config(cfg.PRETRAIN.ACTION_CFG)
    cfg_rec.MODEL.NUM_CLASSES = [cfg.MODEL.TRANSLATION_INPUT_FEATURES]
    cfg_rec.MODEL.HEAD_ACT = None
    self.cfg_recognition = cfg_rec
    out["recognition_model"] = SlowFast(cfg_rec, with_head=False)
    load_recognition_backbone(out["recognition_model"], cfg_rec.CHECKPOINT_FILE_PATH)
    if cfg.PRETRAIN.ACTION_FT:
        out["recognition_model"].eval(); freeze_backbone_params(out["recognition_model"])
    return out


class _HeadContainer(nn.Module):
    """Parameter container with MultiTaskHead's state_dict keys (`projections.{z}.weight/bias`)."""

    def __init__(self, dim_in: int, num_classes: List[int]):
        super().__init__()
        self.projections = nn.ModuleList([nn.Linear(dim_in, n, bias=True) for n in num_classes])


class _LTA4Task(TranslatorBase):
    """PNR + OSCC + AR + LTA -> 20 future (verb, noun) distributions: [(B,Z,115), (B,Z,478)]."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg = cfg
        self.sequence_len = cfg.FORECASTING.NUM_INPUT_CLIPS * 4
        self.num_heads = cfg.MODEL.TRANSLATION_HEADS
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        self.dp_rate = cfg.MODEL.TRANSLATION_DROPOUT
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.proj_pnr = nn.Linear(8192, self.feature_dim)
        self.proj_oscc = nn.Linear(8192, self.feature_dim)
        self.proj_lta = nn.Linear(2048, self.feature_dim)
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=self.num_heads,
                                                     dropout=self.dp_rate, batch_first=True),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.ln = nn.LayerNorm(self.feature_dim)
        for p in self.parameters():                      # _init_parameters(): xavier on every dim>1 translator param
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if backbones is None:
            backbones = _reference_lta_backbones(self, cfg)
        for k, m in backbones.items():
            setattr(self, k, m)
        self.num_classes = list(cfg.MODEL.NUM_CLASSES)
        per_head = reduce(lambda a, b: a + b, self.num_classes)
        Z = cfg.FORECASTING.NUM_ACTIONS_TO_PREDICT
        self.head = _HeadContainer(self.feature_dim, [per_head] * Z)     # default nn.Linear init, like the reference
        self.test_noact = bool(cfg.TEST.NO_ACT)
        self._poison_containers(self.proj_pnr, self.proj_oscc, self.proj_lta, self.transformer, self.ln, self.head)
        self._init_translator(hoi_lta_spec(self.feature_dim, self.num_layers, self.num_heads, self.dp_rate,
                                           cfg.FORECASTING.NUM_INPUT_CLIPS, Z, tuple(self.num_classes),
                                           cfg.MODEL.DROPOUT_RATE))

    # --- feature extraction with the frozen backbones (reference loops, lines 321-346) ---
    def encode_clips(self, model, x):
        assert isinstance(x, list) and len(x) >= 1
        feats = [model([pathway[:, i] for pathway in x]) for i in range(x[0].shape[1])]
        return torch.stack(feats, dim=1)                       # (bs, num_inputs, d)

    def encode_clips_pnr(self, model, x):
        feats = [model([x[:, i, ...]], middle=True).mean(dim=1) for i in range(x.shape[1])]
        return torch.stack(feats, dim=1)                       # (bs, num_inputs, 8192)

    def translate(self, pnr, oscc, action, lta):
        """The hot path: per-input-clip features -> [(B,Z,#verbs), (B,Z,#nouns)]."""
        out = self._translate([pnr, oscc, action, lta])        # (B, Z*593) logits
        B = out.shape[0]
        out = out.view(B, len(self.head.projections), -1)
        if not self.training and not self.test_noact:
            out = torch.softmax(out, dim=-1)                   # MultiTaskHead eval activation (head_helper.py:284-286)
        return list(torch.split(out, self.num_classes, dim=-1))

    def forward(self, x_lta, x_pnr):
        pnr = self.encode_clips_pnr(self.pnr_model, x_pnr)
        oscc = self.encode_clips_pnr(self.oscc_model, x_pnr)
        action = self.encode_clips(self.action_model, x_lta)
        lta = self.lta_model(x_lta, None, middle=True).transpose(0, 1)   # (bs, num_input, 2048)
        return self.translate(pnr, oscc, action, lta)

    def generate(self, x_lta, x_pnr, k=1):
        x = self.forward(x_lta, x_pnr)
        results = []
        for head_x in x:
            if k > 1:
                dist = Categorical(logits=head_x)
                preds = [dist.sample() for _ in range(k)]
            elif k == 1:
                preds = [head_x.argmax(2)]
            results.append(torch.stack(preds, dim=1))
        return results


class _LTA2Task(TranslatorBase):
    """LTA 2-task sibling (HOI/models/lta/lta_models_lta_transfer.py:429-526): AR + LTA features of the input clips ->
    20 future (verb, noun) distributions.  TRANSLATION_INPUT_FEATURES == 2048 (proj_lta = Identity) is not built."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg = cfg
        self.sequence_len = cfg.FORECASTING.NUM_INPUT_CLIPS * 2
        self.num_heads = cfg.MODEL.TRANSLATION_HEADS
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        if self.feature_dim == 2048:
            raise L.Egot2Error("TaskFusionMFTransformer2Task with TRANSLATION_INPUT_FEATURES = 2048 (proj_lta = Identity) "
                               "is not built: the LayerNorm kernels stop at H = 1024")
        self.proj_lta = nn.Linear(2048, self.feature_dim)
        self.dp_rate = cfg.MODEL.TRANSLATION_DROPOUT
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.transformer = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.feature_dim, nhead=self.num_heads,
                                                     dropout=self.dp_rate, batch_first=True),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.ln = nn.LayerNorm(self.feature_dim)
        for p in self.parameters():                      # _init_parameters(): xavier on every dim>1 translator param
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if backbones is None:
            backbones = _reference_lta2_backbones(cfg)
        for k, m in backbones.items():
            setattr(self, k, m)
        self.num_classes = list(cfg.MODEL.NUM_CLASSES)
        per_head = reduce(lambda a, b: a + b, self.num_classes)
        Z = cfg.FORECASTING.NUM_ACTIONS_TO_PREDICT
        self.head = _HeadContainer(self.feature_dim, [per_head] * Z)     # default nn.Linear init, like the reference
        self.test_noact = bool(cfg.TEST.NO_ACT)
        self._poison_containers(self.proj_lta, self.transformer, self.ln, self.head)
        self._init_translator(hoi_lta2_spec(self.feature_dim, self.num_layers, self.num_heads, self.dp_rate,
                                            cfg.FORECASTING.NUM_INPUT_CLIPS, Z, tuple(self.num_classes),
                                            cfg.MODEL.DROPOUT_RATE))

    encode_clips = _LTA4Task.encode_clips

    def translate(self, action, lta):
        out = self._translate([action, lta])                   # (B, Z*593) logits
        B = out.shape[0]
        out = out.view(B, len(self.head.projections), -1)
        if not self.training and not self.test_noact:
            out = torch.softmax(out, dim=-1)                   # MultiTaskHead eval activation (head_helper.py:284-286)
        return list(torch.split(out, self.num_classes, dim=-1))

    def forward(self, x, tgts=None):
        action = self.encode_clips(self.action_model, x)                 # (bs, num_input, d)
        lta = self.lta_model(x, None, middle=True).transpose(0, 1)       # (bs, num_input, 2048)
        return self.translate(action, lta)

    def generate(self, x, k=1):
        results = []
        for head_x in self.forward(x):
            if k > 1:
                dist = Categorical(logits=head_x)
                preds = [dist.sample() for _ in range(k)]
            elif k == 1:
                preds = [head_x.argmax(2)]
            results.append(torch.stack(preds, dim=1))
        return results


def _reference_lta2_backbones(cfg):  # pragma: no cover - needs an EgoT2 checkout + checkpoints
    import copy
    try:
        from models.lta.video_model_builder import SlowFast                                        # type: ignore
        from models.lta.lta_models import ForecastingEncoderDecoder                                # type: ignore
        from utils.multitask.load_model import load_lta_backbone, freeze_backbone_params, freeze_params  # type: ignore
    except Exception as e:
        raise L.Egot2Error("the frozen SlowFast/LTA backbones are not part of egot2_b200: run inside an EgoT2 checkout "
                           "or pass backbones={'action_model':..., 'lta_model':...}") from e
    out = {}
    bcfg = copy.deepcopy(cfg)
    bcfg.MODEL.NUM_CLASSES = [cfg.MODEL.TRANSLATION_INPUT_FEATURES]
    bcfg.MODEL.HEAD_ACT = None
    out["action_model"] = SlowFast(bcfg, with_head=True)
    load_lta_backbone(out["action_model"], cfg.CHECKPOINT_FILE_PATH_AR, True, True)
    freeze_backbone_params(out["action_model"])
    out["lta_model"] = ForecastingEncoderDecoder(cfg, build_decoder=False)
    load_lta_backbone(out["lta_model"], cfg.CHECKPOINT_FILE_PATH_LTA)
    freeze_params(out["lta_model"])
    return out


def _reference_lta_backbones(self, cfg):  # pragma: no cover - needs an EgoT2 checkout + checkpoints
    import copy
    try:
        from models.pnr.video_model_builder import KeyframeLocalizationResNet, StateChangeClsResNet  # type: ignore
        from models.lta.video_model_builder import SlowFast                                        # type: ignore
        from models.lta.lta_models import ForecastingEncoderDecoder                                # type: ignore
        from utils.pnr.parser import load_config_file as load_pnr_config                           # type: ignore
        from utils.multitask.load_model import (load_ckpt, load_lta_backbone, freeze_backbone_params,  # type: ignore
                                                freeze_params)
    except Exception as e:
        raise L.Egot2Error("the frozen PNR/OSCC/SlowFast/LTA backbones are not part of egot2_b200: run inside an EgoT2 "
                           "checkout or pass backbones={'pnr_model','oscc_model','action_model','lta_model'}") from e
    out = {}
    cfg_pnr = load_pnr_config(cfg.PRETRAIN.PNR_CFG)
    self.cfg_pnr = cfg_pnr
    out["pnr_model"] = KeyframeLocalizationResNet(cfg_pnr)
    load_ckpt(out["pnr_model"], cfg_pnr.MISC.CHECKPOINT_FILE_PATH); freeze_params(out["pnr_model"])
    cfg_oscc = load_pnr_config(cfg.PRETRAIN.OSCC_CFG)
    cfg_oscc.MODEL.NO_TEMP_POOL = False
    out["oscc_model"] = StateChangeClsResNet(cfg_oscc)
    load_ckpt(out["oscc_model"], cfg_oscc.MISC.CHECKPOINT_FILE_PATH); freeze_params(out["oscc_model"])
    bcfg = copy.deepcopy(cfg)
    bcfg.MODEL.NUM_CLASSES = [cfg.MODEL.TRANSLATION_INPUT_FEATURES]
    bcfg.MODEL.HEAD_ACT = None
    out["action_model"] = SlowFast(bcfg, with_head=True)
    load_lta_backbone(out["action_model"], cfg.CHECKPOINT_FILE_PATH_AR, True, True)
    freeze_backbone_params(out["action_model"])
    out["lta_model"] = ForecastingEncoderDecoder(cfg, build_decoder=True)
    load_lta_backbone(out["lta_model"], cfg.CHECKPOINT_FILE_PATH_LTA); freeze_params(out["lta_model"])
    return out


pnr = SimpleNamespace(TaskFusionMFTransformer3TaskDropout=_PNR3TaskDropout, TaskFusionMFTransformerDropout=_PNR2TaskDropout)
_PNR3TaskDropout.__name__ = _PNR3TaskDropout.__qualname__ = "TaskFusionMFTransformer3TaskDropout"
_PNR2TaskDropout.__name__ = _PNR2TaskDropout.__qualname__ = "TaskFusionMFTransformerDropout"
pnr.MODEL_REGISTRY = {"TaskFusionMFTransformer3TaskDropout": _PNR3TaskDropout,
                      "TaskFusionMFTransformerDropout": _PNR2TaskDropout}
pnr.build_model = lambda cfg, **kw: pnr.MODEL_REGISTRY[cfg.MODEL.MODEL_NAME](cfg, **kw)

lta = SimpleNamespace(TaskFusionMFTransformerLTA4Task=_LTA4Task, TaskFusionMFTransformer3Task=_AR3Task,
                      TaskFusionMFTransformer2TaskAR=_AR2Task, TaskFusionMFTransformer2Task=_LTA2Task)
_LTA2Task.__name__ = _LTA2Task.__qualname__ = "TaskFusionMFTransformer2Task"
_AR2Task.__name__ = _AR2Task.__qualname__ = "TaskFusionMFTransformer2TaskAR"
_LTA4Task.__name__ = _LTA4Task.__qualname__ = "TaskFusionMFTransformerLTA4Task"
_AR3Task.__name__ = _AR3Task.__qualname__ = "TaskFusionMFTransformer3Task"
lta.MODEL_REGISTRY = {"TaskFusionMFTransformerLTA4Task": _LTA4Task, "TaskFusionMFTransformer3Task": _AR3Task,
                      "TaskFusionMFTransformer2TaskAR": _AR2Task, "TaskFusionMFTransformer2Task": _LTA2Task}
lta.build_model = lambda cfg, **kw: lta.MODEL_REGISTRY[cfg.MODEL.MODEL_NAME](cfg, **kw)

"""Drop-in HOI EgoT2-s translators (same class names, ctor `cfg`, forward signatures, state_dict keys).

Reference (relative to /root/reference):
  pnr.TaskFusionMFTransformer3TaskDropout   HOI/models/pnr/video_model_transfer_3task.py:212-258
  lta.TaskFusionMFTransformerLTA4Task       HOI/models/lta/lta_models_lta_transfer.py:257-377
  pnr.TaskFusionMFTransformerDropout        HOI/models/pnr/video_model_transfer.py:70-105   (2-task sibling)
  pnr.TaskFusionMFTransformer3Task          HOI/models/pnr/video_model_transfer_3task.py:128-164 (simple_vit sibling)
  pnr.TaskFusionMFTransformer               HOI/models/pnr/video_model_transfer.py:44-67       (2-task simple_vit sibling)
  lta.TaskFusionMFTransformer3Task          HOI/models/lta/lta_models_transfer.py:96-137    (action-recognition sibling)
  lta.TaskFusionMFTransformer2TaskAR        HOI/models/lta/lta_models_transfer.py:169-235   (AR from recognition + LTA features)
  lta.TaskFusionMFTransformer2Task          HOI/models/lta/lta_models_lta_transfer.py:429-526 (LTA 2-task sibling)
  MultiTaskHead (LTA head)                  HOI/models/lta/head_helper.py:218-291
  multitask.TaskTranslationPromptTransformer       HOI/models/multitask/video_model_builder.py:223-275 (HOI EgoT2-g)
  multitask.TaskTranslationPromptTransformer6Task  HOI/models/multitask/video_model_builder.py:279-383
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
from .engine import TranslatorEngine, _stream
from .functional import translator_apply
from .modules import PrecomputedFeatures, TranslatorBase, device_mean_dim1, device_softmax
from .specs import (hoi_ar2_spec, hoi_ar_spec, hoi_g_spec, hoi_lta2_spec, hoi_lta_spec, hoi_pnr2_spec, hoi_pnr2_vit_spec,
                    hoi_pnr_spec, hoi_pnr_vit_spec)


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


class _VitAttention(nn.Module):
    """Parameter container with simple_vit Attention's keys (`norm`, bias-free `to_qkv` / `to_out`; simple_vit.py:67-78)."""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, heads * dim_head * 3, bias=False)
        self.to_out = nn.Linear(heads * dim_head, dim, bias=False)


class _VitFeedForward(nn.Module):
    """Parameter container with simple_vit FeedForward's keys (`net.0` LayerNorm, `net.1` / `net.3` Linear; :55-63)."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, dim))


class _VitTransformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([_VitAttention(dim, heads, dim_head), _VitFeedForward(dim, mlp_dim)])
                                     for _ in range(depth)])


class _PNR3TaskVit(TranslatorBase):
    """simple_vit sibling (HOI/models/pnr/video_model_transfer_3task.py:128-164): the 48 tokens of the Dropout variant
    through a pre-norm / GELU simple_vit Transformer(dim 256, depth 3, heads 8, dim_head 128, mlp_dim 512); the head shares
    `ln` with the token LayerNorm; no dropout."""

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
        self.feature_dim = 256
        self.proj1 = nn.Linear(8192, self.feature_dim)
        self.proj2 = nn.Linear(8192, self.feature_dim)
        self.proj3_slow = nn.Linear(2048, self.feature_dim)
        self.proj3_fast = nn.Linear(256, self.feature_dim)
        self.pe = nn.Parameter(torch.randn(1, self.sequence_len, self.feature_dim), requires_grad=True)
        self.transformer = _VitTransformer(dim=self.feature_dim, depth=3, heads=8, dim_head=128, mlp_dim=512)
        self.ln = nn.LayerNorm(self.feature_dim)
        self.linear_head = nn.Sequential(self.ln, nn.Linear(self.feature_dim, self.num_classes))   # shared ln
        self._poison_containers(self.proj1, self.proj2, self.proj3_slow, self.proj3_fast, self.transformer,
                                self.linear_head)
        self._init_translator(hoi_pnr_vit_spec(self.num_classes))

    forward = _PNR3TaskDropout.forward


class _PNR2TaskVit(TranslatorBase):
    """2-task simple_vit sibling (HOI/models/pnr/video_model_transfer.py:44-67): PNR + OSCC projections + pe (no token
    LayerNorm) -> simple_vit Transformer -> mean -> Sequential(LayerNorm, Linear)."""

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
        self.transformer = _VitTransformer(dim=self.feature_dim, depth=3, heads=8, dim_head=128, mlp_dim=512)
        self.linear_head = nn.Sequential(nn.LayerNorm(self.feature_dim), nn.Linear(self.feature_dim, self.num_classes))
        self._poison_containers(self.proj1, self.proj2, self.transformer, self.linear_head)
        self._init_translator(hoi_pnr2_vit_spec(self.num_classes))

    def forward(self, x):
        x2 = x.copy()
        pnr_feat = self.pnr_model(x, middle=True)                        # (bs, 16, 8192)
        oscc_feat = self.oscc_model(x2, middle=True)                     # (bs, 16, 8192)
        out = self._translate([pnr_feat, oscc_feat])                     # token order (pnr, oscc)
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
        # FEAT_DROPOUT_MODE > 0 (:95-96): Dropout(FEAT_DROPOUT_RATE) on the projected PNR features only
        self._init_translator(hoi_pnr2_spec(self.num_classes, cfg.MODEL.TRANSFORMER_DROPOUT_RATE,
                                            cfg.MODEL.FEAT_DROPOUT_RATE, self.dpmode))

    def forward(self, x):
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
        feats = [device_mean_dim1(model([x[:, i, ...]], middle=True)) for i in range(x.shape[1])]
        return torch.stack(feats, dim=1)                       # (bs, num_inputs, 8192)

    def translate(self, pnr, oscc, action, lta):
        """The hot path: per-input-clip features -> [(B,Z,#verbs), (B,Z,#nouns)]."""
        out = self._translate([pnr, oscc, action, lta])        # (B, Z*593) logits
        B = out.shape[0]
        out = out.view(B, len(self.head.projections), -1)
        if not self.training and not self.test_noact:
            out = device_softmax(out)                          # MultiTaskHead eval activation (head_helper.py:284-286)
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
    20 future (verb, noun) distributions; at TRANSLATION_INPUT_FEATURES == 2048 (the shipped config) proj_lta = Identity."""

    def __init__(self, cfg, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        self.cfg = cfg
        self.sequence_len = cfg.FORECASTING.NUM_INPUT_CLIPS * 2
        self.num_heads = cfg.MODEL.TRANSLATION_HEADS
        self.num_layers = cfg.MODEL.TRANSLATION_LAYERS
        self.feature_dim = cfg.MODEL.TRANSLATION_INPUT_FEATURES
        # :441-444 - at the shipped width 2048 the LTA features enter as they are
        self.proj_lta = nn.Identity() if self.feature_dim == 2048 else nn.Linear(2048, self.feature_dim)
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
            out = device_softmax(out)                          # MultiTaskHead eval activation (head_helper.py:284-286)
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


# ------------------------------------------------------------------------------------------------ HOI EgoT2-g
class _PromptTranslator(TranslatorBase):
    """HOI EgoT2-g `TaskTranslationPromptTransformer` (HOI/models/multitask/video_model_builder.py:223-275 on the
    TaskPromptTransformer base, :54-160): PNR + OSCC + action (slow | fast) tokens -> nn.TransformerEncoder -> memory;
    an nn.TransformerDecoder over the task prompt ([task word, answer, ...]) and a vocabulary head.  Imported directly by
    HOI/tasks/multitask/video_task.py:20,176 (not through a registry)."""

    _n_tasks = 3

    def __init__(self, args, vocab, oscc_no_temp_pool=True, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__()
        from .hhi import PositionalEncoding
        self.args = args
        self.vocab = vocab
        self.dim = args.hidden_dim
        self.n_tasks = 3
        self.task_dict = {"pnr": 0, "oscc": 1, "action": 2}
        self.n_heads = args.num_heads
        self.num_layers = args.num_layers
        self.dp_rate = args.dropout
        n_vocab = len(vocab)
        # parameter containers in the reference's registration order (:70-90): same RNG draws, same state_dict keys
        self.transformer_encoder = nn.TransformerEncoder(
            encoder_layer=nn.TransformerEncoderLayer(d_model=self.dim, nhead=self.n_heads, dropout=self.dp_rate),
            num_layers=self.num_layers, enable_nested_tensor=False)
        self.transformer_decoder = nn.TransformerDecoder(
            decoder_layer=nn.TransformerDecoderLayer(d_model=self.dim, nhead=self.n_heads, dropout=self.dp_rate),
            num_layers=self.num_layers)
        self.proj_pnr = nn.Linear(8192, self.dim)
        self.proj_oscc = nn.Linear(8192, self.dim)
        self.proj_action_slow = nn.Linear(2048, self.dim)
        self.proj_action_fast = nn.Linear(256, self.dim)
        self.fc = nn.Linear(self.dim, n_vocab)
        self.ln = nn.LayerNorm(self.dim)
        self.task_embed = nn.Parameter(torch.randn(1, self.n_tasks, self.dim), requires_grad=True)
        self.pos_embed = PositionalEncoding(self.dim, dropout=0.1, max_len=200)
        self.embedding = nn.Embedding(n_vocab, self.dim)
        self.seq_len = 5
        self.y_mask = self.get_tgt_mask(self.seq_len)
        for p in self.parameters():                      # _init_parameters (:121-124), before the backbones exist
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if backbones is None:
            backbones = _reference_multitask_backbones(self, args, oscc_no_temp_pool)
        for k, m in backbones.items():
            setattr(self, k, m)
        self._finish()

    def _finish(self):
        self._poison_containers(self.transformer_encoder, self.transformer_decoder, self.ln, self.proj_pnr, self.proj_oscc,
                                self.proj_action_slow, self.proj_action_fast, self.fc)
        self.embedding.forward = None
        self._specs = {"clip": hoi_g_spec(self.dim, self.n_heads, self.num_layers, self.dp_rate, len(self.vocab), "clip",
                                          self._n_tasks)}
        self._init_translator(self._specs["clip"])
        self._mode_engines: Dict[str, TranslatorEngine] = {}

    def get_tgt_mask(self, size) -> torch.Tensor:
        """(:127-142) kept for API parity; the decoder kernels apply the causal mask themselves."""
        mask = torch.tril(torch.ones(size, size) == 1).float()
        return mask.masked_fill(mask == 0, float("-inf")).masked_fill(mask == 1, 0.0)

    def _configure_engine(self, eng: TranslatorEngine):
        eng.set_sinusoid(self.pos_embed.pe)

    def _engine_for(self, mode: str, device: torch.device) -> TranslatorEngine:
        base = self._ensure_engine(device)               # the 'clip' spec owns the arena every parameter lives in
        if mode == "clip":
            return base
        eng = self._mode_engines.get(mode)
        if eng is None or eng.arena is not base.arena or eng.dtype != base.dtype:
            eng = TranslatorEngine(self._specs[mode], device, base.dtype, arena=base.arena)
            eng.set_sinusoid(self.pos_embed.pe)
            self._mode_engines[mode] = eng
        return eng

    def _clip_features(self, video_pnr, video_ac) -> List[torch.Tensor]:
        """The frozen backbones of encode() (:229-233) + the adaptive pooling of the SlowFast maps (:237-238)."""
        video_oscc = video_pnr.copy()
        with torch.no_grad():
            feat_pnr = self.pnr_model(video_pnr, middle=True)              # (bs, 16, 8192)
            feat_oscc = self.oscc_model(video_oscc, middle=True)           # (bs, 16, 8192)
            slow5, fast5 = self.recognition_model(video_ac, middle=True)   # (bs,2048,8,7,7), (bs,256,32,7,7)
        dt = torch.float32 if self.compute_dtype == "fp32" else torch.bfloat16
        slow = slowfast_pool(slow5, slow5.shape[2], dt) if slow5.dim() == 5 else slow5
        fast = slowfast_pool(fast5, 8, dt) if fast5.dim() == 5 else fast5
        return [feat_pnr, feat_oscc, slow, fast]                           # token order (pnr, oscc, slow | fast)

    def _decode(self, feats, tokens, mode="clip"):
        eng = self._engine_for(mode, feats[0].device)
        out = translator_apply(eng, list(feats), self._params(), self._param_names, self.training, self._next_seed(),
                               prompt=tokens)
        rows, S = tokens.shape
        return out.view(rows, S, -1)                                       # (bs, seq, vocab)

    def _start(self, feats, word) -> torch.Tensor:
        return torch.full((feats[0].shape[0], 1), int(self.vocab[word]), dtype=torch.int64, device=feats[0].device)

    def forward(self, video_pnr, video_ac, target):
        feats = self._clip_features(video_pnr, video_ac)
        return self._decode(feats, target).permute(0, 2, 1)                # (bs, vocab_size, seq_y)

    def predict(self, video_pnr, video_ac, task):
        assert task in ["pnr", "oscc", "action_verb", "action_noun"]
        feats = self._clip_features(video_pnr, video_ac)
        out = self._decode(feats, self._start(feats, task))[:, 0]          # (bs, vocab)
        return torch.argmax(out, dim=-1) if "action" in task else out

    def _greedy(self, feats, start_word: str, seq_len: int = 3, mode: str = "clip") -> torch.Tensor:
        """Greedy decoding of a fixed-length answer (:264-275): one encoder pass, then the decoder alone per new token."""
        B, dev = feats[0].shape[0], feats[0].device
        toks = torch.ones((B, seq_len), dtype=torch.int64, device=dev)
        toks[:, 0] = int(self.vocab[start_word])
        if self.training:                                                  # dropout on: the regular (autograd) path
            for sy in range(1, seq_len):
                toks[:, sy] = self._decode(feats, toks[:, :sy].contiguous(), mode)[:, -1].argmax(dim=-1)
            return toks[:, 1:]
        eng = self._engine_for(mode, dev)
        with torch.no_grad():
            act = None
            for sy in range(1, seq_len):
                y = toks[:, :sy].contiguous()
                act = eng.forward(list(feats), training=False, prompt=y) if act is None else eng.decode_again(act, y)
                toks[:, sy] = act.t["out"].view(B, sy, -1)[:, -1].argmax(dim=-1)
        return toks[:, 1:]                                                 # idx in vocab

    def predict_ac(self, video_pnr, video_ac):
        return self._greedy(self._clip_features(video_pnr, video_ac), "action")


class _PromptTranslator6Task(_PromptTranslator):
    """`TaskTranslationPromptTransformer6Task` (:279-383): the same clip encode for the pnr / oscc / action prompts with
    a 4-row task_embed, plus an 'lta' encode over (pnr, oscc, action, lta) features of the input clips (:325-339)."""

    _n_tasks = 4

    def __init__(self, args, vocab, backbones: Optional[Dict[str, nn.Module]] = None):
        super().__init__(args, vocab, backbones=backbones)

    def _finish(self):
        # registered after the base class is complete (:282-283): plain randn / default Linear init, no xavier
        self.task_embed = nn.Parameter(torch.randn(1, 4, self.dim), requires_grad=True)
        self.proj_lta = nn.Linear(2048, self.dim)
        if not hasattr(self, "lta_model"):
            self.lta_model = _reference_multitask_lta_backbone(self, self.args)
        super()._finish()
        self._poison_containers(self.proj_lta)                             # 'lta' specs: built on first use (num_input
                                                                           # comes with the data), see _features()

    encode_clips = None          # bound below (same loops as the LTA translators)
    encode_clips_pnr = None

    def _features(self, video_pnr, video_ac, task):
        if "lta" not in task:
            return self._clip_features(video_pnr, video_ac), "clip"
        import copy
        with torch.no_grad():
            video_oscc = copy.deepcopy(video_pnr)
            feat_pnr = self.encode_clips_pnr(self.pnr_model, video_pnr)    # (bs, num_input, 8192)
            feat_oscc = self.encode_clips_pnr(self.oscc_model, video_oscc)
            feat_action = self.encode_clips(self.recognition_model, video_ac)          # (bs, num_input, dim)
            feat_lta = self.lta_model(video_ac, None, middle=True).transpose(0, 1)     # (bs, num_input, 2048)
        n = feat_pnr.shape[1]
        key = f"lta{n}"
        if self._specs.get(key) is None:
            self._specs[key] = hoi_g_spec(self.dim, self.n_heads, self.num_layers, self.dp_rate, len(self.vocab), "lta", 4, n)
        return [feat_pnr.contiguous(), feat_oscc.contiguous(), feat_action.contiguous(), feat_lta.contiguous()], key

    def forward(self, video_pnr, video_ac, target, task):
        feats, mode = self._features(video_pnr, video_ac, task)
        return self._decode(feats, target, mode).permute(0, 2, 1)          # (bs, vocab_size, seq_y)

    def predict(self, video_pnr, video_ac, task, predict_verb_only=False, predict_noun_only=False):
        assert task in ["pnr", "oscc", "action", "lta"]
        feats, mode = self._features(video_pnr, video_ac, task)
        if task in ["action", "lta"]:
            if not predict_noun_only:
                out_verb = self._decode(feats, self._start(feats, task + "_verb"), mode)[:, 0]
            if predict_verb_only:
                return
            out_noun = self._decode(feats, self._start(feats, task + "_noun"), mode)[:, 0]
            if predict_noun_only:
                return
            return torch.stack((out_verb.argmax(dim=-1), out_noun.argmax(dim=-1)), dim=1)      # (bs, 2)
        return self._decode(feats, self._start(feats, task), mode)[:, 0]

    def predict_ac(self, video_pnr, video_ac):
        raise AttributeError("TaskTranslationPromptTransformer6Task has no predict_ac (the reference class predicts verb "
                             "and noun with two one-token prompts: predict(..., 'action'))")


_PromptTranslator6Task.encode_clips = _LTA4Task.encode_clips
_PromptTranslator6Task.encode_clips_pnr = _LTA4Task.encode_clips_pnr


def _reference_multitask_backbones(self, args, oscc_no_temp_pool=True):  # pragma: no cover - needs an EgoT2 checkout
    """video_model_builder.py:96-119: PNR / OSCC ResNets (frozen) and the recognition SlowFast with a hidden_dim head."""
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
    cfg_pnr = load_config_file(args.pnr_cfg_file)
    out["pnr_model"] = KeyframeLocalizationResNet(cfg_pnr)
    load_checkpoint(out["pnr_model"], cfg_pnr.MISC.CHECKPOINT_FILE_PATH)
    freeze_params(out["pnr_model"])
    cfg_oscc = load_config_file(args.oscc_cfg_file)
    cfg_oscc.MODEL.NO_TEMP_POOL = oscc_no_temp_pool
    out["oscc_model"] = StateChangeClsResNet(cfg_oscc)
    load_checkpoint(out["oscc_model"], cfg_oscc.MISC.CHECKPOINT_FILE_PATH)
    freeze_params(out["oscc_model"])
    cfg_rec = load_lta_config(args.action_cfg_file)
    cfg_rec.MODEL.NUM_CLASSES = [self.dim]
    cfg_rec.MODEL.HEAD_ACT = None
    out["recognition_model"] = SlowFast(cfg_rec, with_head=True)
    load_recognition_backbone(out["recognition_model"], cfg_rec.CHECKPOINT_FILE_PATH)
    freeze_backbone_params(out["recognition_model"])                       # the head stays trainable
    self.cfg_pnr, self.cfg_oscc, self.cfg_action = cfg_pnr, cfg_oscc, cfg_rec
    return out


def _reference_multitask_lta_backbone(self, args):  # pragma: no cover - needs an EgoT2 checkout
    """video_model_builder.py:285-290: the LTA encoder (no decoder), frozen."""
    import copy
    try:
        from models.lta.lta_models import ForecastingEncoderDecoder                                # type: ignore
        from utils.lta.parser import load_config_from_file as load_lta_config                      # type: ignore
        from utils.multitask.load_model import load_lta_backbone, freeze_params                    # type: ignore
    except Exception as e:
        raise L.Egot2Error("the frozen LTA backbone is not part of egot2_b200: run inside an EgoT2 checkout or pass "
                           "backbones={..., 'lta_model': ...}") from e
    cfg_lta = load_lta_config(args.lta_cfg_file)
    self.cfg_lta = copy.deepcopy(cfg_lta)
    cfg_lta.FORECASTING.NUM_ACTIONS_TO_PREDICT = 20
    m = ForecastingEncoderDecoder(cfg_lta, build_decoder=False)
    load_lta_backbone(m, cfg_lta.CHECKPOINT_FILE_PATH_LTA)
    freeze_params(m)
    return m


_PromptTranslator.__name__ = _PromptTranslator.__qualname__ = "TaskTranslationPromptTransformer"
_PromptTranslator6Task.__name__ = _PromptTranslator6Task.__qualname__ = "TaskTranslationPromptTransformer6Task"
multitask = SimpleNamespace(TaskTranslationPromptTransformer=_PromptTranslator,
                            TaskTranslationPromptTransformer6Task=_PromptTranslator6Task)

pnr = SimpleNamespace(TaskFusionMFTransformer3TaskDropout=_PNR3TaskDropout, TaskFusionMFTransformerDropout=_PNR2TaskDropout,
                      TaskFusionMFTransformer3Task=_PNR3TaskVit, TaskFusionMFTransformer=_PNR2TaskVit)
_PNR3TaskDropout.__name__ = _PNR3TaskDropout.__qualname__ = "TaskFusionMFTransformer3TaskDropout"
_PNR2TaskDropout.__name__ = _PNR2TaskDropout.__qualname__ = "TaskFusionMFTransformerDropout"
_PNR3TaskVit.__name__ = _PNR3TaskVit.__qualname__ = "TaskFusionMFTransformer3Task"
_PNR2TaskVit.__name__ = _PNR2TaskVit.__qualname__ = "TaskFusionMFTransformer"
pnr.MODEL_REGISTRY = {"TaskFusionMFTransformer3TaskDropout": _PNR3TaskDropout,
                      "TaskFusionMFTransformerDropout": _PNR2TaskDropout,
                      "TaskFusionMFTransformer3Task": _PNR3TaskVit,
                      "TaskFusionMFTransformer": _PNR2TaskVit}
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

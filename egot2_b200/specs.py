"""Static description of each EgoT2 translator variant on the hot path.

A `TranslatorSpec` says which per-task feature streams enter the translator, how each is
projected to `hidden`, what is added after the shared LayerNorm (HHI: learned task embedding
+ sinusoidal PE that restarts per task; HOI: one learned `pe`), the encoder geometry and the
head.  The spec also fixes the reference's state_dict key of every parameter, so the drop-in
modules, the synthetic-weight generator and the parity tests all agree on names/shapes.

Reference (relative to /root/reference):
  HHI/models/ttm/model_taskspecific.py:154-245   TaskFusionMFTransformer{2,3}Task (TTM)
  HHI/models/asd/model_taskspecific.py:109-158   TaskFusionMFTransformer3Task (ASD)
  HOI/models/pnr/video_model_transfer_3task.py:212-258   TaskFusionMFTransformer3TaskDropout
  HOI/models/lta/lta_models_lta_transfer.py:257-377      TaskFusionMFTransformerLTA4Task
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class Segment:
    """One task's feature stream = one contiguous run of tokens in every clip."""
    name: str                     # "ttm", "lam", "asd", "pnr", "oscc", "slow", "fast", "action", "lta"
    in_dim: int                   # feature width K of this task's frozen backbone
    proj: Optional[str]           # state_dict prefix of its nn.Linear(K, hidden); None = already `hidden` wide
    tokens: Optional[int] = None  # tokens per clip if fixed by the model (HOI); None = taken from the input (HHI)
    task_id: Optional[int] = None # row of `task_embed` added to these tokens (task_sinusoid embedding only)
    pos_run: bool = False         # the sinusoid positions CONTINUE the previous segment's run instead of restarting at 0
                                  # (HOI EgoT2-g: the action task = slow8 | fast8 sharing positions 0..15)


@dataclass(frozen=True)
class TranslatorSpec:
    family: str                   # "hhi_ttm" | "hhi_asd" | "hoi_pnr" | "hoi_lta"
    hidden: int
    heads: int
    ffn: int
    layers: int
    segments: Tuple[Segment, ...]
    embed: str                    # "task_sinusoid" (HHI) | "learned_pe" (HOI)
    encoder_prefix: str           # "transformer_encoder." (HHI) | "transformer." (HOI)
    head: str                     # "pool_ln_linear" | "tokens" | "pool_multilinear"
    n_out: int                    # logits per clip (2 / 16 / Z*593); for "tokens": hidden
    head_ln_shared: bool = False  # HOI PNR: linear_head.0 IS self.ln (one gamma/beta, two uses)
    p_layer: float = 0.0          # dropout inside every encoder layer (4 sites)
    p_embed: float = 0.0          # dropout after the embedding add (HHI PositionalEncoding: fixed 0.1)
    p_feat: float = 0.0           # dropout on projected features before the LN (HOI self.dp)
    p_head: float = 0.0           # dropout before the head projections (LTA MultiTaskHead)
    n_task_embed: int = 0
    head_groups: Tuple[int, ...] = ()   # LTA: class-group sizes inside each of the Z heads (115, 478)
    n_heads_out: int = 1                # LTA: Z = number of independent Linear heads
    decoder_layers: int = 0             # EgoT2-g: nn.TransformerDecoder depth (head == "decoder")
    vocab: int = 0                      # EgoT2-g: task-prompt vocabulary size (7)
    g_mode: str = ""                    # EgoT2-g: "lam" | "ttm" | "asd" (which forward of the shared model this spec runs)
    encoder: str = "torch"              # "torch": nn.TransformerEncoderLayer (post-norm, ReLU) | "simple_vit": pre-norm, GELU,
                                        # bias-free attention with `dim_head` independent of `hidden` (HOI/models/pnr/simple_vit.py)
    dim_head: int = 0                   # simple_vit only: inner = heads * dim_head
    embed_ln: bool = True               # False: tokens = projected features + pe, no LayerNorm (2-task simple_vit sibling)
    feat_drop_tokens: int = 0           # > 0: p_feat reaches only the first N tokens of a clip (2-task PNR, FEAT_DROPOUT_MODE > 0)

    @property
    def fixed_tokens(self) -> Optional[int]:
        if any(s.tokens is None for s in self.segments):
            return None
        return sum(s.tokens for s in self.segments)

    def table_runs(self, seg_tokens) -> List[Tuple[int, int]]:
        """(tokens, task_id) of every position run of the task_sinusoid token table: one run per segment, except that a
        segment flagged `pos_run` extends the run of the segment before it (same task row, positions carry on)."""
        runs: List[Tuple[int, int]] = []
        for s, d in zip(self.segments, seg_tokens):
            if s.pos_run and runs:
                assert runs[-1][1] == s.task_id, f"{s.name}: a continued position run keeps its task row"
                runs[-1] = (runs[-1][0] + int(d), runs[-1][1])
            else:
                runs.append((int(d), s.task_id))
        return runs

    # ---- parameter inventory: reference state_dict key -> shape ------------------------
    def param_shapes(self, tokens: Optional[int] = None) -> Dict[str, Tuple[int, ...]]:
        H, FF = self.hidden, self.ffn
        out: Dict[str, Tuple[int, ...]] = {}
        if self.embed == "task_sinusoid" and self.family == "hoi_g":
            # one parameter set behind both modes of the 6-task model (n_task_embed == 4 adds proj_lta)
            out["proj_pnr.weight"], out["proj_pnr.bias"] = (H, 8192), (H,)
            out["proj_oscc.weight"], out["proj_oscc.bias"] = (H, 8192), (H,)
            out["proj_action_slow.weight"], out["proj_action_slow.bias"] = (H, 2048), (H,)
            out["proj_action_fast.weight"], out["proj_action_fast.bias"] = (H, 256), (H,)
            if self.n_task_embed == 4:
                out["proj_lta.weight"], out["proj_lta.bias"] = (H, 2048), (H,)
            out["task_embed"] = (1, self.n_task_embed, H)
        elif self.embed == "task_sinusoid":
            # reference registration order: proj_lam, proj_ttm[, proj_asd], task_embed, ..., ln, linear_head
            names = [s.name for s in self.segments]
            for nm in ("lam", "ttm", "asd"):
                if nm in names or self.family == "hhi_g":      # EgoT2-g: one parameter set behind all three modes
                    out[f"proj_{nm}.weight"] = (H, 256)
                    out[f"proj_{nm}.bias"] = (H,)
            out["task_embed"] = (1, self.n_task_embed, H)
        else:
            T = self.fixed_tokens
            out["pe"] = (1, T, H)
            for s in self.segments:
                if s.proj is not None:
                    out[f"{s.proj}.weight"] = (H, s.in_dim)
                    out[f"{s.proj}.bias"] = (H,)
        if self.family == "hoi_pnr" and (self.embed_ln or self.head_ln_shared):
            out["ln.weight"] = (H,)
            out["ln.bias"] = (H,)
        for i in range(self.layers if self.encoder == "simple_vit" else 0):
            a, f = f"{self.encoder_prefix}layers.{i}.0.", f"{self.encoder_prefix}layers.{i}.1.net."
            inner = self.heads * self.dim_head
            out[a + "norm.weight"], out[a + "norm.bias"] = (H,), (H,)
            out[a + "to_qkv.weight"] = (3 * inner, H)
            out[a + "to_out.weight"] = (H, inner)
            out[f + "0.weight"], out[f + "0.bias"] = (H,), (H,)
            out[f + "1.weight"], out[f + "1.bias"] = (FF, H), (FF,)
            out[f + "3.weight"], out[f + "3.bias"] = (H, FF), (H,)
        for i in range(self.layers if self.encoder == "torch" else 0):
            p = f"{self.encoder_prefix}layers.{i}."
            out[p + "self_attn.in_proj_weight"] = (3 * H, H)
            out[p + "self_attn.in_proj_bias"] = (3 * H,)
            out[p + "self_attn.out_proj.weight"] = (H, H)
            out[p + "self_attn.out_proj.bias"] = (H,)
            out[p + "linear1.weight"] = (FF, H)
            out[p + "linear1.bias"] = (FF,)
            out[p + "linear2.weight"] = (H, FF)
            out[p + "linear2.bias"] = (H,)
            out[p + "norm1.weight"] = (H,)
            out[p + "norm1.bias"] = (H,)
            out[p + "norm2.weight"] = (H,)
            out[p + "norm2.bias"] = (H,)
        for i in range(self.decoder_layers):
            p = f"transformer_decoder.layers.{i}."
            for att in ("self_attn", "multihead_attn"):
                out[p + att + ".in_proj_weight"] = (3 * H, H)
                out[p + att + ".in_proj_bias"] = (3 * H,)
                out[p + att + ".out_proj.weight"] = (H, H)
                out[p + att + ".out_proj.bias"] = (H,)
            out[p + "linear1.weight"] = (FF, H)
            out[p + "linear1.bias"] = (FF,)
            out[p + "linear2.weight"] = (H, FF)
            out[p + "linear2.bias"] = (H,)
            for nrm in ("norm1", "norm2", "norm3"):
                out[p + nrm + ".weight"] = (H,)
                out[p + nrm + ".bias"] = (H,)
        if self.family != "hoi_pnr":
            out["ln.weight"] = (H,)
            out["ln.bias"] = (H,)
        if self.family in ("hhi_g", "hoi_g"):
            out["embedding.weight"] = (self.vocab, H)
            out["fc.weight"] = (self.vocab, H)
            out["fc.bias"] = (self.vocab,)
        if self.family in ("hhi_ttm", "hhi_asd"):
            out["linear_head.0.weight"] = (H,)
            out["linear_head.0.bias"] = (H,)
            out["linear_head.1.weight"] = (2, H)
            out["linear_head.1.bias"] = (2,)
        elif self.family == "hoi_pnr" and self.head == "pool_linear":
            out["linear_head.weight"] = (self.n_out, H)       # 2-task variant: a bare nn.Linear head
            out["linear_head.bias"] = (self.n_out,)
        elif self.family == "hoi_pnr":
            # linear_head.0.* alias ln.* in the reference state_dict (same tensor) unless the head owns its LayerNorm
            if not self.head_ln_shared:
                out["linear_head.0.weight"] = (H,)
                out["linear_head.0.bias"] = (H,)
            out["linear_head.1.weight"] = (self.n_out, H)
            out["linear_head.1.bias"] = (self.n_out,)
        elif self.family == "hoi_ar":
            # linear_head{1,2}.0.* alias ln.* in the reference state_dict (one LayerNorm, three uses)
            for gi, n_cls in enumerate(self.head_groups):
                out[f"linear_head{gi + 1}.1.weight"] = (n_cls, H)
                out[f"linear_head{gi + 1}.1.bias"] = (n_cls,)
        elif self.family == "hoi_lta":
            per = sum(self.head_groups)
            for z in range(self.n_heads_out):
                out[f"head.projections.{z}.weight"] = (per, H)
                out[f"head.projections.{z}.bias"] = (per,)
        return out

    def n_params(self) -> int:
        n = 0
        for shp in self.param_shapes().values():
            k = 1
            for d in shp:
                k *= d
            n += k
        return n

    # ---- algorithmic work per clip --------------------------------------
    def flops_per_clip(self, seg_tokens: Tuple[int, ...], backward: bool = False) -> float:
        H, FF, L = self.hidden, self.ffn, self.layers
        T = sum(seg_tokens)
        proj = sum(2.0 * t * s.in_dim * H for t, s in zip(seg_tokens, self.segments) if s.proj is not None)
        layer = 2.0 * T * H * 3 * H + 4.0 * T * T * H + 2.0 * T * H * H + 4.0 * T * H * FF
        if self.head == "tokens":
            head = 0.0
        else:
            head = 2.0 * H * self.n_out
        rest = L * layer + head
        return 2.0 * proj + 3.0 * rest if backward else proj + rest

    def feature_elems_per_clip(self, seg_tokens: Tuple[int, ...]) -> int:
        return sum(t * s.in_dim for t, s in zip(seg_tokens, self.segments))


# --------------------------------------------------------------------------------------
# factories for the shipped variants
# --------------------------------------------------------------------------------------
def hhi_ttm_spec(hidden=128, heads=4, layers=1, dropout=0.5, three_task=True, ffn=2048) -> TranslatorSpec:
    """TTM-of-interest: token order (ttm id0, lam id1[, asd id2]) — model_taskspecific.py:188-190,238-241."""
    segs = [Segment("ttm", 256, "proj_ttm", None, 0), Segment("lam", 256, "proj_lam", None, 1)]
    if three_task:
        segs.append(Segment("asd", 256, "proj_asd", None, 2))
    return TranslatorSpec("hhi_ttm", hidden, heads, ffn, layers, tuple(segs), "task_sinusoid",
                          "transformer_encoder.", "pool_ln_linear", 2, False, dropout, 0.1, 0.0, 0.0,
                          3 if three_task else 2)


def hhi_asd_spec(hidden=128, heads=4, layers=1, dropout=0.5, ffn=2048) -> TranslatorSpec:
    """ASD-of-interest: token order (asd id2, ttm id0, lam id1), returns the first D tokens —
    HHI/models/asd/model_taskspecific.py:151-157."""
    segs = (Segment("asd", 256, "proj_asd", None, 2), Segment("ttm", 256, "proj_ttm", None, 0),
            Segment("lam", 256, "proj_lam", None, 1))
    return TranslatorSpec("hhi_asd", hidden, heads, ffn, layers, segs, "task_sinusoid",
                          "transformer_encoder.", "tokens", hidden, False, dropout, 0.1, 0.0, 0.0, 3)


def hhi_g_spec(hidden=256, heads=4, layers=3, dropout=0.1, mode="ttm", ffn=2048, vocab=7) -> TranslatorSpec:
    """EgoT2-g TaskTranslationPromptTransformer (HHI/models/multitask/task_prompt_model.py:174-293): encoder over
    (lam id0, ttm id1, asd id2) tokens - or the LAM tokens alone for mode "lam" (:231-234) - and an nn.TransformerDecoder
    over the 2-token task prompt; mode "asd" regroups the memory to 3 tokens per frame (:251-257)."""
    assert mode in ("lam", "ttm", "asd")
    if mode == "lam":
        segs = (Segment("lam", 256, "proj_lam", None, 0),)
    else:
        segs = (Segment("lam", 256, "proj_lam", None, 0), Segment("ttm", 256, "proj_ttm", None, 1),
                Segment("asd", 256, "proj_asd", None, 2))
    return TranslatorSpec("hhi_g", hidden, heads, ffn, layers, segs, "task_sinusoid", "transformer_encoder.", "decoder",
                          vocab, False, dropout, 0.1, 0.0, 0.0, 3, decoder_layers=layers, vocab=vocab, g_mode=mode)


def hoi_g_spec(hidden=256, heads=4, layers=3, dropout=0.1, vocab=600, mode="clip", n_tasks=3, num_input=2,
               ffn=2048) -> TranslatorSpec:
    """HOI EgoT2-g `TaskTranslationPromptTransformer` (HOI/models/multitask/video_model_builder.py:223-275) and the
    `...6Task` sibling (:279-383).  mode "clip" (pnr / oscc / action prompts): encoder over (pnr16 id0, oscc16 id1,
    action id2 = slow8 | fast8 with ONE position run 0..15, encode() :236-243), decoder memory = the clip's 48 tokens.
    mode "lta" (6Task only, n_tasks = 4, :325-345): tokens (pnr, oscc, action [already `hidden` wide], lta) x num_input
    clips with task rows 0..3.  Positional dropout 0.1 (PositionalEncoding, :56), FF 2048 (torch default), xavier init."""
    assert mode in ("clip", "lta") and n_tasks in (3, 4) and (mode == "clip" or n_tasks == 4)
    if mode == "clip":
        segs = (Segment("pnr", 8192, "proj_pnr", 16, 0), Segment("oscc", 8192, "proj_oscc", 16, 1),
                Segment("slow", 2048, "proj_action_slow", 8, 2), Segment("fast", 256, "proj_action_fast", 8, 2, True))
    else:
        n = num_input
        segs = (Segment("pnr", 8192, "proj_pnr", n, 0), Segment("oscc", 8192, "proj_oscc", n, 1),
                Segment("action", hidden, None, n, 2), Segment("lta", 2048, "proj_lta", n, 3))
    return TranslatorSpec("hoi_g", hidden, heads, ffn, layers, segs, "task_sinusoid", "transformer_encoder.", "decoder",
                          vocab, False, dropout, 0.1, 0.0, 0.0, n_tasks, decoder_layers=layers, vocab=vocab, g_mode=mode)


def hoi_pnr_spec(hidden=128, layers=6, n_cls=16, feat_dropout=0.5, tr_dropout=0.1) -> TranslatorSpec:
    """PNR/OSCC EgoT2-s: tokens (pnr16, oscc16, slow8, fast8), nh=8, FF=2H — video_model_transfer_3task.py:219-236."""
    segs = (Segment("pnr", 8192, "proj1", 16), Segment("oscc", 8192, "proj2", 16),
            Segment("slow", 2048, "proj3_slow", 8), Segment("fast", 256, "proj3_fast", 8))
    return TranslatorSpec("hoi_pnr", hidden, 8, 2 * hidden, layers, segs, "learned_pe", "transformer.",
                          "pool_ln_linear", n_cls, True, tr_dropout, 0.0, feat_dropout, 0.0)


def hoi_pnr_vit_spec(n_cls=16) -> TranslatorSpec:
    """simple_vit sibling of the PNR/OSCC translator, `TaskFusionMFTransformer3Task` of the pnr registry
    (HOI/models/pnr/video_model_transfer_3task.py:128-164): the same 48 tokens and the same shared-ln head as the Dropout
    variant, H = 256 fixed, encoder = simple_vit Transformer(dim 256, depth 3, heads 8, dim_head 128, mlp_dim 512); no
    dropout anywhere."""
    segs = (Segment("pnr", 8192, "proj1", 16), Segment("oscc", 8192, "proj2", 16),
            Segment("slow", 2048, "proj3_slow", 8), Segment("fast", 256, "proj3_fast", 8))
    return TranslatorSpec("hoi_pnr", 256, 8, 512, 3, segs, "learned_pe", "transformer.", "pool_ln_linear", n_cls, True,
                          0.0, 0.0, 0.0, 0.0, encoder="simple_vit", dim_head=128)


def hoi_pnr2_vit_spec(n_cls=16) -> TranslatorSpec:
    """2-task simple_vit sibling `TaskFusionMFTransformer` (HOI/models/pnr/video_model_transfer.py:44-67): tokens (pnr16,
    oscc16) = projections + pe with NO token LayerNorm (:63), the same simple_vit Transformer(256, depth 3, 8 x 128, mlp 512),
    head = Sequential(its own LayerNorm, Linear)."""
    segs = (Segment("pnr", 8192, "proj1", 16), Segment("oscc", 8192, "proj2", 16))
    return TranslatorSpec("hoi_pnr", 256, 8, 512, 3, segs, "learned_pe", "transformer.", "pool_ln_linear", n_cls, False,
                          0.0, 0.0, 0.0, 0.0, encoder="simple_vit", dim_head=128, embed_ln=False)


def hoi_pnr2_spec(n_cls=16, tr_dropout=0.1, feat_dropout=0.0, feat_dropout_mode=0) -> TranslatorSpec:
    """2-task PNR/OSCC sibling `TaskFusionMFTransformerDropout` (HOI/models/pnr/video_model_transfer.py:70-105): tokens
    (pnr16, oscc16), H=256, nh=8, FF=2H, 3 layers, shared-nothing head = a bare Linear(H, n_cls) on the mean token (no
    LayerNorm).  FEAT_DROPOUT_MODE = 0 (the shipped default, configs/pnr/defaults.py:240) = no feature dropout; any mode
    > 0 drops the projected PNR features alone (:95-96; the `elif dpmode > 1` branch behind it can never run)."""
    segs = (Segment("pnr", 8192, "proj1", 16), Segment("oscc", 8192, "proj2", 16))
    on = feat_dropout_mode > 0 and feat_dropout > 0.0
    return TranslatorSpec("hoi_pnr", 256, 8, 512, 3, segs, "learned_pe", "transformer.", "pool_linear", n_cls, False,
                          tr_dropout, 0.0, feat_dropout if on else 0.0, 0.0, feat_drop_tokens=16 if on else 0)


def hoi_ar_spec(hidden=128, layers=3, heads=8, dropout=0.1, num_classes=(115, 478), ffn=2048) -> TranslatorSpec:
    """Action-recognition EgoT2-s sibling `TaskFusionMFTransformer3Task` (HOI/models/lta/lta_models_transfer.py:96-137):
    tokens (slow8, fast8, pnr16, oscc16), FF = 2048 (torch default), `ln` shared by the token LayerNorm and BOTH heads
    (linear_head1 -> verbs, linear_head2 -> nouns); shipped: H=128, L=3, p=0.1 (configs/recognition/ts_ar.yaml:53-55)."""
    segs = (Segment("slow", 2048, "proj3_slow", 8), Segment("fast", 256, "proj3_fast", 8),
            Segment("pnr", 8192, "proj1", 16), Segment("oscc", 8192, "proj2", 16))
    return TranslatorSpec("hoi_ar", hidden, heads, ffn, layers, segs, "learned_pe", "transformer.",
                          "pool_ln_multilinear", sum(num_classes), True, dropout, 0.0, 0.0, 0.0, 0, tuple(num_classes), 1)


def hoi_ar2_spec(hidden=128, layers=3, heads=8, dropout=0.1, num_classes=(115, 478), ffn=2048) -> TranslatorSpec:
    """`TaskFusionMFTransformer2TaskAR` (HOI/models/lta/lta_models_transfer.py:169-235): AR from the recognition backbone's
    (slow8, fast8) tokens plus the LTA backbone's 2 clip features = 18 tokens; same shared-ln two-head output as the
    3-task AR translator; all dim>1 parameters xavier-initialised (:211-214)."""
    segs = (Segment("slow", 2048, "proj_slow", 8), Segment("fast", 256, "proj_fast", 8), Segment("lta", 2048, "proj_lta", 2))
    return TranslatorSpec("hoi_ar", hidden, heads, ffn, layers, segs, "learned_pe", "transformer.",
                          "pool_ln_multilinear", sum(num_classes), True, dropout, 0.0, 0.0, 0.0, 0, tuple(num_classes), 1)


def hoi_lta2_spec(hidden=512, layers=1, heads=4, dropout=0.5, num_input_clips=2, num_actions=20,
                  num_classes=(115, 478), head_dropout=0.5, ffn=2048) -> TranslatorSpec:
    """LTA 2-task sibling `TaskFusionMFTransformer2Task` (HOI/models/lta/lta_models_lta_transfer.py:429-526): tokens
    (action, lta) x num_input_clips, the action features arrive `hidden` wide from the SlowFast head, `proj_lta` is a
    Linear(2048, hidden) - or nn.Identity when hidden == 2048 (:441-444; the shipped ts_lta_2task.yaml: H 2048, 4 heads =
    head dim 512, served by the wide-head attention kernels) - same MultiTaskHead as the 4-task translator."""
    n = num_input_clips
    segs = (Segment("action", hidden, None, n), Segment("lta", 2048, None if hidden == 2048 else "proj_lta", n))
    return TranslatorSpec("hoi_lta", hidden, heads, ffn, layers, segs, "learned_pe", "transformer.",
                          "pool_multilinear", num_actions * sum(num_classes), False, dropout, 0.0, 0.0,
                          head_dropout, 0, tuple(num_classes), num_actions)


def hoi_lta_spec(hidden=512, layers=4, heads=8, dropout=0.5, num_input_clips=2, num_actions=20,
                 num_classes=(115, 478), head_dropout=0.5, ffn=2048) -> TranslatorSpec:
    """LTA EgoT2-s: tokens (pnr, oscc, action, lta) x num_input_clips, FF=2048 (torch default),
    head = Z x Linear(H, 593) — lta_models_lta_transfer.py:262-275,306-313,354-363."""
    n = num_input_clips
    segs = (Segment("pnr", 8192, "proj_pnr", n), Segment("oscc", 8192, "proj_oscc", n),
            Segment("action", hidden, None, n), Segment("lta", 2048, "proj_lta", n))
    return TranslatorSpec("hoi_lta", hidden, heads, ffn, layers, segs, "learned_pe", "transformer.",
                          "pool_multilinear", num_actions * sum(num_classes), False, dropout, 0.0, 0.0,
                          head_dropout, 0, tuple(num_classes), num_actions)

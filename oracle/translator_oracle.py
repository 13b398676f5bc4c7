"""TEST INFRASTRUCTURE ONLY — CPU restatement (the parity oracle) of EgoT2's task-translation path.

Plain torch-on-CPU arithmetic written out op by op (matmul / softmax / mean / var), NOT a
wrapper around nn.TransformerEncoder, so that every line can be checked against the reference
lines it restates.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product (`egot2_b200/`) never does.

Pinning: the reference ships no tests and no golden vectors — the oracle is
pinned instead against outputs of the reference classes themselves executed in the build
container (`oracle/ref_shims.py` + `oracle/make_golden.py` → `tests/golden/*.npz`); see
`tests/test_oracle_golden.py` (fixtures, runs anywhere) and
`tests/test_oracle_vs_reference.py` (live, only where /root/reference exists).

Parameters are passed as a dict keyed by the reference module's own state_dict names.
All citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# primitives (torch.nn semantics the reference relies on; torch 1.12 == 2.x for these)
# --------------------------------------------------------------------------------------
def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def _drop(x: Tensor, p: float, training: bool) -> Tensor:
    return F.dropout(x, p, True) if (training and p > 0.0) else x


def sinusoid_table(n: int, dim: int) -> Tensor:
    """PositionalEncoding buffer `pe` — HHI/models/ttm/model_taskspecific.py:141-147:
    pe[p,2i]=sin(p*exp(-2i*ln(1e4)/dim)), pe[p,2i+1]=cos(same)."""
    pos = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0) / dim))
    pe = torch.zeros(n, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def self_attention(x: Tensor, P: Params, pre: str, n_heads: int, p_drop: float, training: bool) -> Tensor:
    """nn.MultiheadAttention with packed in_proj (q,k,v rows of in_proj_weight in that order),
    scale 1/sqrt(dh) applied to q, softmax over keys, dropout on the probabilities, out_proj.
    x: (B,T,H) batch-first (the reference's seq-first HHI layout is the same math)."""
    B, T, H = x.shape
    dh = H // n_heads
    qkv = linear(x, P[pre + "in_proj_weight"], P[pre + "in_proj_bias"])       # (B,T,3H)
    q, k, v = qkv.split(H, dim=-1)
    q = q.reshape(B, T, n_heads, dh).transpose(1, 2)                              # (B,nh,T,dh)
    k = k.reshape(B, T, n_heads, dh).transpose(1, 2)
    v = v.reshape(B, T, n_heads, dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)                         # (B,nh,T,T)
    a = torch.softmax(s, dim=-1)
    a = _drop(a, p_drop, training)
    o = (a @ v).transpose(1, 2).reshape(B, T, H)
    return linear(o, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])


def encoder_layer(x: Tensor, P: Params, pre: str, n_heads: int, p_drop: float = 0.0,
                  training: bool = False) -> Tensor:
    """nn.TransformerEncoderLayer defaults used by every nn.Transformer translator in the
    reference: post-norm, ReLU, eps 1e-5 (SURVEY.md F2;
    HHI/models/ttm/model_taskspecific.py:168-171, HOI/models/pnr/video_model_transfer_3task.py:231-235):
        x = norm1(x + dropout1(self_attn(x)));  x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))"""
    a = self_attention(x, P, pre + "self_attn.", n_heads, p_drop, training)
    x = layer_norm(x + _drop(a, p_drop, training), P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    h = torch.relu(linear(x, P[pre + "linear1.weight"], P[pre + "linear1.bias"]))
    h = _drop(h, p_drop, training)
    f = linear(h, P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    return layer_norm(x + _drop(f, p_drop, training), P[pre + "norm2.weight"], P[pre + "norm2.bias"])


def encoder(x: Tensor, P: Params, pre: str, n_layers: int, n_heads: int, p_drop: float = 0.0,
            training: bool = False) -> Tensor:
    """nn.TransformerEncoder without a final norm (none of the call sites passes `norm=`)."""
    for i in range(n_layers):
        x = encoder_layer(x, P, f"{pre}layers.{i}.", n_heads, p_drop, training)
    return x


def count_layers(P: Params, pre: str) -> int:
    n = 0
    while f"{pre}layers.{n}.norm1.weight" in P:
        n += 1
    return n


# --------------------------------------------------------------------------------------
# HHI EgoT2-s translators
# --------------------------------------------------------------------------------------
_HHI_TASK_ID = {"ttm": 0, "lam": 1, "asd": 2}   # model_taskspecific.py:238-240 (encode_prepare ids)


def hhi_tokens(P: Params, feats: Dict[str, Tensor], order: Sequence[str], training: bool = False) -> Tensor:
    """proj_k -> shared ln -> + task_embed[k] -> + sinusoidal pe restarting at 0 per task ->
    Dropout(0.1) -> concat over tokens.  HHI/models/ttm/model_taskspecific.py:178-182,188-190
    (2-task), :222-226,238-241 (3-task); HHI/models/asd/model_taskspecific.py:133-137,151-154."""
    H = P["ln.weight"].numel()
    segs = []
    for name in order:
        f = feats[name]                                                     # (B,D_k,256)
        x = linear(f, P[f"proj_{name}.weight"], P[f"proj_{name}.bias"])
        x = layer_norm(x, P["ln.weight"], P["ln.bias"]) + P["task_embed"][:, _HHI_TASK_ID[name], :]
        x = x + sinusoid_table(f.shape[1], H).unsqueeze(0)
        segs.append(_drop(x, 0.1, training))
    return torch.cat(segs, dim=1)                                           # (B,T,H)


def hhi_ttm_forward(P: Params, feats: Dict[str, Tensor], n_heads: int, p_drop: float = 0.0,
                    training: bool = False) -> Tensor:
    """TaskFusionMFTransformer2Task / 3Task (TTM of interest) -> (B,2) logits.
    HHI/models/ttm/model_taskspecific.py:184-194 and :228-245.  Token order (ttm, lam[, asd])."""
    order = ("ttm", "lam", "asd") if "proj_asd.weight" in P else ("ttm", "lam")
    x = hhi_tokens(P, feats, order, training)
    x = encoder(x, P, "transformer_encoder.", count_layers(P, "transformer_encoder."), n_heads, p_drop, training)
    g = x.mean(dim=1)
    g = layer_norm(g, P["linear_head.0.weight"], P["linear_head.0.bias"])
    return linear(g, P["linear_head.1.weight"], P["linear_head.1.bias"])


def hhi_asd_forward(P: Params, feats: Dict[str, Tensor], n_heads: int, p_drop: float = 0.0,
                    training: bool = False) -> Tensor:
    """ASD-of-interest TaskFusionMFTransformer3Task -> (B*D, H): the first D (= asd) encoded
    tokens of each clip, no pooling, no head.  HHI/models/asd/model_taskspecific.py:139-158;
    token order (asd, ttm, lam)."""
    x = hhi_tokens(P, feats, ("asd", "ttm", "lam"), training)
    x = encoder(x, P, "transformer_encoder.", count_layers(P, "transformer_encoder."), n_heads, p_drop, training)
    D = feats["asd"].shape[1]
    return x[:, :D, :].reshape(-1, x.shape[-1])


def loss_av(P: Params, x: Tensor, labels: Tensor):
    """lossAV — HHI/tasks/asd/loss.py:11-30: FC(H->2), CE(weight [1,4]), softmax score,
    round(softmax)[:,1] label, correct count.  `P` holds 'FC.weight', 'FC.bias'."""
    z = linear(x, P["FC.weight"], P["FC.bias"])
    loss = ce_loss(z, labels, torch.tensor([1.0, 4.0]))
    score = torch.softmax(z, dim=-1)
    label = torch.round(score)[:, 1]
    return loss, score, label, (label == labels).sum().float()


# --------------------------------------------------------------------------------------
# HOI EgoT2-s translators
# --------------------------------------------------------------------------------------
def pool_slowfast(slow5: Tensor, fast5: Tensor):
    """AdaptiveAvgPool3d((None,1,1)) / ((8,1,1)) + squeeze + permute —
    HOI/models/pnr/video_model_transfer_3task.py:226-227,245-247.
    slow5 (B,2048,8,h,w) -> (B,8,2048); fast5 (B,256,32,h,w) -> (B,8,256) (mean over h,w and
    over groups of 32/8 consecutive frames)."""
    slow = slow5.mean(dim=(-1, -2)).permute(0, 2, 1)
    B, C, Tf = fast5.shape[:3]
    fast = fast5.mean(dim=(-1, -2)).reshape(B, C, 8, Tf // 8).mean(dim=-1).permute(0, 2, 1)
    return slow, fast


def hoi_pnr_forward(P: Params, pnr: Tensor, oscc: Tensor, slow: Tensor, fast: Tensor, n_heads: int = 8,
                    p_feat: float = 0.0, p_drop: float = 0.0, training: bool = False) -> Tensor:
    """TaskFusionMFTransformer3TaskDropout -> (B, n_cls) logits (before the unsqueeze).
    HOI/models/pnr/video_model_transfer_3task.py:238-258.  Token order (pnr, oscc, slow, fast);
    learned pe; `ln` shared between the token LN and the head LN (F5).
    slow/fast may be the raw 5-D SlowFast maps or the already pooled (B,8,C) features."""
    if slow.dim() == 5:
        slow, fast = pool_slowfast(slow, fast)
    z = torch.cat([
        _drop(linear(pnr, P["proj1.weight"], P["proj1.bias"]), p_feat, training),
        _drop(linear(oscc, P["proj2.weight"], P["proj2.bias"]), p_feat, training),
        _drop(linear(slow, P["proj3_slow.weight"], P["proj3_slow.bias"]), p_feat, training),
        _drop(linear(fast, P["proj3_fast.weight"], P["proj3_fast.bias"]), p_feat, training)], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    g = layer_norm(x.mean(dim=1), P["ln.weight"], P["ln.bias"])
    return linear(g, P["linear_head.1.weight"], P["linear_head.1.bias"])


def hoi_pnr2_forward(P: Params, pnr: Tensor, oscc: Tensor, n_heads: int = 8, p_drop: float = 0.0,
                     training: bool = False) -> Tensor:
    """TaskFusionMFTransformerDropout (the 2-task sibling) -> (B, n_cls) logits before the unsqueeze.
    HOI/models/pnr/video_model_transfer.py:91-105 with FEAT_DROPOUT_MODE = 0 (configs/pnr/defaults.py:240): tokens
    (pnr, oscc) -> ln + learned pe -> 3-layer encoder -> mean over tokens -> bare Linear head (no LayerNorm)."""
    z = torch.cat([linear(pnr, P["proj1.weight"], P["proj1.bias"]), linear(oscc, P["proj2.weight"], P["proj2.bias"])], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    return linear(x.mean(dim=1), P["linear_head.weight"], P["linear_head.bias"])


def hoi_ar_forward(P: Params, slow: Tensor, fast: Tensor, pnr: Tensor, oscc: Tensor, n_heads: int = 8,
                   p_drop: float = 0.0, training: bool = False) -> Tensor:
    """Action-recognition TaskFusionMFTransformer3Task -> (B, 115 + 478) = [verb logits | noun logits].
    HOI/models/lta/lta_models_transfer.py:124-137.  Token order (slow, fast, pnr, oscc); learned pe; the ONE LayerNorm
    `ln` normalises the tokens and sits in front of both heads (linear_head{1,2} = Sequential(self.ln, Linear), :119-121).
    slow/fast may be the raw 5-D SlowFast maps or the already pooled (B,8,C) features."""
    if slow.dim() == 5:
        slow, fast = pool_slowfast(slow, fast)
    z = torch.cat([linear(slow, P["proj3_slow.weight"], P["proj3_slow.bias"]),
                   linear(fast, P["proj3_fast.weight"], P["proj3_fast.bias"]),
                   linear(pnr, P["proj1.weight"], P["proj1.bias"]),
                   linear(oscc, P["proj2.weight"], P["proj2.bias"])], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    g = layer_norm(x.mean(dim=1), P["ln.weight"], P["ln.bias"])
    return torch.cat([linear(g, P["linear_head1.1.weight"], P["linear_head1.1.bias"]),
                      linear(g, P["linear_head2.1.weight"], P["linear_head2.1.bias"])], dim=-1)


def hoi_ar2_forward(P: Params, slow: Tensor, fast: Tensor, lta: Tensor, n_heads: int = 8, p_drop: float = 0.0,
                    training: bool = False) -> Tensor:
    """TaskFusionMFTransformer2TaskAR -> (B, 115 + 478).  HOI/models/lta/lta_models_transfer.py:227-235: tokens
    (slow8, fast8, lta2) -> ln + pe -> encoder -> mean -> shared ln -> two Linear heads."""
    if slow.dim() == 5:
        slow, fast = pool_slowfast(slow, fast)
    z = torch.cat([linear(slow, P["proj_slow.weight"], P["proj_slow.bias"]),
                   linear(fast, P["proj_fast.weight"], P["proj_fast.bias"]),
                   linear(lta, P["proj_lta.weight"], P["proj_lta.bias"])], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    g = layer_norm(x.mean(dim=1), P["ln.weight"], P["ln.bias"])
    return torch.cat([linear(g, P["linear_head1.1.weight"], P["linear_head1.1.bias"]),
                      linear(g, P["linear_head2.1.weight"], P["linear_head2.1.bias"])], dim=-1)


def ar_loss(stacked: Tensor, labels: Tensor, num_classes=(115, 478)) -> Tensor:
    """RecognitionTask2Loader.training_step (HOI/tasks/lta/long_term_anticipation_taskspecfic.py:26-33):
    CE(verb logits, labels[:, 0]) + CE(noun logits, labels[:, 1])."""
    v, n = torch.split(stacked, list(num_classes), dim=-1)
    return torch.nn.functional.cross_entropy(v, labels[:, 0]) + torch.nn.functional.cross_entropy(n, labels[:, 1])


def hoi_lta_forward(P: Params, pnr: Tensor, oscc: Tensor, action: Tensor, lta: Tensor, n_heads: int = 8,
                    p_drop: float = 0.0, p_head: float = 0.0, training: bool = False,
                    eval_softmax: bool = False) -> Tensor:
    """TaskFusionMFTransformerLTA4Task -> (B, Z, n_verbs+n_nouns) stacked head outputs.
    HOI/models/lta/lta_models_lta_transfer.py:354-363 + decode :348-352 +
    MultiTaskHead HOI/models/lta/head_helper.py:262-290.  Inputs are the per-input-clip
    features: pnr/oscc (B,2,8192) (already temporally averaged, :339-346), action (B,2,H),
    lta (B,2,2048).  Token order (pnr, oscc, action, lta).  eval_softmax=True reproduces the
    eval-mode Softmax(dim=4) over all classes (head_helper.py:284-286)."""
    z = torch.cat([linear(pnr, P["proj_pnr.weight"], P["proj_pnr.bias"]),
                   linear(oscc, P["proj_oscc.weight"], P["proj_oscc.bias"]),
                   action,
                   linear(lta, P["proj_lta.weight"], P["proj_lta.bias"])], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    g = _drop(x.mean(dim=1), p_head, training)
    outs = []
    z_idx = 0
    while f"head.projections.{z_idx}.weight" in P:
        o = linear(g, P[f"head.projections.{z_idx}.weight"], P[f"head.projections.{z_idx}.bias"])
        outs.append(torch.softmax(o, dim=-1) if eval_softmax else o)
        z_idx += 1
    return torch.stack(outs, dim=1)


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# EgoT2-g: TaskTranslationPromptTransformer — HHI/models/multitask/task_prompt_model.py:174-293
# --------------------------------------------------------------------------------------
def mha(q_in: Tensor, kv_in: Tensor, P: Params, pre: str, n_heads: int, mask: Optional[Tensor], p_drop: float,
        training: bool) -> Tensor:
    """nn.MultiheadAttention(query, key=value=kv_in), written batch-first (N,S,H) / (N,M,H) (same math as the
    reference's seq-first call): packed in_proj rows q|k|v, softmax(q k^T / sqrt(dh) + mask), dropout on the
    probabilities, out_proj."""
    N, S, H = q_in.shape
    M = kv_in.shape[1]
    dh = H // n_heads
    W, b = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = linear(q_in, W[:H], b[:H]).reshape(N, S, n_heads, dh).transpose(1, 2)
    k = linear(kv_in, W[H:2 * H], b[H:2 * H]).reshape(N, M, n_heads, dh).transpose(1, 2)
    v = linear(kv_in, W[2 * H:], b[2 * H:]).reshape(N, M, n_heads, dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    p = _drop(torch.softmax(s, dim=-1), p_drop, training)
    o = (p @ v).transpose(1, 2).reshape(N, S, H)
    return linear(o, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])


def decoder_layer(y: Tensor, mem: Tensor, P: Params, pre: str, n_heads: int, mask: Tensor, p_drop: float,
                  training: bool) -> Tensor:
    """nn.TransformerDecoderLayer, post-norm, ReLU (CustomDecoderLayer only forces need_weights=True, :163-172)."""
    y = layer_norm(y + _drop(mha(y, y, P, pre + "self_attn.", n_heads, mask, p_drop, training), p_drop, training),
                   P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    y = layer_norm(y + _drop(mha(y, mem, P, pre + "multihead_attn.", n_heads, None, p_drop, training), p_drop, training),
                   P[pre + "norm2.weight"], P[pre + "norm2.bias"])
    f = linear(_drop(torch.relu(linear(y, P[pre + "linear1.weight"], P[pre + "linear1.bias"])), p_drop, training),
               P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    return layer_norm(y + _drop(f, p_drop, training), P[pre + "norm3.weight"], P[pre + "norm3.bias"])


_G_TASK_ID = {"lam": 0, "ttm": 1, "asd": 2}     # task_prompt_model.py:234,247-249


def hhi_g_forward(P: Params, feats: Dict[str, Tensor], target_in: Tensor, mode: str, n_heads: int, p_drop: float = 0.0,
                  training: bool = False) -> Tensor:
    """forward(video, ..., target, task) -> (rows, vocab, seq) logits (:271-275).
    encode (:230-258): 'lam' uses the LAM tokens only (task id 0); otherwise cat(lam id0, ttm id1, asd id2);
    'asd' regroups the encoder output to a 3-token memory per frame.  decode (:260-269): Embedding*sqrt(H) + PE(+dropout 0.1)
    -> TransformerDecoder with the causal mask -> fc."""
    order = ("lam",) if mode == "lam" else ("lam", "ttm", "asd")
    H = P["ln.weight"].shape[0]
    toks = []
    for nm in order:
        f = feats[nm]
        x = linear(f, P[f"proj_{nm}.weight"], P[f"proj_{nm}.bias"])
        x = layer_norm(x, P["ln.weight"], P["ln.bias"]) + P["task_embed"][:, _G_TASK_ID[nm], :]
        x = x + sinusoid_table(f.shape[1], H).unsqueeze(0)
        toks.append(_drop(x, 0.1, training))
    x = torch.cat(toks, dim=1)                                    # (B, T, H)
    mem = encoder(x, P, "transformer_encoder.", count_layers(P, "transformer_encoder."), n_heads, p_drop, training)
    if mode == "asd":                                             # (B, 3T, H) -> (B*T, 3, H)
        B, T3, _ = mem.shape
        T = T3 // 3
        mem = torch.stack([mem[:, 0:T].reshape(-1, H), mem[:, T:2 * T].reshape(-1, H), mem[:, 2 * T:3 * T].reshape(-1, H)], dim=1)
    S = target_in.shape[1]
    y = P["embedding.weight"][target_in] * math.sqrt(H)           # (rows, S, H)
    y = _drop(y + sinusoid_table(S, H).unsqueeze(0), 0.1, training)
    mask = torch.full((S, S), float("-inf")).triu(1)
    for i in range(count_layers(P, "transformer_decoder.")):
        y = decoder_layer(y, mem, P, f"transformer_decoder.layers.{i}.", n_heads, mask, p_drop, training)
    out = linear(y, P["fc.weight"], P["fc.bias"])                 # (rows, S, V)
    return out.transpose(1, 2)                                    # (rows, V, S) like the reference's permute(1, 2, 0)


def hoi_lta2_forward(P: Params, action: Tensor, lta: Tensor, n_heads: int = 4, p_drop: float = 0.0, p_head: float = 0.0,
                     training: bool = False, eval_softmax: bool = False) -> Tensor:
    """LTA TaskFusionMFTransformer2Task -> stacked head logits (B, Z, 593).
    HOI/models/lta/lta_models_lta_transfer.py:510-518: tokens (action x n, proj_lta(lta) x n) -> ln + pe -> encoder ->
    mean -> MultiTaskHead (Z x [Dropout -> Linear(H, 593)], softmax over 593 in eval unless TEST.NO_ACT)."""
    # :441-444 - proj_lta is nn.Identity when the translator is 2048 wide (no proj_lta.* in the state_dict)
    z = torch.cat([action, linear(lta, P["proj_lta.weight"], P["proj_lta.bias"]) if "proj_lta.weight" in P else lta], dim=1)
    x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
    x = encoder(x, P, "transformer.", count_layers(P, "transformer."), n_heads, p_drop, training)
    g = x.mean(dim=1)
    outs = []
    zi = 0
    while f"head.projections.{zi}.weight" in P:
        outs.append(linear(_drop(g, p_head, training), P[f"head.projections.{zi}.weight"], P[f"head.projections.{zi}.bias"]))
        zi += 1
    out = torch.stack(outs, dim=1)
    return torch.softmax(out, dim=-1) if eval_softmax else out


# --------------------------------------------------------------------------------------
# HOI EgoT2-g (engine path: egot2_b200.hoi.multitask, specs.hoi_g_spec)
# --------------------------------------------------------------------------------------
def _hoi_g_tokens(P: Params, tasks: Sequence[Tensor], n_heads: int, p_drop: float, training: bool) -> Tensor:
    """encode_prepare per TASK (HOI/models/multitask/video_model_builder.py:144-148: ln + task_embed[id] + sinusoid that
    restarts at 0 for every task + Dropout(0.1)), concat along tokens, nn.TransformerEncoder."""
    H = P["ln.weight"].shape[0]
    toks = []
    for k, f in enumerate(tasks):
        x = layer_norm(f, P["ln.weight"], P["ln.bias"]) + P["task_embed"][:, k, :]
        x = x + sinusoid_table(f.shape[1], H).unsqueeze(0)
        toks.append(_drop(x, 0.1, training))
    x = torch.cat(toks, dim=1)
    return encoder(x, P, "transformer_encoder.", count_layers(P, "transformer_encoder."), n_heads, p_drop, training)


def hoi_g_encode(P: Params, pnr: Tensor, oscc: Tensor, slow: Tensor, fast: Tensor, n_heads: int, p_drop: float = 0.0,
                 training: bool = False) -> Tensor:
    """TaskTranslationPromptTransformer.encode (HOI/models/multitask/video_model_builder.py:228-246; same lines of the
    6-task class for the pnr / oscc / action prompts, :333-343) -> memory (B,48,H).  Three TASKS: pnr (id 0), oscc (id 1)
    and action (id 2) = cat(proj_action_slow(slow8), proj_action_fast(fast8)): the action task's 16 tokens share one
    position run."""
    if slow.dim() == 5:
        slow, fast = pool_slowfast(slow, fast)
    tasks = [linear(pnr, P["proj_pnr.weight"], P["proj_pnr.bias"]),
             linear(oscc, P["proj_oscc.weight"], P["proj_oscc.bias"]),
             torch.cat([linear(slow, P["proj_action_slow.weight"], P["proj_action_slow.bias"]),
                        linear(fast, P["proj_action_fast.weight"], P["proj_action_fast.bias"])], dim=1)]
    return _hoi_g_tokens(P, tasks, n_heads, p_drop, training)


def hoi_g_lta_encode(P: Params, pnr: Tensor, oscc: Tensor, action: Tensor, lta: Tensor, n_heads: int, p_drop: float = 0.0,
                     training: bool = False) -> Tensor:
    """TaskTranslationPromptTransformer6Task.encode for the 'lta' prompts (:325-339): per input clip one temporal-mean
    PNR and OSCC feature (8192), one recognition feature (already H wide: the SlowFast head projects to hidden_dim) and
    one LTA backbone feature (2048); four tasks (ids 0..3) of num_input tokens each -> memory (B, 4*num_input, H)."""
    tasks = [linear(pnr, P["proj_pnr.weight"], P["proj_pnr.bias"]),
             linear(oscc, P["proj_oscc.weight"], P["proj_oscc.bias"]),
             action,
             linear(lta, P["proj_lta.weight"], P["proj_lta.bias"])]
    return _hoi_g_tokens(P, tasks, n_heads, p_drop, training)


def hoi_g_lta_forward(P: Params, pnr: Tensor, oscc: Tensor, action: Tensor, lta: Tensor, target_in: Tensor, n_heads: int,
                      p_drop: float = 0.0, training: bool = False) -> Tensor:
    """6Task forward(video_pnr, video_ac, target, 'lta*') -> (B, V, S) logits (:347-350)."""
    mem = hoi_g_lta_encode(P, pnr, oscc, action, lta, n_heads, p_drop, training)
    return hoi_g_decode(P, mem, target_in, n_heads, p_drop, training).transpose(1, 2)


def hoi_g_decode(P: Params, mem: Tensor, target_in: Tensor, n_heads: int, p_drop: float = 0.0, training: bool = False) -> Tensor:
    """decode (:150-160): Embedding * sqrt(H) + PE (+ dropout 0.1) -> TransformerDecoder with the causal mask -> fc.
    Returns (B, S, V)."""
    H = P["ln.weight"].shape[0]
    S = target_in.shape[1]
    y = P["embedding.weight"][target_in] * math.sqrt(H)
    y = _drop(y + sinusoid_table(S, H).unsqueeze(0), 0.1, training)
    mask = torch.full((S, S), float("-inf")).triu(1)
    for i in range(count_layers(P, "transformer_decoder.")):
        y = decoder_layer(y, mem, P, f"transformer_decoder.layers.{i}.", n_heads, mask, p_drop, training)
    return linear(y, P["fc.weight"], P["fc.bias"])


def hoi_g_forward(P: Params, pnr: Tensor, oscc: Tensor, slow: Tensor, fast: Tensor, target_in: Tensor, n_heads: int,
                  p_drop: float = 0.0, training: bool = False) -> Tensor:
    """forward(video_pnr, video_ac, target) -> (B, V, S) logits (:248-251)."""
    mem = hoi_g_encode(P, pnr, oscc, slow, fast, n_heads, p_drop, training)
    return hoi_g_decode(P, mem, target_in, n_heads, p_drop, training).transpose(1, 2)


def hoi_g_predict_ac(P: Params, pnr: Tensor, oscc: Tensor, slow: Tensor, fast: Tensor, start_token: int, n_heads: int) -> Tensor:
    """predict_ac (:264-275): greedy autoregressive decoding of a fixed 3-token sequence [action, verb, noun]; returns the two
    generated vocabulary indices (B, 2)."""
    mem = hoi_g_encode(P, pnr, oscc, slow, fast, n_heads)
    B = pnr.shape[0]
    toks = torch.ones((B, 3), dtype=torch.int64)
    toks[:, 0] = start_token
    for sy in range(1, 3):
        out = hoi_g_decode(P, mem, toks[:, :sy], n_heads)            # (B, sy, V)
        toks[:, sy] = out[:, -1].argmax(dim=-1)
    return toks[:, 1:]


# --------------------------------------------------------------------------------------
# simple_vit translators (engine path: csrc/vit.cu, specs.hoi_pnr_vit_spec / hoi_pnr2_vit_spec)
# --------------------------------------------------------------------------------------
def simple_vit_transformer(x: Tensor, P: Params, pre: str, heads: int = 8) -> Tensor:
    """HOI/models/pnr/simple_vit.py:55-107 `Transformer`: depth x [x = Attention(x) + x ; x = FeedForward(x) + x], both
    PRE-norm; Attention = LayerNorm -> bias-free to_qkv (dim -> 3*heads*dim_head, dim_head independent of dim) ->
    softmax(q k^T * dim_head^-0.5) v -> bias-free to_out; FeedForward = LayerNorm -> Linear -> exact (erf) GELU -> Linear.
    No dropout anywhere, no final norm.  State_dict keys: `{pre}layers.{i}.0.{norm,to_qkv,to_out}`, `{pre}layers.{i}.1.net.{0,1,3}`."""
    i = 0
    while f"{pre}layers.{i}.0.norm.weight" in P:
        a, f = f"{pre}layers.{i}.0.", f"{pre}layers.{i}.1.net."
        h = layer_norm(x, P[a + "norm.weight"], P[a + "norm.bias"])
        qkv = linear(h, P[a + "to_qkv.weight"], None)
        B, N, inner3 = qkv.shape
        dh = inner3 // 3 // heads
        q, k, v = (t.reshape(B, N, heads, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        p = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
        o = (p @ v).transpose(1, 2).reshape(B, N, heads * dh)
        x = linear(o, P[a + "to_out.weight"], None) + x
        h = layer_norm(x, P[f + "0.weight"], P[f + "0.bias"])
        h = torch.nn.functional.gelu(linear(h, P[f + "1.weight"], P[f + "1.bias"]))
        x = linear(h, P[f + "3.weight"], P[f + "3.bias"]) + x
        i += 1
    return x


def hoi_pnr_vit_forward(P: Params, pnr: Tensor, oscc: Tensor, slow: Optional[Tensor] = None, fast: Optional[Tensor] = None) -> Tensor:
    """simple_vit siblings -> (B, n_cls) logits before the unsqueeze.
    3-task `TaskFusionMFTransformer3Task` (HOI/models/pnr/video_model_transfer_3task.py:128-164): tokens (pnr, oscc, slow, fast)
    -> ln + pe -> simple_vit Transformer(dim 256, depth 3, heads 8, dim_head 128, mlp 512) -> mean -> Sequential(self.ln, Linear)
    (the SAME ln as the token LayerNorm).  2-task `TaskFusionMFTransformer` (video_model_transfer.py:44-67): tokens (pnr, oscc),
    NO token LayerNorm (feat = cat(proj1, proj2) + pe), head = Sequential(its own LayerNorm, Linear)."""
    if slow is not None:
        if slow.dim() == 5:
            slow, fast = pool_slowfast(slow, fast)
        z = torch.cat([linear(pnr, P["proj1.weight"], P["proj1.bias"]), linear(oscc, P["proj2.weight"], P["proj2.bias"]),
                       linear(slow, P["proj3_slow.weight"], P["proj3_slow.bias"]),
                       linear(fast, P["proj3_fast.weight"], P["proj3_fast.bias"])], dim=1)
        x = layer_norm(z, P["ln.weight"], P["ln.bias"]) + P["pe"]
        x = simple_vit_transformer(x, P, "transformer.")
        g = layer_norm(x.mean(dim=1), P["ln.weight"], P["ln.bias"])
    else:
        z = torch.cat([linear(pnr, P["proj1.weight"], P["proj1.bias"]), linear(oscc, P["proj2.weight"], P["proj2.bias"])], dim=1)
        x = simple_vit_transformer(z + P["pe"], P, "transformer.")
        g = layer_norm(x.mean(dim=1), P["linear_head.0.weight"], P["linear_head.0.bias"])
    return linear(g, P["linear_head.1.weight"], P["linear_head.1.bias"])


def ce_loss(logits: Tensor, target: Tensor, weight: Optional[Tensor] = None) -> Tensor:
    """nn.CrossEntropyLoss(weight=w), reduction='mean':  sum_b w[y_b]*nll_b / sum_b w[y_b].
    TTM weight [0.266,0.734]: HHI/configs/ttm/config.py:36, HHI/tasks/ttm/video_task.py:23-24."""
    lse = torch.logsumexp(logits, dim=-1)
    nll = lse - logits.gather(-1, target.unsqueeze(-1)).squeeze(-1)
    if weight is None:
        return nll.mean()
    w = weight.to(logits.dtype)[target]
    return (w * nll).sum() / w.sum()


def bce_sigmoid_loss(logits: Tensor, onehot: Tensor) -> Tensor:
    """PNR: nn.BCELoss(mean)(sigmoid(logits), onehot) — HOI/tasks/pnr/video_taskspecific_pnr.py:29-31.
    (BCELoss clamps log terms at -100.)"""
    p = torch.sigmoid(logits)
    lp = torch.clamp(torch.log(p), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - p), min=-100.0)
    return -(onehot * lp + (1.0 - onehot) * l1p).mean()


def lta_loss(stacked: Tensor, labels: Tensor, num_classes=(115, 478)) -> Tensor:
    """sum over heads (verb, noun) and the Z future steps of mean-CE —
    HOI/tasks/lta/long_term_anticipation_taskspecfic.py:177-183.  stacked (B,Z,593), labels (B,Z,2)."""
    parts = torch.split(stacked, list(num_classes), dim=-1)
    loss = stacked.new_zeros(())
    for h, part in enumerate(parts):
        for z in range(part.shape[1]):
            loss = loss + ce_loss(part[:, z], labels[:, z, h])
    return loss


def keyframe_index(logits: Tensor) -> Tensor:
    """argmax over the 16 PNR logits — HOI/evaluation/pnr/metrics.py:56, HOI/submission/eval_pnr.py:23."""
    return logits.argmax(dim=-1)

"""TEST INFRASTRUCTURE ONLY - a CPU stand-in for the forward entry points of libegot2's C ABI (fp32, eval mode).

`install(monkeypatch)` replaces `egot2_b200._lib.call` by a dispatcher that decodes the SAME arguments the product code
hands to the library (descriptor structs, raw pointers into its own buffers) and computes each stage with the oracle's
ops, writing the results through those pointers.  Driving the public modules through it on CPU checks the HOST WIRING of
a case - which parameter lands in which descriptor field, segment offsets, token-table runs, the decoder's memory
mapping, buffer chaining between stages, greedy decoding - against the committed goldens.  It says nothing about the
CUDA kernels (the `-m gpu` parity tests do); cases whose GPU parity is green double as the check that this emulator
reads the ABI the way the kernels do.  The backward entry points re-run their stage's forward under autograd from the
inputs the ABI hands them and accumulate the gradients through the gradient pointers (fp32, eval mode)."""
import ctypes as C
import math

import numpy as np
import torch

from egot2_b200 import _lib as L
from egot2_b200 import engine as E
from egot2_b200 import modules as M
from oracle import translator_oracle as O


def _f32(ptr, *shape):
    n = int(np.prod(shape))
    arr = np.ctypeslib.as_array(C.cast(C.c_void_p(ptr), C.POINTER(C.c_float)), shape=(n,))
    return torch.from_numpy(arr).view(*shape)


def _i64(ptr, *shape):
    n = int(np.prod(shape))
    arr = np.ctypeslib.as_array(C.cast(C.c_void_p(ptr), C.POINTER(C.c_int64)), shape=(n,))
    return torch.from_numpy(arr).view(*shape)


def _st(arg, typ):
    return C.cast(arg, C.POINTER(typ)).contents


def _need_fp32(dtype):
    assert dtype == L.F32, "the ABI emulator covers the fp32 mode only"


# ---------------------------------------------------------------------------------------------------- stages
def tok_table_fwd(task_embed, pe, pe_len, n_seg, segs, ids, H, table, st):
    te = _f32(task_embed, max(ids[:n_seg]) + 1, H)
    pe_t = _f32(pe, pe_len, H)
    T = sum(segs[:n_seg])
    out = _f32(table, T, H)
    off = 0
    for k in range(n_seg):
        out[off:off + segs[k]] = te[ids[k]] + pe_t[:segs[k]]
        off += segs[k]


def embed_fwd(desc, ein, eout, ws, ws_bytes, st):
    d, i, o = _st(desc, L.EmbedDesc), _st(ein, L.EmbedIn), _st(eout, L.EmbedOut)
    _need_fp32(d.dtype)
    assert not d.training and d.feat_dtype == L.F32
    z = _f32(o.z, d.B, d.T, d.H)
    for k in range(d.n_seg):
        Dk, Kk, off = d.seg_tokens[k], d.seg_in_dim[k], d.seg_offset[k]
        f = _f32(i.feat[k], d.B, Dk, Kk)
        if d.seg_has_proj[k]:
            z[:, off:off + Dk] = O.linear(f, _f32(i.proj_w[k], d.H, Kk), _f32(i.proj_b[k], d.H))
        else:
            z[:, off:off + Dk] = f
    table = _f32(i.tok_table, d.T, d.H)
    x = z + table if d.no_ln else O.layer_norm(z, _f32(i.ln_g, d.H), _f32(i.ln_b, d.H), d.ln_eps) + table
    _f32(o.x, d.B, d.T, d.H).copy_(x)


def encoder_layer_fwd(desc, params, x_in, x_out, saved, ws, ws_bytes, st):
    d, p = _st(desc, L.LayerDesc), _st(params, L.LayerParams)
    _need_fp32(d.dtype)
    assert not d.training
    H, FF = d.H, d.FF
    P = {"self_attn.in_proj_weight": _f32(p.in_proj_w, 3 * H, H), "self_attn.in_proj_bias": _f32(p.in_proj_b, 3 * H),
         "self_attn.out_proj.weight": _f32(p.out_proj_w, H, H), "self_attn.out_proj.bias": _f32(p.out_proj_b, H),
         "linear1.weight": _f32(p.lin1_w, FF, H), "linear1.bias": _f32(p.lin1_b, FF),
         "linear2.weight": _f32(p.lin2_w, H, FF), "linear2.bias": _f32(p.lin2_b, H),
         "norm1.weight": _f32(p.norm1_g, H), "norm1.bias": _f32(p.norm1_b, H),
         "norm2.weight": _f32(p.norm2_g, H), "norm2.bias": _f32(p.norm2_b, H)}
    x = _f32(x_in, d.B, d.T, H)
    _f32(x_out, d.B, d.T, H).copy_(O.encoder_layer(x, P, "", d.heads))


def vit_layer_fwd(desc, params, x_in, x_out, saved, st):
    d, p = _st(desc, L.VitDesc), _st(params, L.VitParams)
    _need_fp32(d.dtype)
    D, inner, mlp = d.D, d.heads * d.dim_head, d.mlp
    P = {"layers.0.0.norm.weight": _f32(p.norm_a_g, D), "layers.0.0.norm.bias": _f32(p.norm_a_b, D),
         "layers.0.0.to_qkv.weight": _f32(p.qkv_w, 3 * inner, D), "layers.0.0.to_out.weight": _f32(p.out_w, D, inner),
         "layers.0.1.net.0.weight": _f32(p.norm_f_g, D), "layers.0.1.net.0.bias": _f32(p.norm_f_b, D),
         "layers.0.1.net.1.weight": _f32(p.ff1_w, mlp, D), "layers.0.1.net.1.bias": _f32(p.ff1_b, mlp),
         "layers.0.1.net.3.weight": _f32(p.ff2_w, D, mlp), "layers.0.1.net.3.bias": _f32(p.ff2_b, D)}
    x = _f32(x_in, d.B, d.T, D)
    _f32(x_out, d.B, d.T, D).copy_(O.simple_vit_transformer(x, P, "", d.heads))


def prompt_embed_fwd(dtype, rows, S, H, tokens, embedding, pe, p_drop, training, seed, y, st):
    _need_fp32(dtype)
    assert not training
    tok = _i64(tokens, rows, S)
    emb = _f32(embedding, int(tok.max()) + 1, H)
    _f32(y, rows, S, H).copy_(emb[tok] * math.sqrt(H) + _f32(pe, S, H).unsqueeze(0))


def decoder_layer_fwd(desc, params, y_in, mem, y_out, saved, st):
    d, p = _st(desc, L.DecoderDesc), _st(params, L.DecoderParams)
    _need_fp32(d.dtype)
    assert not d.training
    H, FF = d.H, d.FF
    P = {}
    for att, (iw, ib, ow, ob) in (("self_attn", (p.sa_in_w, p.sa_in_b, p.sa_out_w, p.sa_out_b)),
                                  ("multihead_attn", (p.ca_in_w, p.ca_in_b, p.ca_out_w, p.ca_out_b))):
        P[att + ".in_proj_weight"], P[att + ".in_proj_bias"] = _f32(iw, 3 * H, H), _f32(ib, 3 * H)
        P[att + ".out_proj.weight"], P[att + ".out_proj.bias"] = _f32(ow, H, H), _f32(ob, H)
    P["linear1.weight"], P["linear1.bias"] = _f32(p.lin1_w, FF, H), _f32(p.lin1_b, FF)
    P["linear2.weight"], P["linear2.bias"] = _f32(p.lin2_w, H, FF), _f32(p.lin2_b, H)
    for i, (g, b) in enumerate(((p.norm1_g, p.norm1_b), (p.norm2_g, p.norm2_b), (p.norm3_g, p.norm3_b)), 1):
        P[f"norm{i}.weight"], P[f"norm{i}.bias"] = _f32(g, H), _f32(b, H)
    memory = _f32(mem, d.mem_rows, H)
    n = torch.arange(d.rows)
    j = torch.arange(d.M)
    # include/egot2.h: row n, key j reads encoder token (n / kv_inner) * kv_outer + j * kv_jstride + (n % kv_inner) * kv_istride
    idx = (n // d.kv_inner)[:, None] * d.kv_outer + j[None, :] * d.kv_jstride + (n % d.kv_inner)[:, None] * d.kv_istride
    mem_rows = memory[idx]                                              # (rows, M, H)
    y = _f32(y_in, d.rows, d.S, H)
    mask = torch.full((d.S, d.S), float("-inf")).triu(1)
    _f32(y_out, d.rows, d.S, H).copy_(O.decoder_layer(y, mem_rows, P, "", d.heads, mask, 0.0, False))


def pool_fwd(dtype, B, T, H, pool, row_tokens, x, out, st):
    _need_fp32(dtype)
    xt = _f32(x, B, T, H)
    if pool:
        _f32(out, B, H).copy_(xt.mean(dim=1))
    else:
        _f32(out, B * row_tokens, H).copy_(xt[:, :row_tokens].reshape(-1, H))


def row_softmax(rows, n, x, out, st):
    _f32(out, rows, n).copy_(torch.softmax(_f32(x, rows, n), dim=-1))


def head_loss_fwd(desc, hin, hout, st):
    d, i, o = _st(desc, L.HeadDesc), _st(hin, L.HeadIn), _st(hout, L.HeadOut)
    _need_fp32(d.dtype)
    assert not (d.training and d.p_head > 0)
    x = _f32(i.x, d.B, d.T, d.H)
    g = x.mean(dim=1) if d.pool else x[:, :d.row_tokens].reshape(-1, d.H)
    rows = g.shape[0]
    if d.use_ln:
        g = O.layer_norm(g, _f32(i.ln_g, d.H), _f32(i.ln_b, d.H), d.ln_eps)
    logits = O.linear(g, _f32(i.w, d.n_out, d.H), _f32(i.b, d.n_out))
    _f32(o.logits, rows, d.n_out).copy_(logits)
    if d.loss == L.LOSS_CE:
        lab = _i64(i.labels, rows)
        w = _f32(i.class_weight, d.n_out) if i.class_weight else None
        _f32(o.loss, 2)[0] = O.ce_loss(logits, lab, w)
    elif d.loss == L.LOSS_BCE_SIGMOID:
        lab = _i64(i.labels, rows)
        _f32(o.loss, 2)[0] = O.bce_sigmoid_loss(logits, torch.nn.functional.one_hot(lab, d.n_out).float())
    elif d.loss == L.LOSS_CE_GROUPS:
        groups = [d.group_size[k] for k in range(d.n_groups)]
        lab = _i64(i.labels, rows, d.sub_rows, d.n_groups)
        _f32(o.loss, 2)[0] = O.lta_loss(logits.view(rows, d.sub_rows, -1), lab, groups)


# ---------------------------------------------------------------------------------------------------- backward stages
def _leaf(t):
    return t.clone().requires_grad_(True)


def _acc(ptr, g, *shape):
    if ptr and g is not None:
        _f32(ptr, *shape).add_(g.reshape(*shape))


def _layer_P(p, H, FF):
    return {"self_attn.in_proj_weight": _f32(p.in_proj_w, 3 * H, H), "self_attn.in_proj_bias": _f32(p.in_proj_b, 3 * H),
            "self_attn.out_proj.weight": _f32(p.out_proj_w, H, H), "self_attn.out_proj.bias": _f32(p.out_proj_b, H),
            "linear1.weight": _f32(p.lin1_w, FF, H), "linear1.bias": _f32(p.lin1_b, FF),
            "linear2.weight": _f32(p.lin2_w, H, FF), "linear2.bias": _f32(p.lin2_b, H),
            "norm1.weight": _f32(p.norm1_g, H), "norm1.bias": _f32(p.norm1_b, H),
            "norm2.weight": _f32(p.norm2_g, H), "norm2.bias": _f32(p.norm2_b, H)}


_LAYER_FIELDS = {"self_attn.in_proj_weight": "in_proj_w", "self_attn.in_proj_bias": "in_proj_b",
                 "self_attn.out_proj.weight": "out_proj_w", "self_attn.out_proj.bias": "out_proj_b", "linear1.weight": "lin1_w",
                 "linear1.bias": "lin1_b", "linear2.weight": "lin2_w", "linear2.bias": "lin2_b", "norm1.weight": "norm1_g",
                 "norm1.bias": "norm1_b", "norm2.weight": "norm2_g", "norm2.bias": "norm2_b"}


def _backprop(out, dy, x, P, fields, grads):
    names = list(P)
    gs = torch.autograd.grad(out, [x] + [P[k] for k in names], dy, allow_unused=True)
    for k, g in zip(names, gs[1:]):
        _acc(getattr(grads, fields[k]), g, *P[k].shape)
    return gs[0]


def encoder_layer_bwd(desc, params, x_in, saved, dx_out, dx_in, grads, ws, ws_bytes, st):
    d, p, g = _st(desc, L.LayerDesc), _st(params, L.LayerParams), _st(grads, L.LayerGrads)
    P = {k: _leaf(v) for k, v in _layer_P(p, d.H, d.FF).items()}
    x = _leaf(_f32(x_in, d.B, d.T, d.H))
    dy = _f32(dx_out, d.B, d.T, d.H).clone()
    dx = _backprop(O.encoder_layer(x, P, "", d.heads), dy, x, P, _LAYER_FIELDS, g)
    _f32(dx_in, d.B, d.T, d.H).copy_(dx)


_VIT_FIELDS = {"layers.0.0.norm.weight": "norm_a_g", "layers.0.0.norm.bias": "norm_a_b", "layers.0.0.to_qkv.weight": "qkv_w",
               "layers.0.0.to_out.weight": "out_w", "layers.0.1.net.0.weight": "norm_f_g", "layers.0.1.net.0.bias": "norm_f_b",
               "layers.0.1.net.1.weight": "ff1_w", "layers.0.1.net.1.bias": "ff1_b", "layers.0.1.net.3.weight": "ff2_w",
               "layers.0.1.net.3.bias": "ff2_b"}


def _vit_P(p, D, inner, mlp):
    return {"layers.0.0.norm.weight": _f32(p.norm_a_g, D), "layers.0.0.norm.bias": _f32(p.norm_a_b, D),
            "layers.0.0.to_qkv.weight": _f32(p.qkv_w, 3 * inner, D), "layers.0.0.to_out.weight": _f32(p.out_w, D, inner),
            "layers.0.1.net.0.weight": _f32(p.norm_f_g, D), "layers.0.1.net.0.bias": _f32(p.norm_f_b, D),
            "layers.0.1.net.1.weight": _f32(p.ff1_w, mlp, D), "layers.0.1.net.1.bias": _f32(p.ff1_b, mlp),
            "layers.0.1.net.3.weight": _f32(p.ff2_w, D, mlp), "layers.0.1.net.3.bias": _f32(p.ff2_b, D)}


def vit_layer_bwd(desc, params, x_in, saved, dx_out, dx_in, grads, ws, ws_bytes, st):
    d, p, g = _st(desc, L.VitDesc), _st(params, L.VitParams), _st(grads, L.VitGrads)
    P = {k: _leaf(v) for k, v in _vit_P(p, d.D, d.heads * d.dim_head, d.mlp).items()}
    x = _leaf(_f32(x_in, d.B, d.T, d.D))
    dy = _f32(dx_out, d.B, d.T, d.D).clone()
    dx = _backprop(O.simple_vit_transformer(x, P, "", d.heads), dy, x, P, _VIT_FIELDS, g)
    _f32(dx_in, d.B, d.T, d.D).copy_(dx)


def _dec_P(p, H, FF):
    P = {}
    for att, (iw, ib, ow, ob) in (("self_attn", (p.sa_in_w, p.sa_in_b, p.sa_out_w, p.sa_out_b)),
                                  ("multihead_attn", (p.ca_in_w, p.ca_in_b, p.ca_out_w, p.ca_out_b))):
        P[att + ".in_proj_weight"], P[att + ".in_proj_bias"] = _f32(iw, 3 * H, H), _f32(ib, 3 * H)
        P[att + ".out_proj.weight"], P[att + ".out_proj.bias"] = _f32(ow, H, H), _f32(ob, H)
    P["linear1.weight"], P["linear1.bias"] = _f32(p.lin1_w, FF, H), _f32(p.lin1_b, FF)
    P["linear2.weight"], P["linear2.bias"] = _f32(p.lin2_w, H, FF), _f32(p.lin2_b, H)
    for i, (g, b) in enumerate(((p.norm1_g, p.norm1_b), (p.norm2_g, p.norm2_b), (p.norm3_g, p.norm3_b)), 1):
        P[f"norm{i}.weight"], P[f"norm{i}.bias"] = _f32(g, H), _f32(b, H)
    return P


_DEC_FIELDS = {"self_attn.in_proj_weight": "sa_in_w", "self_attn.in_proj_bias": "sa_in_b", "self_attn.out_proj.weight": "sa_out_w",
               "self_attn.out_proj.bias": "sa_out_b", "multihead_attn.in_proj_weight": "ca_in_w",
               "multihead_attn.in_proj_bias": "ca_in_b", "multihead_attn.out_proj.weight": "ca_out_w",
               "multihead_attn.out_proj.bias": "ca_out_b", "linear1.weight": "lin1_w", "linear1.bias": "lin1_b",
               "linear2.weight": "lin2_w", "linear2.bias": "lin2_b", "norm1.weight": "norm1_g", "norm1.bias": "norm1_b",
               "norm2.weight": "norm2_g", "norm2.bias": "norm2_b", "norm3.weight": "norm3_g", "norm3.bias": "norm3_b"}


def _mem_index(d):
    n, j = torch.arange(d.rows), torch.arange(d.M)
    return (n // d.kv_inner)[:, None] * d.kv_outer + j[None, :] * d.kv_jstride + (n % d.kv_inner)[:, None] * d.kv_istride


def decoder_layer_bwd(desc, params, y_in, mem, saved, dy_out, dy_in, dmem, grads, ws, ws_bytes, st):
    d, p, g = _st(desc, L.DecoderDesc), _st(params, L.DecoderParams), _st(grads, L.DecoderGrads)
    H = d.H
    P = {k: _leaf(v) for k, v in _dec_P(p, H, d.FF).items()}
    y = _leaf(_f32(y_in, d.rows, d.S, H))
    memory = _leaf(_f32(mem, d.mem_rows, H))
    mask = torch.full((d.S, d.S), float("-inf")).triu(1)
    out = O.decoder_layer(y, memory[_mem_index(d)], P, "", d.heads, mask, 0.0, False)
    dy = _f32(dy_out, d.rows, d.S, H).clone()
    names = list(P)
    gs = torch.autograd.grad(out, [y, memory] + [P[k] for k in names], dy, allow_unused=True)
    for k, gk in zip(names, gs[2:]):
        _acc(getattr(g, _DEC_FIELDS[k]), gk, *P[k].shape)
    _acc(dmem, gs[1], d.mem_rows, H)                                     # accumulated over the decoder layers
    _f32(dy_in, d.rows, d.S, H).copy_(gs[0])


def prompt_embed_bwd(dtype, rows, S, H, tokens, dy, p_drop, training, seed, d_embedding, st):
    tok = _i64(tokens, rows, S)
    demb = _f32(d_embedding, int(tok.max()) + 1, H)
    demb.index_put_((tok.reshape(-1),), _f32(dy, rows * S, H) * math.sqrt(H), accumulate=True)


def pool_bwd(dtype, B, T, H, pool, row_tokens, dpooled, dx, st):
    out = _f32(dx, B, T, H)
    if pool:
        out.copy_(_f32(dpooled, B, 1, H).expand(B, T, H) / T)
    else:
        out.zero_()
        out[:, :row_tokens] = _f32(dpooled, B, row_tokens, H)


def _head_forward(d, x, lng, lnb, w, b):
    g = x.mean(dim=1) if d.pool else x[:, :d.row_tokens].reshape(-1, d.H)
    if d.use_ln:
        g = O.layer_norm(g, lng, lnb, d.ln_eps)
    return O.linear(g, w, b)


def head_loss_bwd(desc, hin, hsaved, dlogits, dloss_scale, dx, grads, ws, ws_bytes, st):
    d, i, g = _st(desc, L.HeadDesc), _st(hin, L.HeadIn), _st(grads, L.HeadGrads)
    rows = d.B if d.pool else d.B * d.row_tokens
    x = _leaf(_f32(i.x, d.B, d.T, d.H))
    w, b = _leaf(_f32(i.w, d.n_out, d.H)), _leaf(_f32(i.b, d.n_out))
    lng = _leaf(_f32(i.ln_g, d.H)) if d.use_ln else None
    lnb = _leaf(_f32(i.ln_b, d.H)) if d.use_ln else None
    logits = _head_forward(d, x, lng, lnb, w, b)
    if d.loss == L.LOSS_NONE:
        target, seed_grad = logits, _f32(dlogits, rows, d.n_out).clone()
    else:
        if d.loss == L.LOSS_CE:
            cw = _f32(i.class_weight, d.n_out) if i.class_weight else None
            loss = O.ce_loss(logits, _i64(i.labels, rows), cw)
        elif d.loss == L.LOSS_BCE_SIGMOID:
            loss = O.bce_sigmoid_loss(logits, torch.nn.functional.one_hot(_i64(i.labels, rows), d.n_out).float())
        else:
            groups = [d.group_size[k] for k in range(d.n_groups)]
            loss = O.lta_loss(logits.view(rows, d.sub_rows, -1), _i64(i.labels, rows, d.sub_rows, d.n_groups), groups)
        target, seed_grad = loss, torch.tensor(float(dloss_scale))
    leaves = [x, w, b] + ([lng, lnb] if d.use_ln else [])
    gs = torch.autograd.grad(target, leaves, seed_grad)
    _f32(dx, d.B, d.T, d.H).copy_(gs[0])
    _acc(g.w, gs[1], d.n_out, d.H)
    _acc(g.b, gs[2], d.n_out)
    if d.use_ln:
        _acc(g.ln_g, gs[3], d.H)
        _acc(g.ln_b, gs[4], d.H)


def embed_bwd(desc, ein, esaved, dx, grads, ws, ws_bytes, st):
    d, i, g = _st(desc, L.EmbedDesc), _st(ein, L.EmbedIn), _st(grads, L.EmbedGrads)
    B, T, H = d.B, d.T, d.H
    feats = [_leaf(_f32(i.feat[k], B, d.seg_tokens[k], d.seg_in_dim[k])) for k in range(d.n_seg)]
    ws_ = [_leaf(_f32(i.proj_w[k], H, d.seg_in_dim[k])) if d.seg_has_proj[k] else None for k in range(d.n_seg)]
    bs_ = [_leaf(_f32(i.proj_b[k], H)) if d.seg_has_proj[k] else None for k in range(d.n_seg)]
    lng = None if d.no_ln else _leaf(_f32(i.ln_g, H))
    lnb = None if d.no_ln else _leaf(_f32(i.ln_b, H))
    table = _leaf(_f32(i.tok_table, T, H))
    z = torch.cat([O.linear(f, w, b) if w is not None else f for f, w, b in zip(feats, ws_, bs_)], dim=1)
    x = z + table if d.no_ln else O.layer_norm(z, lng, lnb, d.ln_eps) + table
    leaves = [t for t in feats + ws_ + bs_ + [lng, lnb, table] if t is not None]
    gs = dict(zip(map(id, leaves), torch.autograd.grad(x, leaves, _f32(dx, B, T, H).clone(), allow_unused=True)))
    for k in range(d.n_seg):
        if ws_[k] is not None:
            _acc(g.proj_w[k], gs[id(ws_[k])], H, d.seg_in_dim[k])
            _acc(g.proj_b[k], gs[id(bs_[k])], H)
        _acc(g.dfeat[k], gs[id(feats[k])], B, d.seg_tokens[k], d.seg_in_dim[k])
        if g.seg_embed[k]:          # column sums of the table gradient over segment k's tokens (+=: segments may share a row)
            off = d.seg_offset[k]
            _acc(g.seg_embed[k], gs[id(table)][off:off + d.seg_tokens[k]].sum(dim=0), H)
    if not d.no_ln:
        _acc(g.ln_g, gs[id(lng)], H)
        _acc(g.ln_b, gs[id(lnb)], H)
    _acc(g.tok_table, gs[id(table)], T, H)


BACKWARD = {"egot2_encoder_layer_bwd": encoder_layer_bwd, "egot2_vit_layer_bwd": vit_layer_bwd,
            "egot2_decoder_layer_bwd": decoder_layer_bwd, "egot2_prompt_embed_bwd": prompt_embed_bwd,
            "egot2_pool_bwd": pool_bwd, "egot2_head_loss_bwd": head_loss_bwd, "egot2_embed_bwd": embed_bwd}

FORWARD = {"egot2_hhi_tok_table_fwd": tok_table_fwd, "egot2_embed_fwd": embed_fwd, "egot2_encoder_layer_fwd": encoder_layer_fwd,
           "egot2_vit_layer_fwd": vit_layer_fwd, "egot2_prompt_embed_fwd": prompt_embed_fwd,
           "egot2_decoder_layer_fwd": decoder_layer_fwd, "egot2_pool_fwd": pool_fwd, "egot2_head_loss_fwd": head_loss_fwd,
           "egot2_row_softmax": row_softmax}


def install(monkeypatch):
    """Route the product code's C-ABI calls to the emulator; returns the list of (name, args) it saw."""
    calls = []

    def call(name, *args):
        calls.append((name, args))
        fn = FORWARD.get(name)
        if fn is not None:
            with torch.no_grad():
                fn(*args)
        elif name in BACKWARD:
            with torch.enable_grad():
                BACKWARD[name](*args)
        elif name.endswith(("_fwd", "_bwd")):
            raise NotImplementedError(f"abi_emulator: {name}")
        return 0
    monkeypatch.setattr(L, "call", call)
    monkeypatch.setattr(E, "_stream", lambda: 0)
    monkeypatch.setattr(M, "_require_cuda", lambda device: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    return calls

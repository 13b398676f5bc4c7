"""CPU checks of the ALGORITHMS of kernels / launch sequences that have not run on hardware yet: each test re-states the
device code step by step (same loop structure, index expressions and launcher argument conventions as the .cu file, in
float64 torch) and compares it with autograd of the plain formula.  They guard the math and the indexing of
csrc/attention_wide.cu and of the launch sequence in csrc/vit.cu; the CUDA specifics (warp reductions, shared memory,
dtypes) are only exercised by the `-m gpu` tests."""
import math

import torch

from oracle import translator_oracle as O

DT = torch.float64


def test_attention_wide_kernel_indexing():
    """attn_wide_fwd_kernel / attn_wide_bwd_kernel: scores P[i][j], lse, Pd / dS tiles and the three gradient sums."""
    torch.manual_seed(0)
    B, T, heads, dh = 2, 4, 3, 160
    H = heads * dh
    qkv = torch.randn(B, T, 3 * H, dtype=DT, requires_grad=True)
    dout = torch.randn(B, T, H, dtype=DT)
    q, k, v = qkv.split(H, dim=-1)
    sh = lambda t: t.reshape(B, T, heads, dh).transpose(1, 2)
    p = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(dh), -1)
    o = (p @ sh(v)).transpose(1, 2).reshape(B, T, H)
    (gref,) = torch.autograd.grad(o, qkv, dout)

    Q = qkv.detach()
    out = torch.zeros(B, T, H, dtype=DT)
    lse = torch.zeros(B * heads, T, dtype=DT)
    scale = 1 / math.sqrt(dh)
    for bh in range(B * heads):                                  # one CTA per (clip, head)
        b, h = bh // heads, bh % heads
        qs, ks, vs = (Q[b, :, o_ + h * dh:o_ + (h + 1) * dh] for o_ in (0, H, 2 * H))
        P = torch.zeros(T, T, dtype=DT)
        for pair in range(T * T):                                # one warp per (query, key) pair
            i, j = pair // T, pair % T
            P[i, j] = (qs[i] * ks[j]).sum() * scale
        for i in range(T):                                       # one warp per query, lane = key
            mx = P[i].max()
            e = torch.exp(P[i] - mx)
            lse[bh, i] = mx + torch.log(e.sum())
            P[i] = e / e.sum()
        for i in range(T):
            out[b, i, h * dh:(h + 1) * dh] = sum(P[i, j] * vs[j] for j in range(T))
    assert float((out - o.detach()).abs().max()) < 1e-12

    dqkv = torch.zeros_like(Q)
    for bh in range(B * heads):
        b, h = bh // heads, bh % heads
        qs, ks, vs = (Q[b, :, o_ + h * dh:o_ + (h + 1) * dh] for o_ in (0, H, 2 * H))
        dO, Oh = dout[b, :, h * dh:(h + 1) * dh], out[b, :, h * dh:(h + 1) * dh]
        Dv = (dO * Oh).sum(-1)
        Pd, dS = torch.zeros(T, T, dtype=DT), torch.zeros(T, T, dtype=DT)
        for pair in range(T * T):
            i, j = pair // T, pair % T
            pp = torch.exp((qs[i] * ks[j]).sum() * scale - lse[bh, i])
            dp = (dO[i] * vs[j]).sum()
            Pd[i, j] = pp                                        # mask multiplier 1 (eval)
            dS[i, j] = pp * (dp - Dv[i])
        for r in range(T):
            dq = sum(dS[r, t] * ks[t] for t in range(T))
            dk = sum(dS[t, r] * qs[t] for t in range(T))
            dv = sum(Pd[t, r] * dO[t] for t in range(T))
            dqkv[b, r, h * dh:(h + 1) * dh] = dq * scale
            dqkv[b, r, H + h * dh:H + (h + 1) * dh] = dk * scale
            dqkv[b, r, 2 * H + h * dh:2 * H + (h + 1) * dh] = dv
    assert float((dqkv - gref).abs().max()) < 1e-12


def test_vit_layer_launch_sequence():
    """egot2_vit_layer_fwd / _bwd: the order and the operands of the lin / dgrad / wgrad2 / LayerNorm / attention / GELU
    launches (ops.h conventions: lin = A W^T + bias + residual, dgrad = dY W, wgrad2 = dY^T X, LayerNorm backward adds
    `dres`)."""
    torch.manual_seed(1)
    B, T, D, heads, dhd, mlp = 2, 5, 32, 2, 24, 48
    inner, M = heads * dhd, B * T
    r = lambda *s: torch.randn(*s, dtype=DT)
    P = {"layers.0.0.norm.weight": 1 + 0.1 * r(D), "layers.0.0.norm.bias": 0.1 * r(D),
         "layers.0.0.to_qkv.weight": r(3 * inner, D) / D ** .5, "layers.0.0.to_out.weight": r(D, inner) / inner ** .5,
         "layers.0.1.net.0.weight": 1 + 0.1 * r(D), "layers.0.1.net.0.bias": 0.1 * r(D),
         "layers.0.1.net.1.weight": r(mlp, D) / D ** .5, "layers.0.1.net.1.bias": 0.1 * r(mlp),
         "layers.0.1.net.3.weight": r(D, mlp) / mlp ** .5, "layers.0.1.net.3.bias": 0.1 * r(D)}
    Pl = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    x = r(B, T, D).requires_grad_(True)
    y = O.simple_vit_transformer(x, Pl, "", heads)
    dy = torch.randn_like(y)
    names = list(Pl)
    ref = dict(zip(["x"] + names, torch.autograd.grad(y, [x] + [Pl[k] for k in names], dy)))

    def lin(A, W, bias=None, res=None):
        C = A @ W.t()
        C = C if bias is None else C + bias
        return C if res is None else C + res

    dgrad = lambda dY, W: dY @ W
    wgrad2 = lambda dY, X: dY.t() @ X

    def ln_fwd(xx, g, b, eps=1e-5):
        mean = xx.mean(-1, keepdim=True)
        rstd = 1 / torch.sqrt(((xx - mean) ** 2).mean(-1, keepdim=True) + eps)
        return (xx - mean) * rstd * g + b, (mean, rstd)

    def ln_bwd(xx, stat, g, dyy, dres):
        mean, rstd = stat
        xh, dyg = (xx - mean) * rstd, dyy * g
        dx = rstd * (dyg - dyg.mean(-1, keepdim=True) - xh * (dyg * xh).mean(-1, keepdim=True)) + dres
        return dx, (dyy * xh).sum(0), dyy.sum(0)

    def attn_fwd(qkv):
        q, k, v = qkv.split(inner, -1)
        sh = lambda t: t.reshape(B, T, heads, dhd).transpose(1, 2)
        p = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(dhd), -1)
        return (p @ sh(v)).transpose(1, 2).reshape(B, T, inner)

    def attn_bwd(qkv, dout):
        q = qkv.clone().requires_grad_(True)
        return torch.autograd.grad(attn_fwd(q), q, dout)[0]

    gelu = lambda u: 0.5 * u * (1 + torch.erf(u * 0.70710678118654752))                       # gelu_f
    gelu_grad = lambda u: 0.5 * (1 + torch.erf(u * 0.70710678118654752)) + u * 0.39894228040143268 * torch.exp(-0.5 * u * u)
    p = {"norm_a_g": "0.norm.weight", "norm_a_b": "0.norm.bias", "qkv_w": "0.to_qkv.weight", "out_w": "0.to_out.weight",
         "norm_f_g": "1.net.0.weight", "norm_f_b": "1.net.0.bias", "ff1_w": "1.net.1.weight", "ff1_b": "1.net.1.bias",
         "ff2_w": "1.net.3.weight", "ff2_b": "1.net.3.bias"}                                  # engine._VIT_NAMES
    w = {f: P["layers.0." + n] for f, n in p.items()}
    xin = x.detach().reshape(M, D)
    # ---- egot2_vit_layer_fwd
    h, stat_a = ln_fwd(xin, w["norm_a_g"], w["norm_a_b"])
    qkv = lin(h, w["qkv_w"])
    attn = attn_fwd(qkv.reshape(B, T, 3 * inner)).reshape(M, inner)
    x1 = lin(attn, w["out_w"], None, xin)
    h2, stat_f = ln_fwd(x1, w["norm_f_g"], w["norm_f_b"])
    u = lin(h2, w["ff1_w"], w["ff1_b"])
    act = gelu(u)
    xout = lin(act, w["ff2_w"], w["ff2_b"], x1)
    assert float((xout.reshape(B, T, D) - y.detach()).abs().max()) < 1e-12
    # ---- egot2_vit_layer_bwd
    dxo, g = dy.reshape(M, D), {}
    g["ff2_w"], g["ff2_b"] = wgrad2(dxo, act), dxo.sum(0)
    du = dgrad(dxo, w["ff2_w"]) * gelu_grad(u)
    g["ff1_w"], g["ff1_b"] = wgrad2(du, h2), du.sum(0)
    dh = dgrad(du, w["ff1_w"])
    d1, g["norm_f_g"], g["norm_f_b"] = ln_bwd(x1, stat_f, w["norm_f_g"], dh, dxo)
    g["out_w"] = wgrad2(d1, attn)
    dout = dgrad(d1, w["out_w"])
    dqkv = attn_bwd(qkv.reshape(B, T, 3 * inner), dout.reshape(B, T, inner)).reshape(M, 3 * inner)
    g["qkv_w"] = wgrad2(dqkv, h)
    dh = dgrad(dqkv, w["qkv_w"])
    dxin, g["norm_a_g"], g["norm_a_b"] = ln_bwd(xin, stat_a, w["norm_a_g"], dh, d1)
    for f, n in p.items():
        assert float((g[f] - ref["layers.0." + n]).abs().max()) < 1e-11, f
    assert float((dxin.reshape(B, T, D) - ref["x"]).abs().max()) < 1e-11


def test_adamw_update_rule():
    """adam_kernel<DECOUPLED = true>: w *= 1 - lr*wd, then the Adam step with denom = sqrt(v)/sqrt(bc2) + eps."""
    torch.manual_seed(2)
    n, lr, b1, b2, eps, wd = 257, 1e-2, 0.9, 0.999, 1e-8, 0.1
    p = torch.randn(n, dtype=DT)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    m, v = torch.zeros(n, dtype=DT), torch.zeros(n, dtype=DT)
    for step in range(1, 5):
        grad = torch.randn(n, dtype=DT)
        ref.grad = grad.clone()
        opt.step()
        bc1, bc2_sqrt = 1 - b1 ** step, math.sqrt(1 - b2 ** step)
        w = p * (1 - lr * wd)
        m = b1 * m + (1 - b1) * grad
        v = b2 * v + (1 - b2) * grad * grad
        p = w - (lr / bc1) * m / (torch.sqrt(v) / bc2_sqrt + eps)
        assert float((p - ref.detach()).abs().max()) < 1e-12

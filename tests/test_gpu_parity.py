"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed goldens
of the real reference.  fp32 mode: 1e-3 of the output range (we assert far tighter) and
argmax bit-exact; bf16 mode: 2e-2 (north_star tolerances)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from egot2_b200 import _lib as L
from oracle.cases import CASES, UNVALIDATED_ON_GPU, case_inputs, grad_digest, oracle_forward_loss

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# Output/loss tolerances are relative to the reference tensor's absmax (north_star: 1e-3 fp32, 2e-2 bf16; we
# assert tighter in fp32).  Gradients, fp32: max-abs error relative to absmax.  Gradients, bf16: the relative L2 error
# of every tensor is bounded by max(grad_l2, BF16_COND x what stock torch bf16 arithmetic loses on that very tensor)
# (oracle.cases.bf16_conditioning: 1-6e-2 for most, 0.2 for the one ill-conditioned decoder self-attention of
# hoi_g_h128_l2), and element-wise 99.9 % of a tensor within `grad` of the reference absmax (single rows legitimately
# move more when a ReLU gate whose pre-activation is within bf16 rounding of zero flips: a property of bf16
# arithmetic, not of the kernels).
TOL = {"fp32": dict(out=2e-4, loss=2e-4, grad=2e-3, grad_l2=2e-3), "bf16": dict(out=2e-2, loss=2e-2, grad=0.1, grad_l2=4e-2)}
BF16_COND = 2.5      # our path also STORES every activation in bf16 between kernels (autocast keeps LayerNorm / softmax outputs in fp32)


def bf16_grad_bound(case, name, tol):
    from oracle.cases import bf16_conditioning
    return max(tol["grad_l2"], BF16_COND * bf16_conditioning(case).get(name, 0.0))


def _loss_kind(case):
    sp = case.spec
    if sp.family == "hhi_ttm":
        return L.LOSS_CE, torch.tensor([0.266, 0.734])
    if sp.family == "hoi_pnr":
        return (L.LOSS_BCE_SIGMOID if sp.n_out == 16 else L.LOSS_CE), None
    if sp.family in ("hoi_lta", "hoi_ar"):
        return L.LOSS_CE_GROUPS, None
    if sp.family in ("hhi_g", "hoi_g"):
        return L.LOSS_CE, None
    return L.LOSS_NONE, None


def _engine_feats(case, eng, feats, extra, dtype):
    from egot2_b200 import engine as E
    dev = eng.device
    out = []
    for s in case.spec.segments:
        if case.raw_slowfast and s.name in ("slow", "fast"):
            raw = extra["slow5" if s.name == "slow" else "fast5"].to(dev)
            B, Cc, Tin, h, w = raw.shape
            o = torch.empty((B, 8, Cc), device=dev, dtype=torch.float32)
            L.call("egot2_slowfast_pool_fwd", raw.data_ptr(), L.F32, B, Cc, Tin, h * w, 8, o.data_ptr(), L.F32,
                   E._stream())
            out.append(o if dtype == "fp32" else o.to(torch.bfloat16))
        else:
            f = feats[s.name].to(dev)
            out.append(f if dtype == "fp32" else f.to(torch.bfloat16))
    return out


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n not in UNVALIDATED_ON_GPU])
def test_engine_matches_oracle_and_golden(name, dtype):
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    case = CASES[name]
    sp = case.spec
    tol = TOL[dtype]
    sd, feats, labels, extra = case_inputs(case)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))

    eng = TranslatorEngine(sp, "cuda:0", dtype)
    eng.arena.load_state_dict(sd)
    if sp.embed == "task_sinusoid":
        eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
    loss_kind, cw = _loss_kind(case)
    gfeats = _engine_feats(case, eng, feats, extra, dtype)
    if sp.family in ("hhi_g", "hoi_g"):      # decoder reads target[:, :-1], CE on target[:, 1:] (video_tasktranslation.py:48-61)
        act = eng.forward(gfeats, training=False, labels=labels[:, 1:], loss=loss_kind, prompt=labels[:, :-1])
        out = act.t["out"].float().cpu().view(labels.shape[0], 2, -1).permute(0, 2, 1)      # (rows, V, S) like the reference
    else:
        act = eng.forward(gfeats, training=False, labels=labels, loss=loss_kind, class_weight=cw)
        out = act.t["out"].float().cpu()

    # oracle (CPU, fp32) on the same inputs
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_out, o_loss = oracle_forward_loss(case, P, feats, labels, extra)
    ref_out = torch.from_numpy(gold["output"]).reshape(o_out.shape)
    scale = float(ref_out.abs().max())
    assert float((out.reshape(o_out.shape) - o_out.detach()).abs().max()) <= tol["out"] * scale, "vs oracle"
    assert float((out.reshape(o_out.shape) - ref_out).abs().max()) <= tol["out"] * scale, "vs reference golden"
    if dtype == "fp32" and o_out.dim() == 2 and o_out.shape[1] in (2, 16):
        assert torch.equal(out.argmax(-1), ref_out.argmax(-1)), "argmax / keyframe index must be bit-exact in fp32"
        if loss_kind != L.LOSS_NONE:
            assert torch.equal(act.t["argmax"].cpu().long(), ref_out.argmax(-1))

    names = [k[len("grad/"):] for k in gold.files if k.startswith("grad/")]
    if loss_kind != L.LOSS_NONE:
        loss = float(act.t["loss"][0].cpu())
        assert abs(loss - float(gold["loss"])) <= tol["loss"] * abs(float(gold["loss"])) + 1e-6
        grad, _ = eng.backward(act)
    else:  # ASD: the loss (lossAV) lives outside the translator; feed d(loss)/d(out) computed by the oracle
        o2 = out.clone().requires_grad_(True)
        l2 = O.loss_av(extra, o2, labels)[0]
        (dout,) = torch.autograd.grad(l2, o2)
        grad, _ = eng.backward(act, dout=dout.cuda())
    o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
    for k, g_ref in zip(names, o_grads):
        g = eng.arena.view(k, grad).float().cpu()
        if g_ref is None:
            assert float(g.abs().max()) == 0.0, k
            continue
        gscale = float(g_ref.abs().max()) + 1e-12
        diff = (g - g_ref).abs().flatten() / gscale
        err = float(diff.max())
        err_l2 = float((g - g_ref).norm()) / (float(g_ref.norm()) + 1e-12)
        if dtype == "fp32":
            # A ReLU gate whose pre-activation lies within fp32 rounding of zero (|z| ~ 1e-7; with FF = 2048 x a few
            # hundred tokens x L layers every seed has one - measured in float64) flips between two correct fp32
            # evaluation orders and moves ONE token's contribution to one hidden unit: an isolated outlier, not a
            # kernel error.  So: 99.9 % of the elements within the tight bound, every element within 10x of it.
            q = float(torch.quantile(diff[:: max(1, diff.numel() // 1000000)], 0.999)) if diff.numel() > 1 else err
            assert q <= tol["grad"], f"{k}: 99.9th percentile rel err {q:.3e}"
            assert err <= 10 * tol["grad"], f"{k}: max-abs rel err {err:.3e}"
            assert err_l2 <= 3 * tol["grad_l2"], f"{k}: rel L2 err {err_l2:.3e}"
        else:
            q = float(torch.quantile(diff[:: max(1, diff.numel() // 1000000)], 0.999)) if diff.numel() > 1 else err
            bound = bf16_grad_bound(case, k, tol)
            assert err_l2 <= bound, f"{k}: rel L2 err {err_l2:.3e} > {bound:.3e}"
            assert q <= max(tol["grad"], 3.0 * bound), f"{k}: 99.9th percentile rel err {q:.3e}"
        dg = grad_digest(g)
        ref_d = torch.from_numpy(gold["grad/" + k])
        assert abs(float(dg[1] - ref_d[1])) <= 2 * tol["grad"] * float(ref_d[1]) + 1e-7, f"{k}: l2 norm vs golden"


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_gemm_orientations(dtype, ta, tb):
    from egot2_b200 import engine as E
    torch.manual_seed(0)
    M, N, K = 200, 136, 264
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(tdt)
    B = torch.randn(K, N, device="cuda").to(tdt)
    bias = torch.randn(N, device="cuda")
    As = A.t().contiguous() if ta else A.contiguous()
    Bs = B.t().contiguous() if tb else B.contiguous()
    Cout = torch.empty(M, N, device="cuda", dtype=torch.float32)
    L.call("egot2_gemm", E._dt(dtype), M, N, K, As.data_ptr(), ta, Bs.data_ptr(), tb, bias.data_ptr(), 1,
           Cout.data_ptr(), 1, 0, E._stream())
    ref = torch.relu(A.double() @ B.double() + bias.double()).float()
    tol = 1e-5 if dtype == "fp32" else 1e-5   # bf16 inputs are exact in fp32; accumulation is fp32 either way
    assert float((Cout - ref).abs().max()) <= tol * float(ref.abs().max()) + 1e-4


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("T,H,heads", [(7, 128, 4), (90, 128, 4), (48, 128, 8), (150, 256, 4), (8, 1024, 8), (453, 64, 2)])
def test_attention_fwd_bwd(dtype, T, H, heads):
    from egot2_b200 import engine as E
    torch.manual_seed(1)
    B = 3
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    qkv = (torch.randn(B, T, 3 * H, device="cuda") * 0.7).to(tdt)
    dout = torch.randn(B, T, H, device="cuda").to(tdt)
    out = torch.empty(B, T, H, device="cuda", dtype=tdt)
    lse = torch.empty(B, heads, T, device="cuda", dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    L.call("egot2_attention_fwd", E._dt(dtype), B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), 0.0, 0, 0,
           E._stream())
    nws = L.load().egot2_attention_bwd_workspace_bytes(E._dt(dtype), B, T, H, heads)
    ws = torch.empty(nws, device="cuda", dtype=torch.uint8)
    L.call("egot2_attention_bwd", E._dt(dtype), B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
           dout.data_ptr(), dqkv.data_ptr(), 0.0, 0, 0, ws.data_ptr(), nws, E._stream())
    q = qkv.double().requires_grad_(True)
    dh = H // heads
    qq, kk, vv = q.split(H, dim=-1)
    sh = lambda x: x.reshape(B, T, heads, dh).transpose(1, 2)
    s = (sh(qq) / dh ** 0.5) @ sh(kk).transpose(-1, -2)
    ref = (torch.softmax(s, -1) @ sh(vv)).transpose(1, 2).reshape(B, T, H)
    (gref,) = torch.autograd.grad(ref, q, dout.double())
    t_out, t_g = (1e-4, 1e-3) if dtype == "fp32" else (1e-2, 2e-2)
    assert float((out.double() - ref).abs().max()) <= t_out * float(ref.abs().max())
    assert float((lse.double() - torch.logsumexp(s, -1)).abs().max()) <= 1e-3
    assert float((dqkv.double() - gref).abs().max()) <= t_g * float(gref.abs().max())


def test_adam_matches_torch():
    from egot2_b200 import engine as E
    torch.manual_seed(2)
    n = 10007
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, weight_decay=0.01)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        L.call("egot2_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 5e-4, 0.9, 0.999, 1e-8,
               0.01, step, 1.0, E._stream())
    assert float((p - ref.detach()).abs().max()) < 1e-6


@pytest.mark.parametrize("M,N,K,ta,tb,acc", [
    (23040, 2048, 128, 0, 1, 0),     # ffn.linear1 forward
    (23040, 128, 2048, 0, 1, 0),     # ffn.linear2 forward (long K, 4-stage ring wraps)
    (23040, 384, 128, 0, 1, 0),      # packed qkv projection
    (23040, 2048, 128, 0, 0, 0),     # dhid = dy . W2      (B is MN-major)
    (2048, 128, 23040, 1, 0, 1),     # dW1 = dhid^T . x1   (A and B MN-major, split-K atomics)
    (128, 2048, 23040, 1, 0, 1),     # dW2 = dy^T . hid
    (304, 136, 264, 0, 1, 0), (304, 136, 264, 0, 0, 0), (304, 136, 264, 1, 0, 1), (304, 136, 264, 1, 1, 0),
])
def test_tcgen05_gemm(M, N, K, ta, tb, acc):
    """The bf16 GEMMs of the step must be served by the tcgen05/TMEM/TMA kernel and match an fp64 reference."""
    from egot2_b200 import engine as E
    torch.manual_seed(3)
    A = (torch.randn(M, K, device="cuda") / K ** 0.25).to(torch.bfloat16)
    B = (torch.randn(K, N, device="cuda") / K ** 0.25).to(torch.bfloat16)
    As = A.t().contiguous() if ta else A.contiguous()
    Bs = B.t().contiguous() if tb else B.contiguous()
    bias = None if acc else torch.randn(N, device="cuda")
    C0 = torch.randn(M, N, device="cuda") if acc else torch.zeros(M, N, device="cuda")
    Cout = C0.clone()
    L.call("egot2_gemm", L.BF16, M, N, K, As.data_ptr(), ta, Bs.data_ptr(), tb, None if acc else bias.data_ptr(), 0,
           Cout.data_ptr(), 1, acc, E._stream())
    torch.cuda.synchronize()
    assert L.load().egot2_gemm_last_impl() == b"tcgen05"
    ref = A.double() @ B.double() + (C0.double() if acc else bias.double())
    err = float((Cout.double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-5, err
    # bf16 output path
    if not acc:
        Cb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        L.call("egot2_gemm", L.BF16, M, N, K, As.data_ptr(), ta, Bs.data_ptr(), tb, bias.data_ptr(), 1, Cb.data_ptr(), 0, 0,
               E._stream())
        refb = torch.relu(A.double() @ B.double() + bias.double())
        assert float((Cb.double() - refb).abs().max()) / float(refb.abs().max()) < 1e-2


def _np_mix64(x):
    import numpy as np
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
    x ^= x >> np.uint64(33)
    return x


def dropout_mask_oracle(seed, site, layer, n, p, bit_mode=False, raw_bits=False):
    """Bit-exact restatement of csrc/common.cuh drop_scale(): multiplier (0 or 1/(1-p)) for element indices 0..n-1."""
    import numpy as np
    M32 = np.uint64(0xFFFFFFFF)

    def bits(key, idx):
        x = (idx & M32) ^ (key & M32) ^ (((idx >> np.uint64(32)) * np.uint64(0x9E3779B1)) & M32)
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(0x7feb352d)) & M32
        x ^= (x >> np.uint64(15)) ^ (key >> np.uint64(32))
        return (x * np.uint64(0x846ca68b)) & M32

    with np.errstate(over="ignore"):
        key = _np_mix64(np.array([seed], dtype=np.uint64) ^ (np.uint64(0x9E3779B97F4A7C15) * np.uint64(site + 16 * layer + 1)))[0]
        idx = np.arange(n, dtype=np.uint64)
        if raw_bits:
            return bits(key, idx)
        if p == 0.5:      # every elementwise site at p == 0.5: one random bit per element (bit idx%32 of the hash of idx/32)
            keep = ((bits(key, idx >> np.uint64(5)) >> (idx & np.uint64(31))) & np.uint64(1)) == 1
        else:             # other p: a 16-bit field per element (half idx % 2 of the hash of idx / 2) against round(p * 2^16)
            thr = np.uint64(min(65535, int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))))
            h = bits(key, idx >> np.uint64(1))
            keep = np.where((idx & np.uint64(1)) == 1, h >> np.uint64(16), h & np.uint64(0xffff)) >= thr
    return torch.from_numpy(np.where(keep, np.float32(1.0 / (1.0 - p)), np.float32(0.0)))


def attn_mask_oracle(seed, n_bh, T, p):
    """csrc/common.cuh attn_drop_keep() for every (b*heads+h, query, key): multiplier 0 or 1/(1-p), shape (n_bh*T*T,)."""
    import numpy as np
    if p != 0.5:          # 16-bit field key % 2 of the hash of (row * ceil(T/2) + key // 2)
        hp = (T + 1) // 2
        words = dropout_mask_oracle(seed, 3, 0, n_bh * T * hp, 0.25, raw_bits=True)
        rows = np.arange(n_bh * T, dtype=np.int64)[:, None]
        keys = np.arange(T, dtype=np.int64)[None, :]
        h = words[rows * hp + keys // 2]
        f = np.where(keys % 2 == 1, h >> np.uint64(16), h & np.uint64(0xffff))
        thr = np.uint64(min(65535, int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))))
        return torch.from_numpy(np.where(f >= thr, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).reshape(-1))
    wpr = (T + 31) // 32
    words = dropout_mask_oracle(seed, 3, 0, n_bh * T * wpr, 0.25, raw_bits=True)      # hash of row*wpr + key//32
    rows = np.arange(n_bh * T, dtype=np.int64)[:, None]
    keys = np.arange(T, dtype=np.int64)[None, :]
    w = words[rows * wpr + keys // 32]
    keep = (w >> (keys % 32).astype(np.uint64)) & np.uint64(1)
    return torch.from_numpy(np.where(keep == 1, np.float32(2.0), np.float32(0.0)).reshape(-1))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("T,H,heads", [(90, 128, 4), (48, 128, 8), (128, 256, 4), (33, 64, 4), (150, 128, 4), (8, 1024, 8), (200, 128, 4),
                                         (453, 64, 2), (300, 256, 4)])
def test_attention_dropout_mask_is_exact(dtype, T, H, heads):
    """Attention with dropout on the probabilities (train mode): forward and backward must use exactly the mask of
    the documented counter-based generator — checked against a torch reference fed the same mask."""
    from egot2_b200 import engine as E
    torch.manual_seed(4)
    B, seed = 2, 99
    p = 0.5 if T in (48, 128, 150, 453) else 0.25     # p == 0.5: one random bit per (query, key) pair, 32 keys per hash
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    qkv = (torch.randn(B, T, 3 * H, device="cuda") * 0.7).to(tdt)
    dout = torch.randn(B, T, H, device="cuda").to(tdt)
    out = torch.empty(B, T, H, device="cuda", dtype=tdt)
    lse = torch.empty(B, heads, T, device="cuda", dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    L.call("egot2_attention_fwd", E._dt(dtype), B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), p, 1, seed,
           E._stream())
    nws = L.load().egot2_attention_bwd_workspace_bytes(E._dt(dtype), B, T, H, heads)
    ws = torch.empty(max(nws, 16), device="cuda", dtype=torch.uint8)
    L.call("egot2_attention_bwd", E._dt(dtype), B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
           dout.data_ptr(), dqkv.data_ptr(), p, 1, seed, ws.data_ptr(), nws, E._stream())
    mask = attn_mask_oracle(seed, B * heads, T, p).reshape(B, heads, T, T).cuda().double()
    q = qkv.double().requires_grad_(True)
    dh = H // heads
    qq, kk, vv = q.split(H, dim=-1)
    sh = lambda x: x.reshape(B, T, heads, dh).transpose(1, 2)
    s = (sh(qq) / dh ** 0.5) @ sh(kk).transpose(-1, -2)
    ref = ((torch.softmax(s, -1) * mask) @ sh(vv)).transpose(1, 2).reshape(B, T, H)
    (gref,) = torch.autograd.grad(ref, q, dout.double())
    t_out, t_g = (1e-4, 1e-3) if dtype == "fp32" else (1.5e-2, 3e-2)
    assert float((out.double() - ref.detach()).abs().max()) <= t_out * float(ref.abs().max())
    assert float((dqkv.double() - gref).abs().max()) <= t_g * float(gref.abs().max())


@pytest.mark.parametrize("T,H,heads,want", [(90, 128, 4, "attn_mma_"), (8, 1024, 8, "attn_mma_"), (150, 256, 4, "attn_long_"),
                                            (453, 64, 2, "attn_long_"), (512, 128, 4, "attn_long_"), (384, 256, 4, "attn_long_")])
def test_bf16_attention_runs_on_tensor_cores(T, H, heads, want):
    """In bf16 mode no translator shape of the reference may fall back to the FFMA attention kernels: T <= 128 (and head dim
    128 at T <= 32) take attention_mma.cu, 128 < T <= 512 take attention_long.cu - checked on the launcher tags."""
    from egot2_b200 import engine as E
    B = 2
    qkv = (torch.randn(B, T, 3 * H, device="cuda") * 0.5).to(torch.bfloat16)
    dout = torch.randn(B, T, H, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, T, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, heads, T, device="cuda", dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    nws = L.load().egot2_attention_bwd_workspace_bytes(L.BF16, B, T, H, heads)
    ws = torch.empty(max(nws, 16), device="cuda", dtype=torch.uint8)
    L.prof_enable(True)
    try:
        L.call("egot2_attention_fwd", L.BF16, B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), 0.1, 1, 3, E._stream())
        L.call("egot2_attention_bwd", L.BF16, B, T, H, heads, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
               dqkv.data_ptr(), 0.1, 1, 3, ws.data_ptr(), nws, E._stream())
        torch.cuda.synchronize()
        tags = [r[0] for r in L.prof_report()]
    finally:
        L.prof_enable(False)
    assert any(t.startswith(want + "fwd") for t in tags) and any(t.startswith(want + "bwd") for t in tags), tags
    assert torch.isfinite(out.float()).all() and torch.isfinite(dqkv.float()).all()


@pytest.mark.parametrize("name", ["hhi3_h128_d30", "hoi_pnr_h128_l6"])
def test_fused_ffn_matches_unfused_with_dropout(name, monkeypatch):
    """The tcgen05 fused FFN (GEMM+ReLU+dropout+GEMM+residual+LN) must reproduce the unfused kernel sequence in
    TRAIN mode too: same dropout masks (counter-based), same bf16 rounding points; only the fp32 accumulation
    order differs."""
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    case = CASES[name]
    sp = case.spec
    sd, feats, labels, extra = case_inputs(case)
    outs = {}
    for mode in ("unfused", "fused"):
        monkeypatch.setenv("EGOT2_FFN", mode)
        eng = TranslatorEngine(sp, "cuda:0", "bf16")
        eng.arena.load_state_dict(sd)
        if sp.embed == "task_sinusoid":
            eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
        gf = [feats[s.name].cuda().bfloat16() for s in sp.segments]
        act = eng.forward(gf, training=True, seed=11)
        torch.cuda.synchronize()
        outs[mode] = {k: act.t[k].float().clone() for k in ("hid0", "y2_0", "x1", "out")}
    for k in outs["fused"]:
        a, b = outs["fused"][k], outs["unfused"][k]
        scale = float(b.abs().max())
        frac_zero_a, frac_zero_b = float((a == 0).float().mean()), float((b == 0).float().mean())
        assert abs(frac_zero_a - frac_zero_b) < 2e-3, k              # same dropout pattern
        assert float((a - b).abs().max()) <= 2e-2 * scale, k


@pytest.mark.parametrize("batch,slots", [(8, 2), (24, 4), (13, 4)])
def test_ffn_tail_split_matches_unsplit(batch, slots, monkeypatch):
    """The FF-split of the last partial wave (ffn_sm100.cu: tail tiles cut into FF slices that meet in an fp32 scratch,
    finished by the fix-up kernel) against the same kernels with whole tiles, forward and backward, in training mode.
    EGOT2_FFN_SLOTS shrinks the wave so that a few hundred tokens already have a tail (batch 13: a ragged last tile)."""
    from egot2_b200 import synth
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    case = CASES["hhi3_h128_d30"]
    sp = case.spec
    sd = synth.make_state_dict(sp, 3)
    feats = synth.make_features(sp, batch, case.seg_tokens, 3)
    labels = synth.make_labels(sp, batch, case.seg_tokens, 3).cuda()
    loss_kind, cw = _loss_kind(case)
    monkeypatch.setenv("EGOT2_FFN_SLOTS", str(slots))
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EGOT2_FFN_SPLIT", mode)
        eng = TranslatorEngine(sp, "cuda:0", "bf16")
        eng.arena.load_state_dict(sd)
        eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
        gf = [feats[s.name].cuda().bfloat16() for s in sp.segments]
        act = eng.forward(gf, training=True, seed=5, labels=labels, loss=loss_kind, class_weight=cw)
        eng.backward(act)
        torch.cuda.synchronize()
        res[mode] = {k: act.t[k].float().clone() for k in ("hid0", "y2_0", "x_last", "out", "loss")}
        res[mode]["grad"] = eng.arena.grad.clone()
        if mode == "1":
            assert float(act.t["ffn_scratch"].abs().max()) == 0.0          # left clean for the next launch
    for k in res["0"]:
        a, b = res["1"][k], res["0"][k]
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 2e-2 * scale, k
        if k not in ("grad", "loss"):            # bf16 activations: only the fp32 summation order of GEMM2 differs
            assert float((a != b).float().mean()) < 0.05, k

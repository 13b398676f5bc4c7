#!/usr/bin/env python
"""Per-tensor parity table of one oracle case on the GPU (diagnostics; not a test).

    python tools/diag_case.py hoi_g_h128_l2 bf16 [fp32]

Prints, for the output, the loss and every parameter gradient: max-abs error relative to the reference absmax and the
relative L2 error against the CPU oracle.  Environment switches of the library (EGOT2_GEMM=simt, EGOT2_ATTN=simt,
EGOT2_FFN=unfused) bisect which kernel family a mismatch comes from.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(name, dtype):
    import test_gpu_parity as tg
    from egot2_b200 import _lib as L
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    from oracle.cases import CASES, case_inputs, oracle_forward_loss
    case = CASES[name]
    sp = case.spec
    sd, feats, labels, extra = case_inputs(case)
    eng = TranslatorEngine(sp, "cuda:0", dtype)
    eng.arena.load_state_dict(sd)
    if sp.embed == "task_sinusoid":
        eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
    loss_kind, cw = tg._loss_kind(case)
    gfeats = tg._engine_feats(case, eng, feats, extra, dtype)
    if sp.family in ("hhi_g", "hoi_g"):
        act = eng.forward(gfeats, training=False, labels=labels[:, 1:], loss=loss_kind, prompt=labels[:, :-1])
        out = act.t["out"].float().cpu().view(labels.shape[0], 2, -1).permute(0, 2, 1)
    else:
        act = eng.forward(gfeats, training=False, labels=labels, loss=loss_kind, class_weight=cw)
        out = act.t["out"].float().cpu()
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_out, o_loss = oracle_forward_loss(case, P, feats, labels, extra)
    scale = float(o_out.abs().max())
    print(f"== {name} [{dtype}] env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("EGOT2_")))
    print(f"output   max-abs/absmax {float((out.reshape(o_out.shape) - o_out.detach()).abs().max()) / scale:.3e}")
    if loss_kind == L.LOSS_NONE:
        return
    loss = float(act.t["loss"][0].cpu())
    print(f"loss     {loss:.6f} vs {float(o_loss):.6f}  rel {abs(loss - float(o_loss)) / abs(float(o_loss)):.3e}")
    grad, _ = eng.backward(act)
    names = list(sd.keys())
    o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
    rows = []
    for k, g_ref in zip(names, o_grads):
        if k not in eng.arena.offsets:
            continue
        g = eng.arena.view(k, grad).float().cpu()
        if g_ref is None:
            rows.append((0.0, 0.0, k, "unused", float(g.abs().max())))
            continue
        gs = float(g_ref.abs().max()) + 1e-12
        err = float((g - g_ref).abs().max()) / gs
        l2 = float((g - g_ref).norm()) / (float(g_ref.norm()) + 1e-12)
        rows.append((err, l2, k, tuple(g.shape), gs))
    for err, l2, k, shp, gs in sorted(rows, reverse=True):
        print(f"  {k:60s} {str(shp):18s} max {err:.3e}  l2 {l2:.3e}  absmax_ref {gs:.3e}")


if __name__ == "__main__":
    name = sys.argv[1]
    for dt in sys.argv[2:] or ["bf16"]:
        run(name, dt)

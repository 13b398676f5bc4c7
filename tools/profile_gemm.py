#!/usr/bin/env python
"""Launch one GEMM shape through egot2_gemm a few times (for `ncu --set full -k regex:gemm_sm100`)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egot2_b200 import _lib as L
from egot2_b200.engine import _stream

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=23040); ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--k", type=int, default=128); ap.add_argument("--ta", type=int, default=0)
ap.add_argument("--tb", type=int, default=1); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--f32out", type=int, default=0); ap.add_argument("--acc", type=int, default=0)
a = ap.parse_args()
A = torch.randn((a.k, a.m) if a.ta else (a.m, a.k), device="cuda").bfloat16()
B = torch.randn((a.n, a.k) if a.tb else (a.k, a.n), device="cuda").bfloat16()
bias = torch.zeros(a.n, device="cuda")
C = torch.zeros(a.m, a.n, device="cuda", dtype=torch.float32 if a.f32out else torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters + 3):
    if i == 3:
        e0.record()
    L.call("egot2_gemm", L.BF16, a.m, a.n, a.k, A.data_ptr(), a.ta, B.data_ptr(), a.tb, None if a.acc else bias.data_ptr(),
           0 if a.acc else 1, C.data_ptr(), a.f32out, a.acc, _stream())
e1.record()
torch.cuda.synchronize()
print("impl", L.load().egot2_gemm_last_impl(), "us/launch", e0.elapsed_time(e1) * 1e3 / a.iters)

#!/usr/bin/env python
"""Op-level timing of single kernels through the C ABI (CUDA events, rotating buffers):
   python tools/bench_ops.py [--lib path/to/libegot2.so] [--op attn]"""
import argparse, ctypes as C, os, sys
import torch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "egot2_b200", "lib", "libegot2.so"))
    ap.add_argument("--B", type=int, default=256); ap.add_argument("--T", type=int, default=90)
    ap.add_argument("--H", type=int, default=128); ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--p", type=float, default=0.5); ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    lib = C.CDLL(a.lib)
    lib.egot2_last_error.restype = C.c_char_p
    dev = torch.device("cuda:0")
    B, T, H, heads = a.B, a.T, a.H, a.heads
    NB = 8
    qkv = [torch.randn(B, T, 3 * H, device=dev).bfloat16() for _ in range(NB)]
    out = [torch.empty(B, T, H, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    dout = [torch.randn(B, T, H, device=dev).bfloat16() for _ in range(NB)]
    dqkv = [torch.empty(B, T, 3 * H, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    lse = [torch.empty(B, heads, T, device=dev) for _ in range(NB)]
    st = torch.cuda.current_stream().cuda_stream
    vp = C.c_void_p
    def fwd(i):
        rc = lib.egot2_attention_fwd(1, B, T, H, heads, vp(qkv[i].data_ptr()), vp(out[i].data_ptr()), vp(lse[i].data_ptr()),
                                     C.c_float(a.p), 1, C.c_uint64(7), vp(st))
        assert rc == 0, lib.egot2_last_error()
    def bwd(i):
        rc = lib.egot2_attention_bwd(1, B, T, H, heads, vp(qkv[i].data_ptr()), vp(out[i].data_ptr()), vp(lse[i].data_ptr()),
                                     vp(dout[i].data_ptr()), vp(dqkv[i].data_ptr()), C.c_float(a.p), 1, C.c_uint64(7), None, C.c_size_t(0), vp(st))
        assert rc == 0, lib.egot2_last_error()
    for name, fn in (("attn_fwd", fwd), ("attn_bwd", bwd)):
        for i in range(NB): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters): fn(i % NB)
        e1.record(); torch.cuda.synchronize()
        print(f"{os.path.basename(a.lib)} {name} B{B} T{T} H{H} p{a.p}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us")

if __name__ == "__main__":
    main()

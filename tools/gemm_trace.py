#!/usr/bin/env python
"""Phase trace of the tcgen05 GEMM launches of one training step (needs a -DEGOT2_GEMM_TRACE build, passed with EGOT2_LIB):
CTA (0,0,0) of every launch stamps clock64 at its phases; the table says where a one-tile GEMM spends its microseconds.

  EGOT2_BUILD_TAG=gtrace EGOT2_CFLAGS=-DEGOT2_GEMM_TRACE python -m egot2_b200.build
  EGOT2_LIB=$PWD/egot2_b200/lib/libegot2_gtrace.so python tools/gemm_trace.py [--workload hhi_ttm3_train_b256] [--graphs]
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from egot2_b200 import _lib as L, synth  # noqa: E402
from egot2_b200.trainer import TranslatorTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256")
    ap.add_argument("--graphs", action="store_true", help="trace a CUDA-graph replay (launch slots are those of the capture)")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    spec = wl["spec"]()
    B, seg = wl["batch"], wl["seg_tokens"]
    dev = torch.device("cuda:0")
    tr = TranslatorTrainer(spec, dev, "bf16", use_graphs=args.graphs)
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    f = synth.make_features(spec, B, seg, seed=0, dtype=torch.bfloat16)
    feats = [f[s.name].to(dev) for s in spec.segments]
    labels = synth.make_labels(spec, B, seg, seed=0).to(dev)
    lib = L.load()
    dump = lib.egot2_gemm_trace_dump
    dump.restype = ctypes.c_int
    if args.graphs:
        # the capture assigns the slots; replays overwrite the same slots, so dump after the replays
        for _ in range(5):
            tr.train_step(feats, labels, graph_key=0)
        torch.cuda.synchronize()
        print("# after 5 steps with graphs (slots assigned at capture; values = last replay)", flush=True)
        dump()
        return
    for _ in range(3):
        tr.train_step(feats, labels)
    torch.cuda.synchronize()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    dump()                     # discard the warm-up launches
    os.dup2(saved, 1)
    tr.train_step(feats, labels)
    torch.cuda.synchronize()
    print("# one eager step (launches spaced by the host: every kernel starts on an idle GPU)", flush=True)
    dump()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run a few eager (un-graphed) training steps of a bench workload between cudaProfilerStart/Stop so that
`ncu --profile-from-start off ...` sees exactly the kernels of the step.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from egot2_b200 import synth  # noqa: E402
from egot2_b200.trainer import TranslatorTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fwd-only", action="store_true")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    spec = wl["spec"]()
    B, seg = wl["batch"], wl["seg_tokens"]
    dev = torch.device("cuda:0")
    tr = TranslatorTrainer(spec, dev, args.dtype, use_graphs=False)
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    fdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    pool = []
    for i in range(4):
        f = synth.make_features(spec, B, seg, seed=i, dtype=fdt)
        pool.append(([f[s.name].to(dev) for s in spec.segments], synth.make_labels(spec, B, seg, seed=i).to(dev)))

    def step(i):
        fe, la = pool[i % len(pool)]
        if args.fwd_only:
            tr.infer(fe)
        else:
            tr.train_step(fe, la)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.steps):
        step(args.warmup + i)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()

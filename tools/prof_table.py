#!/usr/bin/env python
"""Per-launcher CUDA-event timing table (the library's egot2_prof_* hooks) for a bench workload.

  python tools/prof_table.py [--workload W] [--mode train|eval|train_fwd] [--steps N] [--batch B]
"""
import argparse
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
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--mode", default="train", choices=["train", "eval", "train_fwd"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
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
        if args.mode == "eval":
            tr.infer(fe)
        elif args.mode == "train_fwd":
            tr.engine.forward(fe, training=True, seed=i, labels=la, loss=tr.loss_kind, class_weight=tr.class_weight,
                              persistent=True)
        else:
            tr.train_step(fe, la)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    L.prof_enable(True)
    for i in range(args.steps):
        step(3 + i)
    torch.cuda.synchronize()
    rows = L.prof_report()
    L.prof_enable(False)
    tot = sum(r[2] for r in rows)
    es = 2 if args.dtype == "bf16" else 4
    print(f"# {args.workload} [{args.dtype}] mode={args.mode} B={B}: {tot / args.steps:.1f} us of kernel time per step, "
          f"{sum(r[1] for r in rows) / args.steps:.0f} launchers per step")
    print(f"# {'us/step':>9} {'n/step':>6} {'share':>6} {'TF/s':>7} {'GB/s':>7}  launcher")
    for tag, n, us in rows:
        c = bench.launcher_cost(tag, es)
        tf = c[0] * n / (us * 1e-6) / 1e12 if c else float("nan")
        gb = c[1] * n / (us * 1e-6) / 1e9 if c else float("nan")
        print(f"{us / args.steps:10.1f} {n / args.steps:6.1f} {100 * us / tot:5.1f}% {tf:7.1f} {gb:7.0f}  {tag}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Event-timed FFN launchers (and optionally all launchers) of eager training steps for the library in EGOT2_LIB.
   EGOT2_LIB=egot2_b200/lib/libegot2_dbg1.so python tools/ffn_time.py [--workload hhi_ttm3_train_b256] [--all]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--grep", default="ffn")
    a = ap.parse_args()
    wl = dict(bench.WORKLOADS[a.workload])
    spec = wl["spec"]()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    tr = bench.build_trainer(a.workload, wl, spec, dev, "bf16", use_graphs=False)
    pool, _ = bench.build_pool(wl, spec, wl["batch"], dev, "bf16", 0, max_pool=6)
    rows = bench.profile_step_launchers(tr, pool, len(pool), a.steps, 2)
    tot = sum(r[2] for r in rows)
    tag = os.path.basename(os.environ.get("EGOT2_LIB", "libegot2.so"))
    print(f"{tag}: sum of launcher time {tot / a.steps:.1f} us/step")
    for t, n, us in rows:
        if a.all or a.grep in t:
            print(f"  {tag:22s} {t:60s} {us / n:7.1f} us/launch x {n / a.steps:.0f}")


if __name__ == "__main__":
    main()

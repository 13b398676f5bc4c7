#!/usr/bin/env python
"""A/B runs of bench.py on the GPU box: every (environment setting, workload) pair, one summary line each.

  python tools/ab.py --tag r2b --env "" "EGOT2_DEFER_JOIN=0" --workloads hhi_ttm3_train_b256 hoi_pnr_train_b256 [--breakdown 8]

Full JSON lines land in gpurun_out/ab_<tag>_<i>_<workload>.json.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="ab")
    ap.add_argument("--env", nargs="*", default=[""])
    ap.add_argument("--workloads", nargs="*", default=["hhi_ttm3_train_b256"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--breakdown", type=int, default=0)
    ap.add_argument("--extra", nargs="*", default=[])
    args = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for i, envs in enumerate(args.env):
        env = dict(os.environ)
        for kv in envs.split():
            k, v = kv.split("=", 1)
            env[k] = v
        for wl in args.workloads:
            out = os.path.join(ROOT, "gpurun_out", f"ab_{args.tag}_{i}_{wl}.json")
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--skip-cpu-baseline",
                   "--steps", str(args.steps), "--warmup", str(args.warmup)] + args.extra
            try:
                r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=400)
            except subprocess.TimeoutExpired:
                print(f"[{envs or 'default'}] {wl}: TIMEOUT")
                continue
            open(out, "w").write(r.stdout)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                print(f"[{envs or 'default'}] {wl}: rc={r.returncode} no JSON; stderr tail: {r.stderr[-400:]}")
                continue
            rf = d.get("roofline", {})
            print(f"[{envs or 'default'}] {wl}: {d['value']:.0f} {d['unit']}  {d['ms_per_step']*1000:.1f} us/step  "
                  f"e2e {d['e2e']['value']:.0f}  top {rf.get('kernel')} frac {rf.get('frac')}")
            for b in rf.get("breakdown", [])[:args.breakdown]:
                print("    %-64s %6.1f us/step %4.1f%%" % (b["launcher"], b["us_per_step"], 100 * b["share"]))


if __name__ == "__main__":
    main()

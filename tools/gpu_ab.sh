#!/bin/bash
# A/B of an environment switch on the default bench: tools/gpu_ab.sh TAG VAR=off_value  (runs pytest -m gpu first)
tag=${1:-ab}; sw=${2:-EGOT2_PDL=0}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
for mode in on off; do
  if [ $mode = off ]; then export $sw; fi
  timeout 300 python bench.py --skip-cpu-baseline > gpurun_out/bench_hhi_${tag}_$mode.json 2> gpurun_out/bench_hhi_${tag}_$mode.err
  tail -c 400 gpurun_out/bench_hhi_${tag}_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_hhi_${tag}_$mode.json").read().strip().splitlines()[-1])
    print("$mode value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    r = d["roofline"]; print("roofline", r["kernel"], r["bound"], r["achieved"], r["frac"], "share", r["share_of_step"], "sum_kernel_us", r["sum_kernel_us_per_step"])
    for b in r["breakdown"]: print("  %-60s %5.1f us/step %4.1f%%" % (b["launcher"], b["us_per_step"], 100*b["share"]))
except Exception as e:
    print("bench parse failed", e)
PY
done

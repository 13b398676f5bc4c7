#!/bin/bash
# One GPU round-trip: parity tests, the default bench line, (optional) launch list.  Usage: tools/gpu_check.sh [tag]
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --skip-cpu-baseline > gpurun_out/bench_hhi_$tag.json 2> gpurun_out/bench_hhi_$tag.err
tail -c 600 gpurun_out/bench_hhi_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_hhi_$tag.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    r = d["roofline"]; print("roofline", r["kernel"], r["bound"], r["achieved"], r["frac"], "share", r["share_of_step"])
    for b in r["breakdown"]: print("  %-60s %5.1f us/step %4.1f%%  %s TF %s GB/s" % (b["launcher"], b["us_per_step"], 100*b["share"], round(b.get("tflops",0),1), round(b.get("gbs",0),1)))
except Exception as e:
    print("bench parse failed", e)
PY

#!/bin/bash
# First GPU round-trip for code that was written without hardware: the xfail(non-strict) parity tests of
# oracle.cases.UNVALIDATED_ON_GPU (+ greedy decoding, fused AdamW) with full tracebacks, then the bench lines of the new
# workloads.  Usage (under gpurun): tools/gpu_pending.sh [tag] [memcheck]   -> gpurun_out/pending_<tag>.log, bench_*_<tag>.json
# "memcheck" as the second argument first runs the fp32 tests of the NEW device code (vit.cu, attention_wide.cu, the H = 2048
# LayerNorm instantiations) under compute-sanitizer (slow: a few minutes).
tag=${1:-run}
mkdir -p gpurun_out
if [ "$2" = "memcheck" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_unvalidated_gpu.py -m gpu -q \
    --runxfail -k "fp32 and (vit or lta2_h2048)" -p no:cacheprovider > gpurun_out/memcheck_$tag.log 2>&1
  echo "memcheck rc=$?" >> gpurun_out/memcheck_$tag.log
  grep -E "ERROR SUMMARY|Invalid|passed|failed|rc=" gpurun_out/memcheck_$tag.log | tail -12
fi
# --runxfail: report real pass / fail (with tracebacks) instead of XPASS / XFAIL
timeout 900 python -m pytest tests/test_zz_unvalidated_gpu.py -m gpu -q --runxfail -rA --tb=short \
  > gpurun_out/pending_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pending_$tag.log
grep -E "PASSED|FAILED|ERROR|passed|failed|rc=" gpurun_out/pending_$tag.log | tail -40
for wl in hoi_g_train hhi_g_train; do
  timeout 300 python bench.py --workload $wl --skip-cpu-baseline --steps 20 --warmup 3 \
    > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err
  tail -c 400 gpurun_out/bench_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${wl}_$tag.json").read().strip().splitlines()[-1])
    print("$wl", "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d.get("gpu_launches"))
except Exception as e:
    print("$wl bench parse failed", e)
PY
done
# forward-only throughput (the other half of the metric; bench.py reports the training step)
for wl in hhi_ttm3_train_b256 hoi_pnr_train_b256 hoi_lta_train_b512; do
  timeout 300 python tools/bench_infer.py --workload $wl > gpurun_out/infer_${wl}_$tag.json 2> gpurun_out/infer_${wl}_$tag.err
  tail -c 300 gpurun_out/infer_${wl}_$tag.err; tail -c 400 gpurun_out/infer_${wl}_$tag.json | cut -c1-300
done
# EgoT2-g steps replayed from a CUDA graph (default off until this has passed once)
for wl in hhi_g_train hoi_g_train; do
  EGOT2_G_GRAPH=1 timeout 300 python bench.py --workload $wl --skip-cpu-baseline --steps 20 --warmup 3 \
    > gpurun_out/bench_${wl}_graph_$tag.json 2> gpurun_out/bench_${wl}_graph_$tag.err
  tail -c 300 gpurun_out/bench_${wl}_graph_$tag.err; tail -c 600 gpurun_out/bench_${wl}_graph_$tag.json | cut -c1-200
done

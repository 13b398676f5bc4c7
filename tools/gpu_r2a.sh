#!/bin/bash
# round 2, first GPU round-trip: the whole -m gpu suite with real pass/fail of the formerly-unvalidated cases, the per-tensor
# diagnosis of hoi_g_h128_l2 [bf16] under the kernel-family switches, and baseline bench lines of the other configs.
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --runxfail -rf --tb=short -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
grep -E "FAILED|ERROR|passed|failed|rc=" gpurun_out/pytest_$tag.log | tail -30
for env in "X=1" "EGOT2_GEMM=simt" "EGOT2_ATTN=simt" "EGOT2_GEMM=simt EGOT2_ATTN=simt" "EGOT2_STREAMS=0" "EGOT2_PDL=0"; do
  env $env timeout 300 python tools/diag_case.py hoi_g_h128_l2 bf16 >> gpurun_out/diag_hoi_g_$tag.log 2>&1
done
timeout 300 python tools/diag_case.py hoi_g6_clip_h128_l1 bf16 >> gpurun_out/diag_hoi_g_$tag.log 2>&1
grep -E "^==|output|loss|max [0-9]" gpurun_out/diag_hoi_g_$tag.log | awk '/^==/{n=0} {n++; if (n<=8) print}' | cut -c1-170
for wl in hoi_pnr_train_b256 hoi_lta_train_b512 hhi_g_train; do
  timeout 300 python bench.py --workload $wl --skip-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err
  tail -c 300 gpurun_out/bench_${wl}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${wl}_$tag.json").read().strip().splitlines()[-1])
    print("$wl", "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for b in d["roofline"]["breakdown"]: print("  %-60s %6.1f us/step %4.1f%%" % (b["launcher"], b["us_per_step"], 100*b["share"]))
except Exception as e:
    print("$wl bench parse failed", e)
PY
done
EGOT2_G_GRAPH=1 timeout 300 python bench.py --workload hhi_g_train --skip-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_hhi_g_graph_$tag.json 2> gpurun_out/bench_hhi_g_graph_$tag.err
tail -c 300 gpurun_out/bench_hhi_g_graph_$tag.err; cut -c1-200 gpurun_out/bench_hhi_g_graph_$tag.json

#!/usr/bin/env python
"""Text summary of an .ncu-rep (ncu --set full): one block per kernel launch with the metrics DESIGN.md / bench.py cite.
   python tools/ncu_summary.py report.ncu-rep [> profiles/rNN_ncu_xxx.txt]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== ", r[idx["Kernel Name"]][:140])
        for w in WANT:
            if w in idx:
                print(f"{w:<82} {r[idx[w]]:>16} {units[idx[w]]}")
        stalls = [(h, r[i]) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and r[i] not in ("", "0")]
        for h, v in sorted(stalls, key=lambda kv: -float(kv[1].replace(",", "")))[:7]:
            print(f"{h:<82} {v:>16} warp")
        print()


if __name__ == "__main__":
    main(sys.argv[1])

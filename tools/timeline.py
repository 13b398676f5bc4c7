#!/usr/bin/env python
"""In-graph kernel timeline of one training step (needs a -DEGOT2_TIMELINE build of the library, passed with EGOT2_LIB):
every kernel stamps %globaltimer when it starts; the stamps of one CUDA-graph replay, sorted by time, show the real
start order / gaps of the main chain and the side streams.

  EGOT2_BUILD_TAG=timeline EGOT2_CFLAGS=-DEGOT2_TIMELINE python -m egot2_b200.build      (-> egot2_b200/lib/libegot2_timeline.so)
  EGOT2_LIB=$PWD/egot2_b200/lib/libegot2_timeline.so python tools/timeline.py
"""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from egot2_b200 import _lib as L, synth  # noqa: E402
from egot2_b200.trainer import TranslatorTrainer  # noqa: E402

FILES = {1: "rowops.cu", 2: "loss.cu", 3: "head_fused.cu", 4: "gemm_sm100.cu", 5: "ffn_sm100.cu", 6: "attention_mma.cu",
         7: "decoder.cu", 8: "attention_small.cu", 9: "attention_simt.cu", 10: "gemm_simt.cu", 11: "api.cu", 12: "vit.cu",
         13: "attention_wide.cu", 14: "embed_extra.cu", 15: "metrics.cu", 16: "peer.cu", 17: "attention_long.cu"}


def kernel_at(fid, line, cache={}):
    """name of the __global__ function enclosing `line` of csrc file `fid`"""
    f = FILES.get(fid)
    if f is None:
        return f"?{fid}:{line}"
    if f not in cache:
        cache[f] = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "egot2_b200", "csrc", f)).read().splitlines()
    src = cache[f]
    for i in range(min(line, len(src)) - 1, -1, -1):
        if "__global__" in src[i] or (i + 1 < len(src) and "__global__" in src[i]):
            m = re.search(r"(\w+)\s*\(", " ".join(src[i:i + 3]).split("__global__", 1)[1].replace("__launch_bounds__", "").replace("__cluster_dims__", ""))
            txt = " ".join(src[i:i + 4])
            m = re.findall(r"(\w+_kernel)\b", txt)
            return m[0] if m else f"{f}:{line}"
    return f"{f}:{line}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256")
    ap.add_argument("--replays", type=int, default=3)
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS[args.workload])
    spec = wl["spec"]()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    tr = bench.build_trainer(args.workload, wl, spec, dev, "bf16", use_graphs=True)      # any workload, incl. the EgoT2-g trainers
    pool, _ = bench.build_pool(wl, spec, wl["batch"], dev, "bf16", 0, max_pool=1)
    feats, labels = pool[0]
    for _ in range(4):
        tr.train_step(feats, labels, graph_key=0)
    torch.cuda.synchronize()
    buf = torch.zeros(1 + 2 * 4000, device=dev, dtype=torch.int64)
    L.call("egot2_timeline_set", buf.data_ptr())
    for r in range(args.replays):
        buf.zero_()
        torch.cuda.synchronize()
        tr.train_step(feats, labels, graph_key=0)
        torch.cuda.synchronize()
        h = buf.cpu().tolist()
        n = min(h[0] & 0xffffffff, 4000)
        ev = sorted((h[1 + 2 * i], h[2 + 2 * i]) for i in range(n))
        if r < args.replays - 1:
            continue
        t0 = ev[0][0]
        print(f"# replay {r}: {n} kernel starts, span {(ev[-1][0] - t0) / 1e3:.1f} us (last start)")
        print("#  start_us   +delta  kernel")
        prev = t0
        for t, loc in ev:
            print(f"{(t - t0) / 1e3:10.2f} {(t - prev) / 1e3:8.2f}  {kernel_at(loc // 100000, loc % 100000)}  [{FILES.get(loc // 100000, '?')}:{loc % 100000}]")
            prev = t
    L.call("egot2_timeline_set", None)


if __name__ == "__main__":
    main()

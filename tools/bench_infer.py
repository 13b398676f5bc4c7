"""Forward-only (inference) throughput of a translator workload: the other half of BASELINE.json's metric ("translator
clips/sec ... forward and forward+backward"); `bench.py` reports the training step.

    python tools/bench_infer.py [--workload hhi_ttm3_train_b256] [--dtype bf16] [--steps 200] [--no-graph]

One JSON line: device-resident clips/s (CUDA events around K forwards over a pool of input batches larger than L2, the
launch sequence replayed from one CUDA graph per batch) and end-to-end clips/s from pinned host buffers (H2D of the
features + forward + D2H of the logits inside the timed region).  Eval mode (no dropout), the trainer's `infer` path.
Written at the end of round 1 without hardware: run it first thing when a GPU is available."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                       # noqa: E402
from egot2_b200 import synth                      # noqa: E402
from egot2_b200.trainer import TranslatorTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256", choices=sorted(w for w, d in bench.WORKLOADS.items() if not d.get("prompt")))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    spec = wl["spec"]()
    B, seg = wl["batch"], wl["seg_tokens"]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    fdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tr = TranslatorTrainer(spec, dev, args.dtype, use_graphs=False)
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    feat_bytes = spec.feature_elems_per_clip(seg) * B * (2 if args.dtype == "bf16" else 4)
    n_pool = min(64, max(3, -(-2 * bench.L2_BYTES // feat_bytes)))
    pool = []
    for i in range(n_pool):
        f = synth.make_features(spec, B, seg, seed=100 + i, dtype=fdt)
        pool.append([f[s.name].to(dev) for s in spec.segments])

    # one graph per pool entry (the persistent activations are shared: replays are serialized on one stream)
    graphs = [None] * n_pool
    for _ in range(max(3, args.warmup)):
        tr.infer(pool[0])
    if args.dtype == "bf16":          # weights are frozen here: cast the arena once, keep the cast out of the timed forwards
        tr.engine.arena.refresh_shadow()
        tr.engine.arena.shadow_fresh = True
    torch.cuda.synchronize(dev)
    if not args.no_graph:
        for i in range(n_pool):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                tr.infer(pool[i])
            graphs[i] = g

    def step(i):
        if graphs[i % n_pool] is not None:
            graphs[i % n_pool].replay()
        else:
            tr.infer(pool[i % n_pool])

    for i in range(max(args.warmup, n_pool)):
        step(i)
    torch.cuda.synchronize(dev)
    clocks = bench.ClockSampler(0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i + 1)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop()

    # end to end: pinned host features -> device -> forward -> logits back on the host, every step
    host = []
    for i in range(2):
        f = synth.make_features(spec, B, seg, seed=900 + i, dtype=fdt)
        host.append([f[s.name].pin_memory() for s in spec.segments])
    dbuf = [[torch.empty_like(t, device=dev) for t in host[0]] for _ in range(2)]
    out_host = None
    e2e_steps = max(5, min(args.steps, 50))

    def e2e(n):
        nonlocal out_host
        for i in range(n):
            for d, h in zip(dbuf[i % 2], host[i % 2]):
                d.copy_(h, non_blocking=True)
            out = tr.infer(dbuf[i % 2])
            if out_host is None:
                out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e(3)
    e0.record()
    e2e(e2e_steps)
    e1.record()
    torch.cuda.synchronize(dev)
    ms2 = e0.elapsed_time(e1)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    line = {"metric": "translator fwd clips/sec", "value": B * args.steps / (ms * 1e-3), "unit": "clips/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "mode": "eval forward (no dropout)", "clips_per_gpu": B,
                       "tokens_per_clip": sum(seg), "cuda_graphs": not args.no_graph,
                       "l2_policy": f"inputs rotate over a pool of {n_pool} batches = {n_pool * feat_bytes / 2**20:.0f} MiB"},
            "model_tflops_per_s": spec.flops_per_clip(seg, backward=False) * B * args.steps / (ms * 1e-3) / 1e12,
            "e2e": {"value": B * e2e_steps / (ms2 * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
            "clocks": clock_info}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Experiment: does running the training step as TWO half-batches side by side (two streams, two branches of one CUDA
graph, gradients summed) beat one full batch?  Every kernel of the b256 step is latency-bound and leaves most SMs idle.
Times forward+backward only (no Adam) of: one engine on the full batch; two engines sharing the parameter arena on
halves of it."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from egot2_b200 import _lib as L, synth  # noqa: E402
from egot2_b200.engine import TranslatorEngine  # noqa: E402
from egot2_b200.trainer import default_loss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hhi_ttm3_train_b256")
    ap.add_argument("--parts", type=int, default=2)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    spec = wl["spec"]()
    B, seg = wl["batch"], wl["seg_tokens"]
    dev = torch.device("cuda:0")
    P = args.parts
    engs = [TranslatorEngine(spec, dev, "bf16")]
    for _ in range(P - 1):
        engs.append(TranslatorEngine(spec, dev, "bf16", arena=engs[0].arena))
    engs[0].arena.load_state_dict(synth.make_state_dict(spec, 0))
    if spec.embed == "task_sinusoid":
        from egot2_b200.hhi import PositionalEncoding
        for e in engs:
            e.set_sinusoid(PositionalEncoding(spec.hidden).pe)
    engs[0].arena.refresh_shadow()
    f = synth.make_features(spec, B, seg, seed=0, dtype=torch.bfloat16)
    feats = [f[s.name].to(dev) for s in spec.segments]
    labels = synth.make_labels(spec, B, seg, seed=0).to(dev)
    loss_kind, cw = default_loss(spec)
    cw = None if cw is None else cw.to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(P - 1)]
    grads = [torch.zeros_like(engs[0].arena.grad) for _ in range(P - 1)]
    step = B // P

    def one():
        act = engs[0].forward(feats, training=True, seed=1, labels=labels, loss=loss_kind, class_weight=cw, persistent=True)
        engs[0].backward(act, zero_grad=False)

    def parts():
        cur = torch.cuda.current_stream(dev)
        for i in range(P):
            fi = [x[i * step:(i + 1) * step] for x in feats]
            li = labels[i * step:(i + 1) * step]
            if i > 0:
                streams[i - 1].wait_stream(cur)
            ctx = torch.cuda.stream(streams[i - 1]) if i > 0 else torch.cuda.stream(cur)
            with ctx:
                act = engs[i].forward(fi, training=True, seed=1 + i, labels=li, loss=loss_kind, class_weight=cw, persistent=True)
                engs[i].backward(act, zero_grad=False, dloss_scale=1.0 / P, grad=grads[i - 1] if i > 0 else None)
        for s in streams:
            cur.wait_stream(s)

    def timed(body, tag):
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"{tag}: {e0.elapsed_time(e1) / args.steps * 1e3:.1f} us per forward+backward of {B} clips", flush=True)

    timed(one, f"{args.workload} one batch of {B}")
    timed(parts, f"{args.workload} {P} x {step} side by side")


if __name__ == "__main__":
    main()

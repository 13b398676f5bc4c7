#!/usr/bin/env python
"""Latency of the peer-memory exchange kernel (csrc/peer.cu) against NCCL all-reduce + fused Adam, N ranks, CUDA events.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egot2_b200 import specs  # noqa: E402
from egot2_b200.engine import TranslatorEngine  # noqa: E402
from egot2_b200.parallel import PeerExchange  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    hp = dict(lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    for name, spec in (("hhi_ttm3 (0.69 M)", specs.hhi_ttm_spec(128, 4, 1, 0.5, True)), ("hoi_pnr (3.2 M)", specs.hoi_pnr_spec(128, 6, 16, 0.5, 0.1)),
                       ("hoi_lta (28 M)", specs.hoi_lta_spec(512, 4, 8, 0.5))):
        eng = TranslatorEngine(spec, dev, "bf16")
        px = PeerExchange(eng, None)
        n, nb = eng.arena.numel, eng.arena.embed_numel
        state = {}
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn, iters=200):
            for _ in range(10):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        t_full = timed(lambda: px.step(state, 1, hp, st))
        t_pre = timed(lambda: px.step(state, 1, hp, st, lo=0, hi=nb, channel=1))
        t_rest = timed(lambda: px.step(state, 1, hp, st, lo=nb, hi=n, channel=0))
        g = torch.zeros(n, device=dev)
        t_nccl = timed(lambda: dist.all_reduce(g))
        t_nccl_pre = timed(lambda: dist.all_reduce(g[:nb]))
        t_adam = timed(lambda: eng.adam_step(state, 1, fused=True))
        if rank == 0:
            print(f"{name:20s} world {world}: peer kernel full {t_full:6.1f} us | prefix ({nb} el) {t_pre:6.1f} | rest {t_rest:6.1f} || "
                  f"NCCL all-reduce full {t_nccl:6.1f} | prefix {t_nccl_pre:6.1f} | fused Adam {t_adam:6.1f}", flush=True)
        px.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

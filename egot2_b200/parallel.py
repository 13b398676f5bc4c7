"""Data-parallel plumbing of the translator path: clips are independent units, so they
shard across ranks with NO data-path collective; the only exchange of a training step is the all-reduce of
the flat translator-gradient arena (what the reference gets from Lightning DDP's bucketed all-reduce:
HOI/scripts/lta/run_lta.py:249, HOI/configs/recognition/defaults.py:494; DP averaging of per-GPU losses:
HOI/tasks/pnr/video_task.py:39-42).

Everything here is host logic over torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_clips: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced clip range [lo, hi) of `rank`: the first n % world ranks get one extra clip
    (reference: DistributedSampler's per-GPU batch = BATCH_SIZE / NUM_GPUS, HOI/dataset/lta/loader.py:73-74)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_clips(tensors: Sequence[torch.Tensor], world: int, rank: int):
    """Slice every per-clip tensor (leading dimension = clips) to this rank's shard."""
    lo, hi = shard_range(int(tensors[0].shape[0]), world, rank)
    return [t[lo:hi] for t in tensors]


def allreduce_gradients(flat_grad: torch.Tensor, group=None, local_weight: Optional[float] = None) -> float:
    """Combine the ranks' flat gradient arenas in place with ONE all-reduce and return the scale the optimizer must
    apply to the summed buffer.

    local_weight None  -> DDP semantics (what the reference does): mean over ranks, scale = 1 / world.
    local_weight w_r   -> exact full-batch gradient for mean-type losses with unequal shards: every rank scales its
                          arena by w_r first (w_r = its clips, or its summed class weights for the weighted CE), the
                          buffers are summed and the returned scale is 1 / sum_r w_r.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 1.0
    if local_weight is None:
        dist.all_reduce(flat_grad, group=group)
        return 1.0 / world
    w = torch.tensor([float(local_weight)], dtype=torch.float64, device=flat_grad.device)
    flat_grad.mul_(float(local_weight))
    dist.all_reduce(flat_grad, group=group)
    dist.all_reduce(w, group=group)
    return 1.0 / float(w.item())


def gather_outputs(local_out: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """Inference: concatenate the ranks' per-clip outputs in clip order (shards may differ by one clip)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_out
    base, extra = divmod(n_clips, world)
    pad_rows = base + (1 if extra else 0)
    padded = local_out.new_zeros((pad_rows,) + tuple(local_out.shape[1:]))
    padded[: local_out.shape[0]] = local_out
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_clips, world, r)
        out.append(p[: hi - lo])
    return torch.cat(out, dim=0)

"""World-size-2 gloo tests (CPU) of the data-parallel host logic: clip sharding with no data-path collective, the
single flat gradient all-reduce, and output gathering.  The per-rank compute is the CPU oracle (test infrastructure);
on the GPUs the same functions run over NCCL with libegot2 producing the gradients."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egot2_b200 import parallel as par


def test_shard_range_covers_and_balances():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def _flat_grads(case_name, lo, hi):
    """Oracle loss + flat gradient (fixed parameter order) on clips [lo, hi) of a golden case."""
    from oracle.cases import CASES, case_inputs, oracle_forward_loss
    case = CASES[case_name]
    sd, feats, labels, extra = case_inputs(case)
    feats = {k: v[lo:hi] for k, v in feats.items()}
    extra = {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == labels.shape[0] else v) for k, v in extra.items()}
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out, loss = oracle_forward_loss(case, P, feats, labels[lo:hi], extra)
    names = sorted(P)
    grads = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(P[k])).reshape(-1) for k, g in zip(names, grads)])
    return out.detach(), float(loss), flat, labels[lo:hi]


def _worker(rank, world, port, case_name, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        lo, hi = par.shard_range(n_clips, world, rank)
        out, loss, flat, labels = _flat_grads(case_name, lo, hi)
        # (1) DDP semantics: mean over ranks
        g = flat.clone()
        scale = par.allreduce_gradients(g)
        ddp = g * scale
        # (2) exact full-batch gradient for the class-weighted CE: weight = the shard's summed class weights
        cw = torch.tensor([0.266, 0.734])
        g2 = flat.clone()
        scale2 = par.allreduce_gradients(g2, local_weight=float(cw[labels].sum()))
        exact = g2 * scale2
        full_out = par.gather_outputs(out, n_clips)
        if rank == 0:
            q.put(tuple(t.detach().numpy().copy() for t in (ddp, exact, full_out)))      # plain arrays: no fd passing race with the exiting worker
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_single_process():
    case_name, n_clips, world = "hhi3_h128_l1", 5, 2          # 5 clips over 2 ranks: unequal shards (3 + 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    ddp, exact, full_out = (torch.from_numpy(a) for a in q.get(timeout=240))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process references
    out_all, _, flat_all, _ = _flat_grads(case_name, 0, n_clips)
    parts = [_flat_grads(case_name, *par.shard_range(n_clips, world, r))[2] for r in range(world)]
    assert torch.allclose(ddp, sum(parts) / world, rtol=1e-5, atol=1e-7)           # reference DDP: mean of rank gradients
    assert torch.allclose(exact, flat_all, rtol=2e-4, atol=1e-6)                   # weighted: the full-batch gradient
    assert torch.equal(full_out, out_all) or torch.allclose(full_out, out_all, rtol=1e-6, atol=1e-7)


def _overlap_worker(rank, world, port, q):
    """The trainer's overlapped schedule (trainer._dp_overlap_step) on the real arena layout: the tail
    [embed_numel:] (head + encoder-layer gradients, complete after the first backward stage) is all-reduced
    asynchronously while the embedding stage still writes the prefix [0, embed_numel), which is reduced afterwards."""
    from egot2_b200 import specs
    from egot2_b200.engine import ParamArena
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arena = ParamArena(specs.hhi_ttm_spec(128, 4, 1, 0.5, True), torch.device("cpu"))
        g = torch.Generator().manual_seed(100 + rank)
        full = torch.randn(arena.numel, generator=g)
        nb = arena.embed_numel
        arena.grad[nb:] = full[nb:]                       # stage 1 done: everything behind the embedding stage
        work = dist.all_reduce(arena.grad[nb:], async_op=True)
        arena.grad[:nb] = full[:nb]                       # stage 2 (embedding backward) lands while the tail is in flight
        dist.all_reduce(arena.grad[:nb])
        work.wait()
        whole = full.clone()
        dist.all_reduce(whole)                            # the single-bucket schedule
        if rank == 0:
            q.put((arena.grad.numpy().copy(), whole.numpy().copy(), nb, arena.numel))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_piece_overlapped_allreduce_equals_single_bucket():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    two_piece, whole, nb, numel = q.get(timeout=240)
    two_piece, whole = torch.from_numpy(two_piece), torch.from_numpy(whole)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert 0 < nb < numel
    assert torch.equal(two_piece, whole)                  # the two slices tile the arena exactly; same sums
    expect = sum(torch.randn(numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    assert torch.allclose(whole, expect, rtol=0, atol=1e-6)


def test_peer_slab_layout_is_aligned_and_disjoint():
    """parallel.slab_layout: the byte offsets every rank derives for its IPC slab (parameters | gradients | bf16 shadow | flag
    words) from the arena size alone - 256-byte aligned, non-overlapping, identical for identical arenas (csrc/peer.cu relies on it)."""
    from egot2_b200.parallel import slab_layout
    for numel in (64, 692928, 28137600):
        for shadow in (True, False):
            p, g, s, f, total = slab_layout(numel, shadow, 256)
            assert p == 0 and g >= p + 4 * numel and f % 256 == 0 and g % 256 == 0 and total == f + 256
            if shadow:
                assert s >= g + 4 * numel and s % 256 == 0 and f >= s + 2 * numel
            else:
                assert s == -1 and f >= g + 4 * numel
            assert slab_layout(numel, shadow, 256) == (p, g, s, f, total)


def test_bind_to_gpu_numa_is_best_effort():
    """No CUDA device here: the binding must decline quietly, never raise."""
    from egot2_b200.parallel import bind_to_gpu_numa
    assert bind_to_gpu_numa(0) is None or isinstance(bind_to_gpu_numa(0), tuple)

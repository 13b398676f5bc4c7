"""Trainer-level checks on the GPU: the whole-step CUDA graph (forward + backward + fused Adam with the step count on the
device) must train exactly like the same graph followed by an eager Adam launch that gets the step count from the host."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(monkeypatch, graph_update: str, steps: int = 4):
    from egot2_b200 import specs, synth
    from egot2_b200.trainer import TranslatorTrainer
    monkeypatch.setenv("EGOT2_GRAPH_UPDATE", graph_update)
    spec = specs.hhi_ttm_spec(128, 4, 1, 0.5, True)
    dev = torch.device("cuda:0")
    tr = TranslatorTrainer(spec, dev, "bf16", use_graphs=True)
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    seg = (30, 30, 30)
    f = synth.make_features(spec, 16, seg, seed=1, dtype=torch.bfloat16)
    feats = [f[s.name].to(dev) for s in spec.segments]
    labels = synth.make_labels(spec, 16, seg, seed=1).to(dev)
    losses = []
    for _ in range(steps):
        losses.append(float(tr.train_step(feats, labels, graph_key=0)))
    torch.cuda.synchronize()
    return tr.engine.arena.param.clone(), tr.engine.arena.shadow.float().clone(), tr.engine.arena.grad.clone(), losses


def test_whole_step_graph_matches_eager_update(monkeypatch):
    p1, s1, g1, l1 = _run(monkeypatch, "1")
    p0, s0, g0, l0 = _run(monkeypatch, "0")
    assert float(g1.abs().max()) == 0.0 and float(g0.abs().max()) == 0.0        # the fused Adam cleared the arena
    # the parameter-gradient sums use fp32 atomics (split-K), so two runs agree to rounding, not bit for bit
    assert float((p1 - p0).abs().max()) <= 5e-4 * float(p0.abs().max())
    assert float((s1 - p1).abs().max()) <= 1e-2 * float(p1.abs().max())          # bf16 shadow tracks the parameters
    assert all(abs(a - b) <= 2e-2 * abs(b) + 1e-4 for a, b in zip(l1, l0))
    assert l1[-1] != l1[0]                                                       # the parameters moved


def test_staged_backward_matches_single_call():
    """backward(stage='pre_embed') + backward(stage='embed') == backward() (the split the data-parallel trainer captures
    as two graphs)."""
    from egot2_b200 import _lib as L, specs, synth
    from egot2_b200.engine import TranslatorEngine
    from egot2_b200.hhi import PositionalEncoding
    spec = specs.hhi_ttm_spec(128, 4, 1, 0.5, True)
    dev = torch.device("cuda:0")
    seg = (30, 30, 30)
    f = synth.make_features(spec, 12, seg, seed=2, dtype=torch.bfloat16)
    feats = [f[s.name].to(dev) for s in spec.segments]
    labels = synth.make_labels(spec, 12, seg, seed=2).to(dev)
    grads = []
    for staged in (False, True):
        eng = TranslatorEngine(spec, dev, "bf16")
        eng.arena.load_state_dict(synth.make_state_dict(spec, 0))
        eng.set_sinusoid(PositionalEncoding(spec.hidden).pe)
        act = eng.forward(feats, training=True, seed=9, labels=labels, loss=L.LOSS_CE,
                          class_weight=torch.tensor([0.266, 0.734], device=dev))
        if staged:
            eng.backward(act, stage="pre_embed")
            assert float(eng.arena.grad[:eng.arena.embed_numel].abs().max()) == 0.0    # embedding bucket untouched so far
            eng.backward(act, zero_grad=False, stage="embed")
        else:
            eng.backward(act)
        torch.cuda.synchronize()
        grads.append(eng.arena.grad.clone())
    assert float((grads[0] - grads[1]).abs().max()) <= 1e-3 * float(grads[0].abs().max())
    assert float(grads[1][:eng.arena.embed_numel].abs().max()) > 0.0


def test_dp_overlap_step_single_rank(monkeypatch):
    """The two-graph data-parallel step (all-reduce of arena[embed_numel:] overlapping the embedding-backward graph) on a
    one-rank NCCL group must train like the plain graphed step."""
    import torch.distributed as dist
    p0, _, _, l0 = _run(monkeypatch, "0")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29577", world_size=1, rank=0)
        created = True
    try:
        monkeypatch.setenv("EGOT2_DP_OVERLAP", "force")
        p1, _, g1, l1 = _run(monkeypatch, "0")
    finally:
        if created:
            dist.destroy_process_group()
    assert float(g1.abs().max()) == 0.0
    assert float((p1 - p0).abs().max()) <= 5e-4 * float(p0.abs().max())
    assert all(abs(a - b) <= 2e-2 * abs(b) + 1e-4 for a, b in zip(l1, l0))

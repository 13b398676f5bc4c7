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

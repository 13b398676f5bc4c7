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


def test_dropout_epoch_device_equals_host_emulation():
    """The device-resident dropout epoch (what makes graph replays draw fresh masks) must act on EVERY dropout site exactly
    like the same epoch folded into the keys on the host: forward activations bit-identical, gradients equal up to the
    fp32 atomics' summation order; and a different epoch must change the masks."""
    from egot2_b200 import _lib as L, specs, synth
    from egot2_b200.engine import TranslatorEngine, _stream
    from egot2_b200.hhi import PositionalEncoding
    dev = torch.device("cuda:0")
    results = {}
    for name, spec, seg, batch in (("hhi", specs.hhi_ttm_spec(128, 4, 1, 0.5, True), (30, 30, 30), 6),
                                   ("pnr", specs.hoi_pnr_spec(128, 2, 16, 0.5, 0.1), (16, 16, 8, 8), 5)):
        f = synth.make_features(spec, batch, seg, seed=4, dtype=torch.bfloat16)
        feats = [f[s.name].to(dev) for s in spec.segments]
        labels = synth.make_labels(spec, batch, seg, seed=4).to(dev)
        loss = L.LOSS_CE if name == "hhi" else L.LOSS_BCE_SIGMOID
        for mode in ("host5", "dev5", "dev6"):
            L.call("egot2_dropout_epoch_host", 5 if mode == "host5" else 0)
            L.call("egot2_dropout_epoch_enable", 0 if mode == "host5" else 1)
            if mode != "host5":
                L.call("egot2_dropout_epoch_set", int(mode[3:]), _stream())
            try:
                eng = TranslatorEngine(spec, dev, "bf16")
                eng.arena.load_state_dict(synth.make_state_dict(spec, 1))
                if spec.embed == "task_sinusoid":
                    eng.set_sinusoid(PositionalEncoding(spec.hidden).pe)
                act = eng.forward(feats, training=True, seed=21, labels=labels, loss=loss)
                eng.backward(act)
                torch.cuda.synchronize()
                results[(name, mode)] = ({k: act.t[k].float().clone() for k in ("x0", "hid0", "x_last", "out")},
                                         eng.arena.grad.clone())
            finally:
                L.call("egot2_dropout_epoch_host", 0)
                L.call("egot2_dropout_epoch_enable", 0)
        a_host, g_host = results[(name, "host5")]
        a_dev, g_dev = results[(name, "dev5")]
        a_other, _ = results[(name, "dev6")]
        for k in a_host:
            assert torch.equal(a_host[k], a_dev[k]), (name, k)
        assert float((g_host - g_dev).abs().max()) <= 1e-3 * float(g_host.abs().max()), name
        assert not torch.equal(a_dev["hid0"], a_other["hid0"]), name


def test_graph_replays_draw_fresh_dropout_masks(monkeypatch):
    from egot2_b200 import specs, synth
    from egot2_b200.trainer import TranslatorTrainer
    spec = specs.hhi_ttm_spec(128, 4, 1, 0.5, True)
    dev = torch.device("cuda:0")
    tr = TranslatorTrainer(spec, dev, "bf16", use_graphs=True, lr=0.0)          # lr 0: only the masks can change
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    seg = (30, 30, 30)
    f = synth.make_features(spec, 8, seg, seed=1, dtype=torch.bfloat16)
    feats = [f[s.name].to(dev) for s in spec.segments]
    labels = synth.make_labels(spec, 8, seg, seed=1).to(dev)
    hids = []
    for _ in range(3):
        tr.train_step(feats, labels, graph_key=0)
        torch.cuda.synchronize()
        act = tr._graphs[0][-1]
        hids.append(act.t["hid0"].clone())
    assert not torch.equal(hids[1], hids[2])
    z1, z2 = float((hids[1] == 0).float().mean()), float((hids[2] == 0).float().mean())
    assert abs(z1 - z2) < 0.02 and z1 > 0.6                                   # same keep rate, different pattern

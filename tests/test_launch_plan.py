"""Host-side launch plan of every parity case, checked WITHOUT a GPU: the C-ABI entry points are replaced by a recorder
(nothing is computed, buffers stay uninitialised), so what runs here is exactly the Python sequencing of the product
path - descriptor filling, buffer shapes, the order of `egot2_*` calls in forward, backward and greedy decoding - through
the public drop-in modules.  It catches host-logic errors (wrong segment geometry, a missing buffer, a mode that is not
wired) before a case reaches hardware; the arithmetic itself is covered by the `-m gpu` parity tests."""
import ctypes as C
import warnings

import pytest
import torch

from egot2_b200 import _lib as L
from egot2_b200 import engine as E
from egot2_b200 import modules as M
from oracle.cases import CASES, case_inputs

import test_modules as tm


@pytest.fixture
def recorder(monkeypatch):
    calls = []

    def fake_call(name, *args):
        calls.append((name, args))
        return 0
    monkeypatch.setattr(L, "call", fake_call)
    monkeypatch.setattr(E, "_stream", lambda: 0)
    monkeypatch.setattr(M, "_require_cuda", lambda device: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    return calls


def _names(calls):
    return [c[0] for c in calls]


def _build(name, dtype="fp32"):
    warnings.filterwarnings("ignore")
    case = CASES[name]
    sd, feats, labels, extra = case_inputs(case)
    m = tm.build_ours(case)
    m.load_state_dict(sd, strict=False)
    m.set_compute_dtype(dtype).eval()
    return case, m, feats, labels, extra


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if not CASES[n].raw_slowfast])
def test_forward_backward_plan(name, dtype, recorder):
    case, m, feats, labels, extra = _build(name, dtype)
    sp = case.spec
    out = tm.run_ours(case, m, feats, extra, torch.device("cpu"), labels)
    fwd = _names(recorder)
    assert fwd.count("egot2_embed_fwd") == 1
    layer = "egot2_vit_layer" if sp.encoder == "simple_vit" else "egot2_encoder_layer"
    assert fwd.count(layer + "_fwd") == sp.layers
    assert fwd.count("egot2_decoder_layer_fwd") == sp.decoder_layers
    if sp.embed == "task_sinusoid":
        assert fwd.index("egot2_hhi_tok_table_fwd") < fwd.index("egot2_embed_fwd")
    if sp.head == "decoder":
        assert fwd.index("egot2_prompt_embed_fwd") < fwd.index("egot2_decoder_layer_fwd") < fwd.index("egot2_head_loss_fwd")
        assert tuple(out.shape) == (labels.shape[0], sp.vocab, labels.shape[1] - 1)
    del recorder[:]
    out.float().sum().backward()
    bwd = _names(recorder)
    assert bwd.count(layer + "_bwd") == sp.layers
    assert bwd.count("egot2_decoder_layer_bwd") == sp.decoder_layers
    if sp.layers and layer == "egot2_encoder_layer" and m._engine.defer_joins:
        # deferred side-stream joins: scoped around the encoder layers + embedding stage, one join at the very end
        assert bwd[-3:] == ["egot2_embed_bwd", "egot2_side_defer", "egot2_side_join_all"]
        i0 = bwd.index("egot2_side_defer")
        assert i0 < bwd.index(layer + "_bwd") and bwd.count("egot2_side_defer") == 2
    else:
        assert bwd[-1] == "egot2_embed_bwd"
    for k in m._param_names:
        assert m.get_parameter(k).grad is not None, k


def _tok_table_runs(calls):
    (args,) = [a for n, a in calls if n == "egot2_hhi_tok_table_fwd"]
    n = args[3]
    return list(args[4][:n]), list(args[5][:n])


def test_hoi_g_action_tokens_share_one_position_run(recorder):
    """encode() :236-243: slow8 | fast8 are ONE task (row 2 of task_embed) with positions 0..15, while the projection
    stage still sees four feature streams."""
    case, m, feats, labels, extra = _build("hoi_g_h128_l2")
    tm.run_ours(case, m, feats, extra, torch.device("cpu"), labels)
    assert _tok_table_runs(recorder) == ([16, 16, 16], [0, 1, 2])
    (eargs,) = [a for n, a in recorder if n == "egot2_embed_fwd"]
    d = C.cast(eargs[0], C.POINTER(L.EmbedDesc)).contents
    assert d.n_seg == 4 and list(d.seg_tokens[:4]) == [16, 16, 8, 8] and list(d.seg_offset[:4]) == [0, 16, 32, 40]
    assert list(d.seg_in_dim[:4]) == [8192, 8192, 2048, 256] and d.T == 48


def test_hoi_g6_lta_mode_uses_four_task_rows(recorder):
    case, m, feats, labels, extra = _build("hoi_g6_lta_h128_l2")
    tm.run_ours(case, m, feats, extra, torch.device("cpu"), labels)
    assert _tok_table_runs(recorder) == ([2, 2, 2, 2], [0, 1, 2, 3])
    (eargs,) = [a for n, a in recorder if n == "egot2_embed_fwd"]
    d = C.cast(eargs[0], C.POINTER(L.EmbedDesc)).contents
    assert list(d.seg_has_proj[:4]) == [1, 1, 0, 1] and d.T == 8          # the action features arrive hidden-wide
    # both modes of the 6-task model live in ONE arena
    assert m._mode_engines["lta2"].arena is m._engine.arena


def test_hoi_g_greedy_decoding_encodes_once(recorder):
    """predict_ac (:264-275): the encoder runs once, the decoder + vocabulary head once per generated token with a
    prompt that grows by one."""
    case, m, feats, labels, extra = _build("hoi_g_h128_l2")
    vid, ac = [{"pnr": feats["pnr"], "oscc": feats["oscc"]}], {"slowfast": [feats["slow"], feats["fast"]]}
    toks = m.predict_ac(vid, ac)
    assert tuple(toks.shape) == (case.batch, 2) and toks.dtype == torch.int64
    names = _names(recorder)
    assert names.count("egot2_embed_fwd") == 1 and names.count("egot2_encoder_layer_fwd") == case.spec.layers
    assert names.count("egot2_prompt_embed_fwd") == 2 and names.count("egot2_head_loss_fwd") == 2
    assert names.count("egot2_decoder_layer_fwd") == 2 * case.spec.decoder_layers
    S = [a[2] for n, a in recorder if n == "egot2_prompt_embed_fwd"]
    assert S == [1, 2]
    # predict: one-token prompt; 'action_*' returns indices, 'pnr' / 'oscc' the vocabulary logits
    assert tuple(m.predict(vid, ac, "action_verb").shape) == (case.batch,)
    assert tuple(m.predict(vid, ac, "pnr").shape) == (case.batch, case.spec.vocab)


def test_hoi_g6_predict_returns_verb_noun_pairs(recorder):
    case, m, feats, labels, extra = _build("hoi_g6_clip_h128_l1")
    vid, ac = [{"pnr": feats["pnr"], "oscc": feats["oscc"]}], {"slowfast": [feats["slow"], feats["fast"]]}
    assert tuple(m.predict(vid, ac, "action").shape) == (case.batch, 2)
    assert tuple(m.predict(vid, ac, "oscc").shape) == (case.batch, case.spec.vocab)
    assert m.predict(vid, ac, "action", predict_verb_only=True) is None


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_hoi_g_training_step_plan(dtype, recorder):
    """Unified3TaskTranslation.training_step (HOI/tasks/multitask/video_task.py:182-204): three forward/backward passes
    into one gradient arena (cleared by the first only), one fused AdamW launch."""
    from egot2_b200 import synth
    from egot2_b200.trainer import HoiPromptTranslatorTrainer
    tr = HoiPromptTranslatorTrainer(hidden=128, heads=4, layers=2, vocab=40, device="cpu", dtype=dtype)
    sp = tr.spec
    tr.engine.arena.param.fill_(1.0)
    feats, labels = [], []
    for i, B in enumerate((2, 3, 4)):
        f = synth.make_features(sp, B, seed=50 + i)
        feats += [f[s.name] for s in sp.segments]
        labels.append(synth.make_labels(sp, B, seed=50 + i))
    tr.train_step(feats, torch.cat(labels))
    names = _names(recorder)
    assert names.count("egot2_embed_fwd") == 3 and names.count("egot2_embed_bwd") == 3
    assert names.count("egot2_encoder_layer_fwd") == 3 * sp.layers == names.count("egot2_encoder_layer_bwd")
    assert names.count("egot2_decoder_layer_fwd") == 3 * sp.decoder_layers == names.count("egot2_decoder_layer_bwd")
    assert names.count("egot2_adamw_step_fused") == 1 and names[-1] == "egot2_adamw_step_fused"
    rows = [a[1] for n, a in recorder if n == "egot2_prompt_embed_fwd"]
    assert rows == [2, 3, 4]
    (adam,) = [a for n, a in recorder if n == "egot2_adamw_step_fused"]
    assert adam[5] == pytest.approx(1e-4) and adam[9] == pytest.approx(1e-4)          # lr, decoupled weight decay (:265-268)
    assert float(tr.engine.arena.param[0]) == 1.0                                      # nothing touches the arena on the host
    assert tr._grad_clean
    assert (adam[12] is not None) == (dtype == "bf16")                                 # the launch also writes the bf16 shadow
    del recorder[:]
    tr.train_step(feats, torch.cat(labels))                                            # second step: shadow and gradients are current
    assert "egot2_cast_f32_to_bf16" not in _names(recorder)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n].spec.embed == "task_sinusoid"])
def test_table_runs_default_to_one_run_per_segment(name):
    case = CASES[name]
    sp = case.spec
    runs = sp.table_runs(case.seg_tokens)
    if any(s.pos_run for s in sp.segments):
        assert sum(r[0] for r in runs) == sum(case.seg_tokens) and len(runs) < len(sp.segments)
    else:
        assert runs == [(d, s.task_id) for s, d in zip(sp.segments, case.seg_tokens)]


def test_simple_vit_layers_replace_the_torch_encoder(recorder):
    """HOI PNR simple_vit sibling: the embed stage and the shared-ln head are those of the Dropout variant, the encoder is
    depth x egot2_vit_layer_fwd/bwd with dim_head independent of the model width."""
    case, m, feats, labels, extra = _build("hoi_pnr_vit_h256_l3")
    out = tm.run_ours(case, m, feats, extra, torch.device("cpu"), labels)
    names = _names(recorder)
    assert names.count("egot2_vit_layer_fwd") == 3 and "egot2_encoder_layer_fwd" not in names
    assert names.index("egot2_embed_fwd") < names.index("egot2_vit_layer_fwd") < names.index("egot2_head_loss_fwd")
    (a0, a1, a2) = [a for n, a in recorder if n == "egot2_vit_layer_fwd"]
    d = C.cast(a0[0], C.POINTER(L.VitDesc)).contents
    assert (d.B, d.T, d.D, d.heads, d.dim_head, d.mlp) == (case.batch, 48, 256, 8, 128, 512)
    assert a0[3] == a1[2] and a1[3] == a2[2]                      # layer i's output buffer is layer i+1's input
    del recorder[:]
    out.float().sum().backward()
    names = _names(recorder)
    assert names.count("egot2_vit_layer_bwd") == 3 and names[-1] == "egot2_embed_bwd"
    for k in m._param_names:
        assert m.get_parameter(k).grad is not None, k


@pytest.mark.parametrize("kind", ["hhi", "hoi"])
def test_prompt_trainer_graph_mode_plan(kind, recorder, monkeypatch):
    """EGOT2_G_GRAPH=1 (default off, not yet run on hardware): per graph key the three forward/backward passes are captured
    once and replayed, accumulating into an arena that the fused optimizer launch of the previous step left clean; the
    dropout epoch advances once per step.  The CUDA-graph capture itself is replaced by an eager stand-in here."""
    from egot2_b200 import specs, synth, trainer as T
    monkeypatch.setenv("EGOT2_G_GRAPH", "1")
    monkeypatch.setattr(T, "_cur_stream", lambda device: 0)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: None)
    captured = []

    def fake_capture(self, body):
        captured.append(1)
        return (lambda: body()), body()
    monkeypatch.setattr(T._PromptStepGraphs, "_capture_graph", fake_capture)
    if kind == "hoi":
        tr = T.HoiPromptTranslatorTrainer(hidden=128, heads=4, layers=1, vocab=40, device="cpu", dtype="bf16")
        sp = tr.spec
        feats, labels = [], []
        for i, B in enumerate((2, 3, 2)):
            f = synth.make_features(sp, B, seed=60 + i)
            feats += [f[s.name] for s in sp.segments]
            labels.append(synth.make_labels(sp, B, seed=60 + i))
        opt = "egot2_adamw_step_fused"
    else:
        tr = T.PromptTranslatorTrainer(hidden=128, heads=4, layers=1, device="cpu", dtype="bf16")
        feats, labels = [], []
        for mode, (B, D) in (("lam", (4, 7)), ("ttm", (2, 6)), ("asd", (2, 6))):
            sp = specs.hhi_g_spec(128, 4, 1, 0.1, mode)
            seg = (D,) if mode == "lam" else (D, D, D)
            f = synth.make_features(sp, B, seg, seed=70)
            feats += [f[s.name] for s in sp.segments]
            labels.append(synth.make_labels(sp, B, seg, seed=70))
        opt = "egot2_adam_step_fused"
    assert tr.use_graphs and tr.dropout_epoch
    assert _names(recorder)[:2] == ["egot2_dropout_epoch_enable", "egot2_dropout_epoch_set"]
    lab = torch.cat(labels)
    for step in range(3):
        del recorder[:]
        tr.train_step(feats, lab, graph_key=5)
        names = _names(recorder)
        assert names.count(opt) == 1 and names[-2:] == [opt, "egot2_dropout_epoch_advance"]
        assert tr._grad_clean
        if step > 0:                                  # replay only: exactly one forward/backward sequence, no re-cast
            assert names.count("egot2_embed_fwd") == 3 and "egot2_cast_f32_to_bf16" not in names
    assert len(captured) == 1
    tr.train_step(feats, lab, graph_key=6)
    assert len(captured) == 2
    # an eager step in between (bench.py counts launches that way) must not leave stale state behind
    tr.use_graphs = False
    tr.train_step(feats, lab, graph_key=5)
    tr.use_graphs = True
    if kind == "hhi":
        assert not tr._grad_clean                     # plain Adam keeps the gradients: the next replay clears them first


@pytest.mark.parametrize("wl", ["hhi_ttm3_train_b256", "hoi_pnr_train_b256", "hoi_lta_train_b512"])
def test_bench_workload_eager_step_plan(wl, recorder, monkeypatch):
    """The benchmark's trainer on its workload specs (small batch, eager launches): one fused Adam per step, the bf16 shadow
    is cast once and then kept current by the optimizer launch, `infer` runs the forward only."""
    import bench
    from egot2_b200 import synth, trainer as T
    monkeypatch.setenv("EGOT2_DROPOUT_EPOCH", "0")
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: None)
    w = bench.WORKLOADS[wl]
    spec, seg = w["spec"](), w["seg_tokens"]
    tr = T.TranslatorTrainer(spec, "cpu", "bf16", use_graphs=False)
    tr.load_state_dict(synth.make_state_dict(spec, 0))
    f = synth.make_features(spec, 4, seg, seed=1, dtype=torch.bfloat16)
    fe, la = [f[s.name] for s in spec.segments], synth.make_labels(spec, 4, seg, seed=1)
    tr.train_step(fe, la)
    del recorder[:]
    tr.train_step(fe, la)
    names = _names(recorder)
    assert names.count("egot2_adam_step_fused") == 1 and names[-1] == "egot2_adam_step_fused"
    assert names.count("egot2_encoder_layer_fwd") == spec.layers == names.count("egot2_encoder_layer_bwd")
    assert "egot2_cast_f32_to_bf16" not in names
    del recorder[:]
    out = tr.infer(fe)
    assert tuple(out.shape) == (4, spec.n_out) and not any(n.endswith("_bwd") for n in _names(recorder))


def test_pnr2_feature_dropout_mode_reaches_the_pnr_segment_only(recorder):
    """FEAT_DROPOUT_MODE > 0 (HOI/models/pnr/video_model_transfer.py:95-96): Dropout(FEAT_DROPOUT_RATE) on the projected PNR
    features alone = the first 16 tokens of every clip; mode 0 (shipped default) = no feature dropout at all."""
    from egot2_b200 import hoi
    from egot2_b200.modules import PrecomputedFeatures
    for mode, want in ((0, (0.0, 0)), (1, (0.5, 16)), (2, (0.5, 16))):
        cfg = tm.CfgNode(DATA=tm.CfgNode(TASK="keyframe_localization"),
                         MODEL=tm.CfgNode(FEAT_DROPOUT_RATE=0.5, FEAT_DROPOUT_MODE=mode, TRANSFORMER_DROPOUT_RATE=0.1))
        m = hoi.pnr.TaskFusionMFTransformerDropout(cfg, backbones={"pnr_model": PrecomputedFeatures("pnr"),
                                                                   "oscc_model": PrecomputedFeatures("oscc")})
        m.train()
        del recorder[:]
        m([{"pnr": torch.randn(2, 16, 8192), "oscc": torch.randn(2, 16, 8192)}])
        (eargs,) = [a for n, a in recorder if n == "egot2_embed_fwd"]
        d = C.cast(eargs[0], C.POINTER(L.EmbedDesc)).contents
        assert d.training == 1 and (round(d.p_feat, 6), d.feat_drop_tokens) == want

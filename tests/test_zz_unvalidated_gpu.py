"""GPU parity of cases that were added after the round's GPU budget was spent (oracle.cases.UNVALIDATED_ON_GPU).
Their CPU-side checks are green (state_dict keys, same-seed init, oracle pinned against the reference class, golden);
the hardware run is pending, so these are xfail(strict=False): a pass shows up as XPASS, a failure does not turn the suite
red.  Move a name out of UNVALIDATED_ON_GPU once it has passed on a B200 and it joins the regular parametrisations."""
import pytest

from oracle.cases import UNVALIDATED_ON_GPU

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="not yet run on hardware (GPU budget of round 1 spent)")]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(UNVALIDATED_ON_GPU))
def test_engine_parity_unvalidated(name, dtype):
    import test_gpu_parity as tg          # the tests directory is on sys.path (rootdir conftest / prepend import mode)
    tg.test_engine_matches_oracle_and_golden(name, dtype)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(UNVALIDATED_ON_GPU))
def test_module_parity_unvalidated(name, dtype):
    import test_modules as tm
    tm.test_module_forward_backward_vs_oracle(name, dtype)


def test_hoi_g_greedy_predict_ac_matches_reference_golden():
    """HOI EgoT2-g greedy decoding (predict_ac, HOI/models/multitask/video_model_builder.py:264-275) through the drop-in
    module: the generated [verb, noun] vocabulary indices equal the real reference class's (tests/golden/next_hoi_g.npz),
    and decoding over the kept encoder memory equals a full forward with the same prompt."""
    import warnings
    from types import SimpleNamespace

    import numpy as np
    import torch

    from egot2_b200 import hoi
    from egot2_b200.modules import PrecomputedFeatures
    from oracle import next_rows as NR
    warnings.filterwarnings("ignore")
    dev = torch.device("cuda:0")
    sd, feats, target = NR.inputs()
    gold = np.load(NR.GOLDEN)
    args = SimpleNamespace(hidden_dim=NR.H, num_heads=NR.HEADS, num_layers=NR.LAYERS, dropout=0.1)
    vocab = {("action" if i == 4 else f"w{i}"): i for i in range(NR.VOCAB)}
    bb = {"pnr_model": PrecomputedFeatures("pnr"), "oscc_model": PrecomputedFeatures("oscc"),
          "recognition_model": PrecomputedFeatures("slowfast")}
    m = hoi.multitask.TaskTranslationPromptTransformer(args, vocab, backbones=bb)
    m.load_state_dict(sd, strict=False)
    m.to(dev).set_compute_dtype("fp32").eval()
    f = {k: v.to(dev) for k, v in feats.items()}
    vid, ac = [{"pnr": f["pnr"], "oscc": f["oscc"]}], {"slowfast": [f["slow"], f["fast"]]}
    out = m(vid, ac, target[:, :-1].to(dev)).float().cpu()
    ref = torch.from_numpy(gold["output"])
    assert float((out - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    toks = m.predict_ac(vid, ac).cpu()
    assert torch.equal(toks, torch.from_numpy(gold["predict_ac"]))
    # the decoder-only second step == a full forward on the two-token prompt
    start = torch.full((NR.B, 1), 4, dtype=torch.int64)
    full = m(vid, ac, torch.cat([start, toks[:, :1]], dim=1).to(dev))          # (B, V, 2)
    assert torch.equal(full[:, :, -1].argmax(dim=1).cpu(), toks[:, 1])


def test_fused_adamw_matches_torch():
    """egot2_adamw_step_fused == torch.optim.AdamW (decoupled decay) over a flat arena, incl. the bf16 shadow it writes and
    the gradient clear."""
    import torch

    from egot2_b200 import _lib as L
    torch.manual_seed(0)
    n = 4096 + 37
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for step in range(1, 4):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        L.call("egot2_adamw_step_fused", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-2, 0.9, 0.999, 1e-8, 0.1,
               step, 1.0, shadow.data_ptr(), 1, st)
        assert float((p - ref.detach()).abs().max()) < 1e-6
        assert float(g.abs().max()) == 0.0
        assert torch.equal(shadow, p.to(torch.bfloat16))


@pytest.mark.parametrize("kind", ["hhi", "hoi"])
def test_prompt_trainer_graph_mode_trains(kind, monkeypatch):
    """EGOT2_G_GRAPH=1: the replayed EgoT2-g step optimises like the eager one - on one fixed batch the loss falls from its
    initial value within a few steps in both modes and ends up in the same range (the two modes draw different dropout
    masks by construction, so the comparison is statistical)."""
    import torch

    from egot2_b200 import specs, synth, trainer as T

    def batch():
        feats, labels = [], []
        if kind == "hoi":
            sp = specs.hoi_g_spec(128, 4, 1, 0.1, 40)
            for i, B in enumerate((8, 8, 8)):
                f = synth.make_features(sp, B, seed=60 + i, dtype=torch.bfloat16)
                feats += [f[s.name].cuda() for s in sp.segments]
                labels.append(synth.make_labels(sp, B, seed=60 + i))
        else:
            for mode, (B, D) in (("lam", (16, 7)), ("ttm", (4, 10)), ("asd", (4, 10))):
                sp = specs.hhi_g_spec(128, 4, 1, 0.1, mode)
                seg = (D,) if mode == "lam" else (D, D, D)
                f = synth.make_features(sp, B, seg, seed=70, dtype=torch.bfloat16)
                feats += [f[s.name].cuda() for s in sp.segments]
                labels.append(synth.make_labels(sp, B, seg, seed=70))
        return feats, torch.cat(labels).cuda()

    def run(graph):
        monkeypatch.setenv("EGOT2_G_GRAPH", "1" if graph else "0")
        if kind == "hoi":
            tr = T.HoiPromptTranslatorTrainer(hidden=128, heads=4, layers=1, vocab=40, device="cuda:0", dtype="bf16", lr=2e-3)
            sd = synth.make_state_dict(tr.spec, 3)
        else:
            tr = T.PromptTranslatorTrainer(hidden=128, heads=4, layers=1, device="cuda:0", dtype="bf16", lr=2e-3)
            sd = synth.make_state_dict(tr.spec, 3)
        tr.load_state_dict(sd)
        assert tr.use_graphs == graph
        feats, lab = batch()
        return [float(tr.train_step(feats, lab, graph_key=0)) for _ in range(12)]

    eager, graph = run(False), run(True)
    for losses in (eager, graph):
        assert all(l == l and abs(l) < 1e4 for l in losses)
        assert min(losses[-3:]) < 0.8 * losses[0]
    assert abs(graph[0] - eager[0]) < 0.25 * eager[0]
    assert abs(sum(graph[-3:]) - sum(eager[-3:])) < 0.5 * sum(eager[-3:])


def test_pnr2_feature_dropout_hits_only_the_pnr_tokens():
    """egot2_embed_desc.feat_drop_tokens: in train mode about p of the projected PNR features (tokens 0..15 of every clip)
    are zero in the saved LayerNorm input, none of the OSCC ones; the same seed reproduces the mask."""
    import torch

    from egot2_b200 import specs, synth
    from egot2_b200.engine import TranslatorEngine
    sp = specs.hoi_pnr2_spec(16, 0.1, 0.5, 1)
    eng = TranslatorEngine(sp, "cuda:0", "fp32")
    eng.arena.load_state_dict(synth.make_state_dict(sp, 5))
    f = synth.make_features(sp, 8, seed=5)
    feats = [f[s.name].cuda() for s in sp.segments]
    z1 = eng.forward(feats, training=True, seed=11).t["z"].clone()
    z2 = eng.forward(feats, training=True, seed=11).t["z"].clone()
    z3 = eng.forward(feats, training=True, seed=12).t["z"].clone()
    assert torch.equal(z1, z2) and not torch.equal(z1, z3)
    frac_pnr = float((z1[:, :16] == 0).float().mean())
    assert abs(frac_pnr - 0.5) < 0.03, frac_pnr
    assert float((z1[:, 16:] == 0).float().mean()) == 0.0
    z_eval = eng.forward(feats, training=False).t["z"]
    kept = z1[:, :16] != 0
    assert torch.allclose(z1[:, :16][kept], 2.0 * z_eval[:, :16][kept], rtol=1e-5, atol=1e-6)       # kept values scaled by 1/(1-p)
    assert torch.equal(z1[:, 16:], z_eval[:, 16:])

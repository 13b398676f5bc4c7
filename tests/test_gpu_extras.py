"""GPU tests beside the per-case parity files: greedy decoding of HOI EgoT2-g, the fused AdamW launch, the EgoT2-g trainers'
CUDA-graph mode, the PNR-only feature dropout, parity at the BENCHMARK shape (256 clips x 90 tokens: FF-split tail tiles,
full-size grids), the feature gradients of trainable backbones, and N-rank == 1-rank over NCCL.  All of them ran green on
a B200 in round 2; none is xfail."""
import pytest

pytestmark = pytest.mark.gpu


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
    # ... and the mask is exactly the documented generator (csrc/common.cuh drop_keep): at p == 0.5 every elementwise site
    # takes ONE random bit per element - bit idx % 32 of the hash of idx / 32 - with idx the flat (clip, token, column) index
    import test_gpu_parity as tg
    B, T, H = z1.shape
    want = tg.dropout_mask_oracle(11, 1, 0, B * T * H, 0.5).reshape(B, T, H)[:, :16] != 0          # SITE_FEAT = 1
    live = z_eval[:, :16] != 0                      # an exactly-zero projection says nothing about its mask bit
    assert torch.equal(kept.cpu()[live.cpu()], want[live.cpu()])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_benchmark_shape_matches_oracle(dtype):
    """Parity AT the shape bench.py times (BASELINE config 2: HHI 3-task, 256 clips x 3 x 30 frames, H 128, FF 2048): the
    kernels then run their full-size grids - FF-split tail tiles of the fused FFN, one-wave split-K weight gradients,
    multi-tile persistent GEMMs - which the 2-8 clip cases never reach.  Output, loss and every gradient vs the CPU
    oracle on the same seeded inputs (eval mode: dropout masks cannot be reproduced by torch)."""
    import torch

    import test_gpu_parity as tg
    from egot2_b200 import _lib as L
    from egot2_b200 import specs
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    from oracle.cases import Case, case_inputs, oracle_forward_loss
    torch.set_num_threads(max(1, (__import__("os").cpu_count() or 2)))
    case = Case("bench_hhi_ttm3_b256", specs.hhi_ttm_spec(128, 4, 1, 0.5, True), 256, (30, 30, 30), 31)
    sp, tol = case.spec, tg.TOL[dtype]
    sd, feats, labels, extra = case_inputs(case)
    eng = TranslatorEngine(sp, "cuda:0", dtype)
    eng.arena.load_state_dict(sd)
    eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
    gf = tg._engine_feats(case, eng, feats, extra, dtype)
    act = eng.forward(gf, training=False, labels=labels, loss=L.LOSS_CE, class_weight=torch.tensor([0.266, 0.734]))
    grad, _ = eng.backward(act)
    out = act.t["out"].float().cpu()
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_out, o_loss = oracle_forward_loss(case, P, feats, labels, extra)
    scale = float(o_out.abs().max())
    assert float((out - o_out.detach()).abs().max()) <= tol["out"] * scale
    assert abs(float(act.t["loss"][0]) - float(o_loss)) <= tol["loss"] * abs(float(o_loss))
    if dtype == "fp32":
        # argmax must agree wherever the reference's own top-1 margin exceeds the fp32 tolerance
        margin = (o_out[:, 0] - o_out[:, 1]).abs().detach()
        ok = margin > 4 * tol["out"] * scale
        assert torch.equal(out.argmax(-1)[ok], o_out.argmax(-1)[ok]) and int(ok.sum()) >= 250
    names = list(sd.keys())
    o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
    for k, g_ref in zip(names, o_grads):
        g = eng.arena.view(k, grad).float().cpu()
        err_l2 = float((g - g_ref).norm()) / (float(g_ref.norm()) + 1e-12)
        bound = 3 * tol["grad_l2"] if dtype == "fp32" else tg.bf16_grad_bound(case, k, tol)
        assert err_l2 <= bound, f"{k}: rel L2 err {err_l2:.3e} > {bound:.3e}"


@pytest.mark.parametrize("name", ["hhi3_h128_d30", "hoi_pnr_h128_l6", "hoi_ar_h128_l3"])
def test_layernorm_in_gemm_epilogue_matches_separate_kernel(name, monkeypatch):
    """EGOT2_GEMM_LN=1 (opt-in, see gemm_sm100_ln_ok) normalises the rows of the embedding projections and of the attention
    out-projection inside the tcgen05 GEMM's epilogue instead of in a LayerNorm kernel of its own.  Both read the same
    bf16-rounded pre-norm rows, so output, loss and gradients must agree to bf16 rounding of the normalised rows."""
    import torch

    import test_gpu_parity as tg
    from egot2_b200 import _lib as L
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    from oracle.cases import CASES, case_inputs
    case = CASES[name]
    sd, feats, labels, extra = case_inputs(case)
    res = {}
    loss_kind, cw = tg._loss_kind(case)
    for mode in ("0", "1"):
        monkeypatch.setenv("EGOT2_GEMM_LN", mode)
        eng = TranslatorEngine(case.spec, "cuda:0", "bf16")
        eng.arena.load_state_dict(sd)
        eng.set_sinusoid(O.sinusoid_table(1000, case.spec.hidden))
        gf = tg._engine_feats(case, eng, feats, extra, "bf16")
        L.prof_enable(True)
        try:
            act = eng.forward(gf, training=False, labels=labels, loss=loss_kind, class_weight=cw)
            torch.cuda.synchronize()
            n_ln = sum(1 for r in L.prof_report() if r[0].startswith("ln_fwd"))
        finally:
            L.prof_enable(False)
        grad, _ = eng.backward(act)
        torch.cuda.synchronize()
        res[mode] = (act.t["out"].float().cpu(), float(act.t["loss"][0]), grad.float().cpu().clone(), n_ln)
    (o0, l0, g0, n0), (o1, l1, g1, n1) = res["0"], res["1"]
    assert n1 < n0, f"the fused path did not replace any LayerNorm launch ({n0} -> {n1})"
    scale = float(o0.abs().max())
    assert float((o0 - o1).abs().max()) <= 2e-2 * scale
    assert abs(l0 - l1) <= 2e-2 * abs(l0)
    assert float((g0 - g1).norm()) <= 5e-2 * float(g0.norm())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name,which", [("hhi3_h128_l1", "ttm"), ("hoi_lta_h512_l4", "action"), ("hoi_lta_h512_l4", "lta")])
def test_feature_gradients_of_trainable_backbones(name, which, dtype):
    """dX for trainable inputs (SURVEY Appendix A 'backward obligations'): HHI --nofreeze (the TTM backbone trains,
    HHI/models/ttm/model_taskspecific.py:30-32) and the SlowFast head behind LTA's pass-through `feat_action` tokens
    (HOI/utils/multitask/load_model.py:105-110), plus a projected LTA stream: egot2_embed_grads.dfeat vs autograd."""
    import torch

    import test_gpu_parity as tg
    from egot2_b200.engine import TranslatorEngine
    from oracle import translator_oracle as O
    from oracle.cases import CASES, case_inputs, oracle_forward_loss
    case = CASES[name]
    sp = case.spec
    sd, feats, labels, extra = case_inputs(case)
    eng = TranslatorEngine(sp, "cuda:0", dtype)
    eng.arena.load_state_dict(sd)
    if sp.embed == "task_sinusoid":
        eng.set_sinusoid(O.sinusoid_table(1000, sp.hidden))
    loss_kind, cw = tg._loss_kind(case)
    gf = tg._engine_feats(case, eng, feats, extra, dtype)
    act = eng.forward(gf, training=False, labels=labels, loss=loss_kind, class_weight=cw)
    want = [s.name == which for s in sp.segments]
    _, dfeats = eng.backward(act, want_dfeat=want)
    fr = {k: v.clone().requires_grad_(k == which) for k, v in feats.items()}
    _, o_loss = oracle_forward_loss(case, {k: v.clone() for k, v in sd.items()}, fr, labels, extra)
    (g_ref,) = torch.autograd.grad(o_loss, [fr[which]])
    g = dfeats[[s.name for s in sp.segments].index(which)].float().cpu()
    assert g.shape == g_ref.shape
    err = float((g - g_ref).norm()) / (float(g_ref.norm()) + 1e-12)
    assert err <= (2e-3 if dtype == "fp32" else 4e-2), f"d(feat {which}): rel L2 err {err:.3e}"
    assert all(d is None for d, w in zip(dfeats, want) if not w)


def _nccl_rank_worker(rank, world, port, q, fused="1", graphs=False, steps=1):
    import os
    os.environ["EGOT2_DP_FUSED"] = fused

    import torch
    import torch.distributed as dist

    from egot2_b200 import specs, synth
    from egot2_b200.trainer import TranslatorTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sp = specs.hhi_ttm_spec(128, 4, 1, 0.0, True)               # dropout off: ranks and the 1-rank run see the same function
        sp = __import__("dataclasses").replace(sp, p_embed=0.0)
        B, seg = 16, (10, 10, 10)
        f = synth.make_features(sp, B, seg, seed=77, dtype=torch.bfloat16)
        lab = synth.make_labels(sp, B, seg, seed=77)
        lo, hi = rank * B // world, (rank + 1) * B // world
        tr = TranslatorTrainer(sp, f"cuda:{rank}", "bf16", use_graphs=graphs)
        assert (tr.peer is not None) == (fused == "1"), "the exchange path is not the one asked for"
        tr.load_state_dict(synth.make_state_dict(sp, 9))
        feats = [f[s.name][lo:hi].cuda() for s in sp.segments]
        labels = lab[lo:hi].cuda()
        for _ in range(steps):
            tr.train_step(feats, labels, graph_key=0 if graphs else None)
        torch.cuda.synchronize()
        q.put((rank, {k: v.cpu() for k, v in tr.state_dict().items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_two_ranks(fused, graphs, steps):
    import socket

    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_rank_worker, args=(r, 2, port, q, fused, graphs, steps)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=170) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got[0], got[1]


def test_two_rank_fused_exchange_multi_step_graphs():
    """The peer-memory exchange kernel (csrc/peer.cu) inside the whole-step CUDA graph: after 4 replayed steps the two ranks
    hold BIT-IDENTICAL parameters (every slice is computed once, by its owner, and written to both arenas), and they agree
    with 4 steps of the NCCL all-reduce + fused-Adam path to bf16-training accuracy."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    a0, a1 = _run_two_ranks("1", True, 4)
    for k in a0:
        assert torch.equal(a0[k], a1[k]), f"{k}: ranks diverged"
    n0, _ = _run_two_ranks("0", True, 4)
    for k in a0:
        assert float((a0[k] - n0[k]).abs().max()) <= 2e-4 + 1e-3 * float(n0[k].abs().max()), k


@pytest.mark.parametrize("fused", ["1", "0"])
def test_two_rank_nccl_step_equals_one_rank_step(fused):
    """SURVEY 4(5) on hardware: one optimisation step on 16 clips sharded over 2 ranks (NCCL all-reduce of the flat
    gradient arena, DDP-mean) leaves the same parameters as the same step on the concatenated batch on one rank.
    CE with class weights normalises by each shard's own weight sum (exactly what DDP does with the reference's loss), so
    the 1-rank comparison uses the mean of the two half-batch losses."""
    import socket

    import torch
    import torch.multiprocessing as mp

    from egot2_b200 import specs, synth
    from egot2_b200.engine import TranslatorEngine
    from egot2_b200 import _lib as L
    from egot2_b200.hhi import PositionalEncoding
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    sd2, sd2_r1 = _run_two_ranks(fused, False, 1)      # fused = "1": the peer-memory kernel; "0": NCCL all-reduce + fused Adam
    if fused == "1":
        for k in sd2:
            assert torch.equal(sd2[k], sd2_r1[k]), f"{k}: ranks diverged"
    # 1-rank reference: the two half-batch gradients averaged = what the 2-rank all-reduce (mean) produced
    sp = specs.hhi_ttm_spec(128, 4, 1, 0.0, True)
    sp = __import__("dataclasses").replace(sp, p_embed=0.0)
    B, seg = 16, (10, 10, 10)
    f = synth.make_features(sp, B, seg, seed=77, dtype=torch.bfloat16)
    lab = synth.make_labels(sp, B, seg, seed=77)
    eng = TranslatorEngine(sp, "cuda:0", "bf16")
    eng.arena.load_state_dict(synth.make_state_dict(sp, 9))
    eng.set_sinusoid(PositionalEncoding(sp.hidden).pe)
    cw = torch.tensor([0.266, 0.734])
    gsum = torch.zeros_like(eng.arena.grad)
    for lo, hi in ((0, 8), (8, 16)):
        act = eng.forward([f[s.name][lo:hi].cuda() for s in sp.segments], training=True, seed=1, labels=lab[lo:hi].cuda(),
                          loss=L.LOSS_CE, class_weight=cw)
        g, _ = eng.backward(act)
        gsum += g
    eng.arena.grad.copy_(gsum)
    eng.adam_step({}, 1, grad_scale=0.5)
    torch.cuda.synchronize()
    for k, v in eng.arena.state_dict().items():
        ref = v.cpu()
        assert float((sd2[k] - ref).abs().max()) <= 1e-6 + 1e-5 * float(ref.abs().max()), k

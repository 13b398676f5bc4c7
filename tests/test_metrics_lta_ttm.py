"""SURVEY 8f-3, second half: LTA top-k / edit distance / AUED and the TTM per-segment scores.
CPU: the restatement (oracle/metrics_oracle.py) against committed goldens made from the reference functions
(oracle/make_golden_metrics.py) and, where the reference exists, against those functions live.
GPU: the on-device drop-ins (egot2_b200/lta_metrics.py, egot2_b200/ttm_postprocess.py -> libegot2.so) against the restatement:
integer results bit-exact, float64 averages to 1e-12, fp32 softmax scores to 1e-6."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import make_golden_metrics as G
from oracle import metrics_oracle as MO

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "metrics_lta_ttm.json")))


def test_oracle_matches_golden():
    for rec in GOLD["topk"]:
        p, l = G.topk_inputs(*rec["case"])
        assert MO.topks_correct(p, l, [1, 5]) == rec["correct"]
    for rec in GOLD["ed"]:
        p, l = G.ed_inputs(*rec["case"])
        assert abs(MO.edit_distance(p.numpy(), l.numpy()) - rec["edit_distance"]) <= 1e-12
        au = MO.aued(p.numpy(), l.numpy())
        assert set(au) == set(rec["aued"])
        for k, v in rec["aued"].items():
            assert np.allclose(np.atleast_1d(au[k]), v, rtol=0, atol=1e-12), k
    for rec in GOLD["ttm"]:
        gt, pred = MO.ttm_segment_scores(G.ttm_inputs(*rec["case"]))
        assert gt == rec["groundtruth"]
        assert [r[:5] for r in pred] == [r[:5] for r in rec["prediction"]]
        assert np.allclose([r[5] for r in pred], [r[5] for r in rec["prediction"]], rtol=0, atol=1e-7)


def test_levenshtein_known_answers():
    # the classic pairs (kitten/sitting = 3, flaw/lawn = 2), empty and identical sequences, a transposition counts 2
    enc = lambda s: [ord(c) for c in s]
    assert MO.levenshtein(enc("kitten"), enc("sitting")) == 3
    assert MO.levenshtein(enc("flaw"), enc("lawn")) == 2
    assert MO.levenshtein([], [1, 2, 3]) == 3 and MO.levenshtein([1, 2, 3], [1, 2, 3]) == 0
    assert MO.levenshtein([1, 2], [2, 1]) == 2


@pytest.mark.requires_reference
def test_oracle_matches_reference_live():
    ref = G.reference_outputs()
    assert ref == GOLD or json.loads(json.dumps(ref)) == GOLD      # the committed goldens ARE what the reference computes here


def _cuda(t):
    return t.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.TOPK_CASES + [(4096, 593, 11), (7, 3, 12)])
def test_topk_on_device(case):
    from egot2_b200 import lta_metrics as M
    p, l = G.topk_inputs(*case)
    ks = [1, 5] if case[1] >= 5 else [1, 2]
    got = M.topks_correct(_cuda(p), _cuda(l), ks)
    assert [float(x) for x in got] == MO.topks_correct(p, l, ks)
    errs = M.topk_errors(_cuda(p), _cuda(l), ks)
    assert [float(e) for e in errs] == [float((1.0 - torch.tensor(c) / p.size(0)) * 100.0) for c in MO.topks_correct(p, l, ks)]


@pytest.mark.gpu
def test_topk_ties_and_bad_labels():
    from egot2_b200 import lta_metrics as M
    p = torch.zeros(6, 10)                      # all equal: the label's rank is its index (ties go to the smaller index)
    l = torch.tensor([0, 1, 4, 5, 9, 3])
    got = [float(x) for x in M.topks_correct(p.cuda(), l.cuda(), [1, 5])]
    assert got == [1.0, 4.0]
    assert [float(x) for x in M.topks_correct(torch.randn(0, 10).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), [1, 5])] == [0.0, 0.0]


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.ED_CASES + [(513, 20, 5, 21), (5, 64, 2, 22)])
def test_edit_distance_on_device(case):
    from egot2_b200 import _lib as L, lta_metrics as M
    from egot2_b200.engine import _stream
    p, l = G.ed_inputs(*case)
    N, Z, K = p.shape
    # every prefix distance, per clip, bit-exact
    md = torch.empty((N, Z), device="cuda", dtype=torch.int32)
    sums = torch.empty(Z, device="cuda", dtype=torch.int64)
    pc, lc = p.cuda().contiguous(), l.cuda().contiguous()
    L.call("egot2_edit_distance_prefix", N, Z, K, pc.data_ptr(), lc.data_ptr(), md.data_ptr(), sums.data_ptr(), _stream())
    ref = torch.tensor([[min(MO.levenshtein(p[n, :z, k], l[n, :z]) for k in range(K)) for z in range(1, Z + 1)] for n in range(N)])
    assert torch.equal(md.cpu().long(), ref)
    assert torch.equal(sums.cpu(), ref.sum(0))
    assert abs(M.edit_distance(pc, lc) - MO.edit_distance(p.numpy(), l.numpy())) <= 1e-12
    if Z > 1:
        a, b = M.AUED(pc, lc.unsqueeze(-1)), MO.aued(p.numpy(), l.numpy())
        assert set(a) == set(b)
        for k in a:
            assert a[k].shape == np.atleast_1d(b[k]).shape and np.allclose(a[k], b[k], rtol=0, atol=1e-12), k


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.TTM_CASES + [(300, 31)])
def test_ttm_postprocessor_on_device(case, tmp_path):
    from egot2_b200.ttm_postprocess import PostProcessor
    batches = G.ttm_inputs(*case)
    pp = PostProcessor(SimpleNamespace(exp_path=str(tmp_path), rank=0))
    for out, tg in batches:
        pp.update(out.cuda(), tg)
    gt, pred = pp.results()
    o_gt, o_pred = MO.ttm_segment_scores(batches)
    assert gt == o_gt
    assert [r[:5] for r in pred] == [r[:5] for r in o_pred]
    assert np.allclose([r[5] for r in pred], [r[5] for r in o_pred], rtol=0, atol=1e-6)
    pp.save()
    import pandas as pd
    df = pd.read_csv(pp.predctionfile, header=None)
    assert len(df) == len(o_pred) and np.allclose(df[5].to_numpy(), [r[5] for r in o_pred], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 20, 593), (1, 2), (257, 7), (0, 5)])
def test_device_softmax(shape):
    from egot2_b200.modules import device_softmax
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g) * 4
    got = device_softmax(x.cuda()).cpu()
    ref = torch.softmax(x.double(), dim=-1)
    assert got.shape == x.shape
    if x.numel():
        assert float((got.double() - ref).abs().max()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_device_mean_dim1(dtype):
    from egot2_b200.modules import device_mean_dim1
    g = torch.Generator().manual_seed(6)
    x = torch.randn(5, 16, 8192, generator=g).to(dtype)
    got = device_mean_dim1(x.cuda()).cpu()
    ref = x.double().mean(dim=1)
    assert got.dtype == dtype and got.shape == (5, 8192)
    assert float((got.double() - ref).abs().max()) <= (1e-6 if dtype == torch.float32 else 4e-3)

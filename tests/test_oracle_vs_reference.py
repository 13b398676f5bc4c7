"""Live check (build container only): restatement == verbatim reference classes on a fresh random draw that is NOT in
the committed goldens - output, loss AND the gradient of every translator parameter, for every case of oracle/cases.py
(encoder-only translators, both EgoT2-g families, the simple_vit siblings)."""
import warnings

import pytest
import torch

from oracle.cases import CASES, Case, case_inputs, oracle_forward_loss

pytestmark = pytest.mark.requires_reference


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shims as rs
    warnings.filterwarnings("ignore")
    torch.backends.mha.set_fastpath_enabled(False)
    return rs.load_hhi(), rs.load_hoi()


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_equals_reference_fresh_seed(ref, name):
    from oracle.make_golden import build_reference, reference_forward_loss
    hhi, hoi = ref
    base = CASES[name]
    case = Case(base.name, base.spec, base.batch + 1, base.seg_tokens, seed=1234, raw_slowfast=base.raw_slowfast)
    sd, feats, labels, extra = case_inputs(case)
    m = build_reference(case, hhi, hoi)
    m.load_state_dict(sd, strict=False)
    m.eval()
    out, loss = reference_forward_loss(case, m, hhi, feats, labels, extra)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_out, o_loss = oracle_forward_loss(case, P, feats, labels, extra)
    torch.testing.assert_close(o_out.detach(), out.detach().reshape(o_out.shape), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(o_loss.detach(), loss.detach(), atol=2e-5, rtol=1e-4)
    # gradients of every translator parameter: autograd through the reference module vs autograd through the restatement
    params = dict(m.named_parameters())
    names = [k for k in sd if k in params]
    g_ref = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    g_o = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
    assert len(names) >= 10
    for k, a, b in zip(names, g_ref, g_o):
        if a is None:
            assert b is None or float(b.abs().max()) == 0.0, k
            continue
        assert b is not None, k
        err = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-12)
        assert err < 2e-4, (k, err)

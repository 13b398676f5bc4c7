"""Live check (build container only): restatement == verbatim reference classes, including a
fresh random draw that is NOT in the committed goldens."""
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


@pytest.mark.parametrize("name", ["hhi2_h128_l1", "hhi3_h128_l1", "hhi_asd_h128_l1", "hoi_pnr_h128_l6", "hoi_lta_h512_l4", "hoi_pnr2_h256_l3", "hoi_ar_h128_l3", "hoi_ar2_h128_l2", "hoi_lta2_h512_l1"])
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
    o_out, o_loss = oracle_forward_loss(case, sd, feats, labels, extra)
    torch.testing.assert_close(o_out, out.detach(), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(o_loss, loss.detach(), atol=2e-5, rtol=1e-4)

"""The CPU oracle (oracle/translator_oracle.py) must reproduce the golden vectors that were
produced by the REAL reference classes (oracle/make_golden.py).  Runs anywhere (no GPU, no
/root/reference): inputs/weights are regenerated from seeds, outputs come from tests/golden."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import CASES, case_inputs, grad_digest, oracle_forward_loss

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    case = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd, feats, labels, extra = case_inputs(case)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out, loss = oracle_forward_loss(case, P, feats, labels, extra)
    ref_out = torch.from_numpy(gold["output"])
    # fp32 tolerance: 1e-3 of the output range (north_star), we are far inside it
    scale = float(ref_out.abs().max())
    assert float((out.detach() - ref_out).abs().max()) <= 1e-4 * scale + 1e-6
    assert abs(float(loss) - float(gold["loss"])) <= 1e-4 * abs(float(gold["loss"])) + 1e-6
    if ref_out.dim() == 2 and ref_out.shape[1] in (2, 16):
        assert torch.equal(out.argmax(-1), ref_out.argmax(-1))        # keyframe / class index bit-exact
    names = [k[len("grad/"):] for k in gold.files if k.startswith("grad/")]
    grads = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    for k, g in zip(names, grads):
        ref_d = torch.from_numpy(gold["grad/" + k])
        d = grad_digest(g if g is not None else torch.zeros_like(P[k]))
        tol = 2e-4 * float(ref_d[2]) + 1e-7                              # relative to the grad's absmax
        assert float((d[3:] - ref_d[3:]).abs().max()) <= tol, k
        assert abs(float(d[1] - ref_d[1])) <= 2e-4 * float(ref_d[1]) + 1e-7, k


def test_sinusoid_table_known_values():
    from oracle.translator_oracle import sinusoid_table
    pe = sinusoid_table(4, 8)
    assert pe[0].tolist() == [0, 1, 0, 1, 0, 1, 0, 1]
    assert abs(float(pe[1, 0]) - np.sin(1.0)) < 1e-6 and abs(float(pe[1, 1]) - np.cos(1.0)) < 1e-6
    assert abs(float(pe[3, 2]) - np.sin(3 * 10000 ** (-2 / 8))) < 1e-6

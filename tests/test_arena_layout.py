"""Host-side invariants of the flat parameter arena (no GPU): every reference state_dict key has a slot, slots are
aligned and disjoint, the embedding-stage parameters form the prefix [0, embed_numel) (what the data-parallel trainer
all-reduces last), and stacked heads are contiguous (one GEMM operand)."""
import pytest
import torch

from egot2_b200 import specs
from egot2_b200.engine import ParamArena, _ALIGN

SPECS = {
    "hhi2": specs.hhi_ttm_spec(128, 4, 1, 0.5, False),
    "hhi3": specs.hhi_ttm_spec(128, 4, 1, 0.5, True),
    "hhi_asd": specs.hhi_asd_spec(128, 4, 1, 0.5),
    "hhi_g": specs.hhi_g_spec(256, 4, 3, 0.1, "ttm"),
    "hoi_g": specs.hoi_g_spec(256, 4, 3, 0.1, 600),
    "hoi_g6": specs.hoi_g_spec(256, 4, 3, 0.1, 600, "clip", 4),
    "hoi_g6_lta": specs.hoi_g_spec(256, 4, 3, 0.1, 600, "lta", 4),
    "pnr": specs.hoi_pnr_spec(128, 6, 16, 0.5, 0.1),
    "pnr2": specs.hoi_pnr2_spec(16, 0.1),
    "ar": specs.hoi_ar_spec(128, 3, 8, 0.1),
    "ar2": specs.hoi_ar2_spec(128, 2, 8, 0.1),
    "lta": specs.hoi_lta_spec(512, 4, 8, 0.5),
}


def _numel(shp):
    n = 1
    for d in shp:
        n *= d
    return n


@pytest.mark.parametrize("name", sorted(SPECS))
def test_arena_layout(name):
    sp = SPECS[name]
    a = ParamArena(sp, torch.device("cpu"))
    shapes = sp.param_shapes()
    assert set(a.offsets) == set(shapes)
    spans = sorted((a.offsets[n], a.offsets[n] + _numel(shapes[n]), n) for n in shapes)
    for (lo0, hi0, n0), (lo1, hi1, n1) in zip(spans, spans[1:]):
        assert hi0 <= lo1, f"{n0} overlaps {n1}"
    assert spans[-1][1] <= a.numel and a.numel % _ALIGN == 0 and a.embed_numel % _ALIGN == 0
    stacked = set(a.head_w_names) | set(a.head_b_names)
    for lo, hi, n in spans:
        if n not in stacked or n in (a.head_w_names[:1] + a.head_b_names[:1]):
            assert lo % _ALIGN == 0, f"{n} is not {_ALIGN}-element aligned"
    # embedding-stage block = prefix
    proj = {s.proj for s in sp.segments if s.proj is not None}
    for lo, hi, n in spans:
        is_embed = n.rsplit(".", 1)[0] in proj or n in ("task_embed", "pe", "ln.weight", "ln.bias")
        assert (hi <= a.embed_numel) == is_embed, n
    assert 0 < a.embed_numel < a.numel
    # stacked heads: back to back, in order, viewable as one (rows, H) matrix
    if a.head_w_names:
        w, b = a.stacked_head()
        rows = sum(shapes[n][0] for n in a.head_w_names)
        assert tuple(w.shape) == (rows, sp.hidden) and tuple(b.shape) == (rows,)
        off = a.offsets[a.head_w_names[0]]
        for n in a.head_w_names:
            assert a.offsets[n] == off
            off += _numel(shapes[n])
        a.view(a.head_w_names[-1]).fill_(3.0)
        assert float(w[-1, -1]) == 3.0 and float(w[0, 0]) == 0.0


def test_views_share_storage_with_the_arena():
    a = ParamArena(SPECS["hhi3"], torch.device("cpu"))
    v = a.view("proj_ttm.weight")
    v.fill_(2.0)
    o = a.offsets["proj_ttm.weight"]
    assert float(a.param[o]) == 2.0 and v.data_ptr() == a.param[o:].data_ptr()
    sd = a.state_dict()
    assert set(sd) == set(a.shapes) and float(sd["proj_ttm.weight"].mean()) == 2.0

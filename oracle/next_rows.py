"""TEST INFRASTRUCTURE ONLY - fixtures of the rows of SURVEY.md §8(f) that were pinned before their engine paths existed
(the same rows now also have regular cases in oracle/cases.py; this file keeps the greedy-decoding golden).

HOI EgoT2-g `TaskTranslationPromptTransformer` (HOI/models/multitask/video_model_builder.py:223-275): the restatement in
translator_oracle.py (hoi_g_*) is pinned here against the real reference class, and a golden (forward logits, CE loss,
gradient digests, greedy predict_ac tokens) is stored under tests/golden/next_hoi_g.npz so that round 2 can build the
engine path against a fixed target.  Run in the build container:   python -m oracle.next_rows
"""
from __future__ import annotations

import math
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from egot2_b200 import synth                               # noqa: E402
from oracle import translator_oracle as O                  # noqa: E402
from oracle.cases import grad_digest                       # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "next_hoi_g.npz")
H, HEADS, LAYERS, VOCAB, B, SEED = 128, 4, 2, 40, 3, 31


def param_shapes():
    out = {"proj_pnr.weight": (H, 8192), "proj_pnr.bias": (H,), "proj_oscc.weight": (H, 8192), "proj_oscc.bias": (H,),
           "proj_action_slow.weight": (H, 2048), "proj_action_slow.bias": (H,),
           "proj_action_fast.weight": (H, 256), "proj_action_fast.bias": (H,),
           "fc.weight": (VOCAB, H), "fc.bias": (VOCAB,), "ln.weight": (H,), "ln.bias": (H,), "task_embed": (1, 3, H),
           "embedding.weight": (VOCAB, H)}
    for i in range(LAYERS):
        for pre, atts, norms in ((f"transformer_encoder.layers.{i}.", ("self_attn",), ("norm1", "norm2")),
                                 (f"transformer_decoder.layers.{i}.", ("self_attn", "multihead_attn"), ("norm1", "norm2", "norm3"))):
            for a in atts:
                out[pre + a + ".in_proj_weight"] = (3 * H, H)
                out[pre + a + ".in_proj_bias"] = (3 * H,)
                out[pre + a + ".out_proj.weight"] = (H, H)
                out[pre + a + ".out_proj.bias"] = (H,)
            out[pre + "linear1.weight"] = (2048, H)
            out[pre + "linear1.bias"] = (2048,)
            out[pre + "linear2.weight"] = (H, 2048)
            out[pre + "linear2.bias"] = (H,)
            for n in norms:
                out[pre + n + ".weight"] = (H,)
                out[pre + n + ".bias"] = (H,)
    return out


def inputs():
    """Seeded weights (same scheme as egot2_b200.synth), features and task-prompt targets."""
    sd = {}
    for name, shape in param_shapes().items():
        g = synth._gen(SEED, name)
        if name in ("task_embed", "embedding.weight"):
            t = torch.randn(shape, generator=g)
        elif name.endswith(("norm1.weight", "norm2.weight", "norm3.weight")) or name == "ln.weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(("norm1.bias", "norm2.bias", "norm3.bias")) or name == "ln.bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[-1])
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        sd[name] = t.contiguous()
    g = synth._gen(SEED, "next.feats")
    feats = {"pnr": torch.randn((B, 16, 8192), generator=g), "oscc": torch.randn((B, 16, 8192), generator=g),
             "slow": torch.randn((B, 8, 2048), generator=g), "fast": torch.randn((B, 8, 256), generator=g)}
    target = torch.randint(5, VOCAB, (B, 3), generator=g)
    target[:, 0] = 4                                        # the task word ('action' in the reference vocabulary)
    return sd, feats, target


def oracle_outputs(P, feats, target):
    out = O.hoi_g_forward(P, feats["pnr"], feats["oscc"], feats["slow"], feats["fast"], target[:, :-1], HEADS)
    loss = torch.nn.functional.cross_entropy(out, target[:, 1:])
    return out, loss


def reference_module():
    from types import SimpleNamespace
    from oracle import ref_shims as rs
    mod = rs.load_hoi().multitask
    mod.load_lta_config = lambda f: rs.CfgNode(MODEL=rs.CfgNode(), CHECKPOINT_FILE_PATH=None)
    args = SimpleNamespace(hidden_dim=H, num_heads=HEADS, num_layers=LAYERS, dropout=0.1, pnr_cfg_file=None,
                           oscc_cfg_file=None, action_cfg_file=None)
    vocab = {("action" if i == 4 else f"w{i}"): i for i in range(VOCAB)}      # len(vocab) words; 'action' is the start token
    m = mod.TaskTranslationPromptTransformer(args, vocab)
    return m, rs


def reference_outputs(sd, feats, target):
    m, rs = reference_module()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("pos_embed.pe", "pnr_model", "oscc_model", "recognition_model")) for k in missing), missing
    m.eval()
    slow5 = feats["slow"].permute(0, 2, 1)[..., None, None]
    fast5 = feats["fast"].permute(0, 2, 1).repeat_interleave(4, dim=2)[..., None, None]

    class _Const(torch.nn.Module):           # a frozen backbone whose features are already known
        def __init__(self, value):
            super().__init__()
            self.value = value

        def forward(self, x, middle=False):
            return self.value
    m.pnr_model, m.oscc_model = _Const(feats["pnr"]), _Const(feats["oscc"])
    m.recognition_model = _Const([slow5, fast5])
    vid = [torch.zeros(B, 1)]                # predict_ac reads video_pnr[0].shape[0] / .type_as(video_pnr[0])
    out = m(vid, None, target[:, :-1])
    loss = torch.nn.functional.cross_entropy(out, target[:, 1:])
    toks = m.predict_ac([torch.zeros(B, 1)], None)
    return m, out, loss, toks


def main(write: bool = True):
    warnings.filterwarnings("ignore")
    torch.backends.mha.set_fastpath_enabled(False)
    sd, feats, target = inputs()
    m, out, loss, toks = reference_outputs(sd, feats, target)
    params = dict(m.named_parameters())
    names = [k for k in sd if k in params]
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o_out, o_loss = oracle_outputs(P, feats, target)
    torch.testing.assert_close(o_out, out.detach(), atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(o_loss.detach(), loss.detach(), atol=2e-5, rtol=1e-4)
    o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
    rec = {"output": out.detach().numpy(), "loss": np.float32(loss.item()), "predict_ac": toks.numpy()}
    for k, g_ref, g_o in zip(names, grads, o_grads):
        if g_ref is None:
            assert g_o is None or float(g_o.abs().max()) == 0.0, k
            g_ref = torch.zeros_like(params[k])
        else:
            err = float((g_ref - g_o).abs().max()) / (float(g_ref.abs().max()) + 1e-12)
            assert err < 2e-4, (k, err)
        rec["grad/" + k] = grad_digest(g_ref).numpy()
    o_toks = O.hoi_g_predict_ac({k: v for k, v in sd.items()}, feats["pnr"], feats["oscc"], feats["slow"], feats["fast"], 4, HEADS)
    assert torch.equal(o_toks, toks), (o_toks, toks)
    if write:
        np.savez_compressed(GOLDEN, **rec)
    print(f"next_hoi_g  out{tuple(out.shape)} loss={loss.item():.6f} predict_ac={toks.tolist()}  oracle==reference OK")
    return rec


# ----------------------------------------------------------------------------------------------- simple_vit siblings
GOLDEN_VIT = os.path.join(os.path.dirname(GOLDEN), "next_pnr_vit.npz")
VIT_B, VIT_SEED = 3, 33


def vit_shapes(three_task: bool):
    D, inner, mlp = 256, 8 * 128, 512
    out = {"proj1.weight": (D, 8192), "proj1.bias": (D,), "proj2.weight": (D, 8192), "proj2.bias": (D,),
           "pe": (1, 48 if three_task else 32, D)}
    if three_task:
        out.update({"proj3_slow.weight": (D, 2048), "proj3_slow.bias": (D,), "proj3_fast.weight": (D, 256),
                    "proj3_fast.bias": (D,), "ln.weight": (D,), "ln.bias": (D,)})
    else:
        out.update({"linear_head.0.weight": (D,), "linear_head.0.bias": (D,)})
    for i in range(3):
        a, f = f"transformer.layers.{i}.0.", f"transformer.layers.{i}.1.net."
        out.update({a + "norm.weight": (D,), a + "norm.bias": (D,), a + "to_qkv.weight": (3 * inner, D),
                    a + "to_out.weight": (D, inner), f + "0.weight": (D,), f + "0.bias": (D,), f + "1.weight": (mlp, D),
                    f + "1.bias": (mlp,), f + "3.weight": (D, mlp), f + "3.bias": (D,)})
    out.update({"linear_head.1.weight": (16, D), "linear_head.1.bias": (16,)})
    return out


def vit_inputs(three_task: bool):
    sd = {}
    for name, shape in vit_shapes(three_task).items():
        g = synth._gen(VIT_SEED + int(three_task), name)
        if name == "pe":
            t = torch.randn(shape, generator=g)
        elif name.endswith(("norm.weight", "net.0.weight", "ln.weight", "linear_head.0.weight")):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(("norm.bias", "net.0.bias", "ln.bias", "linear_head.0.bias")):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[-1])
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        sd[name] = t.contiguous()
    g = synth._gen(VIT_SEED, "next.vit.feats")
    feats = {"pnr": torch.randn((VIT_B, 16, 8192), generator=g), "oscc": torch.randn((VIT_B, 16, 8192), generator=g),
             "slow": torch.randn((VIT_B, 8, 2048), generator=g), "fast": torch.randn((VIT_B, 8, 256), generator=g)}
    labels = torch.randint(0, 16, (VIT_B,), generator=g)
    return sd, feats, labels


def vit_oracle(P, feats, labels, three_task: bool):
    out = O.hoi_pnr_vit_forward(P, feats["pnr"], feats["oscc"], feats["slow"] if three_task else None,
                                feats["fast"] if three_task else None)
    return out, O.bce_sigmoid_loss(out, torch.nn.functional.one_hot(labels, 16).float())


def vit_reference(sd, feats, labels, three_task: bool):
    from oracle import ref_shims as rs
    hoi = rs.load_hoi()
    cfg = rs.hoi_pnr_cfg(256, 3, 0.5, 0.1, "keyframe_localization")
    cfg.PRETRAIN.PNR_FT = cfg.PRETRAIN.OSCC_FT = True

    class _Const(torch.nn.Module):
        def __init__(self, value):
            super().__init__()
            self.value = value

        def forward(self, x, middle=False):
            return self.value
    if three_task:
        m = hoi.pnr3.TaskFusionMFTransformer3Task(cfg)
    else:
        m = hoi.pnr2.TaskFusionMFTransformer(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("linear_head.0.") for k in missing), missing          # 3-task: alias of ln
    m.eval()
    m.pnr_model, m.oscc_model = _Const(feats["pnr"]), _Const(feats["oscc"])
    if three_task:
        slow5 = feats["slow"].permute(0, 2, 1)[..., None, None]
        fast5 = feats["fast"].permute(0, 2, 1).repeat_interleave(4, dim=2)[..., None, None]
        m.recognition_model = _Const([slow5, fast5])
        out = m([torch.zeros(1)], None).squeeze(1)
    else:
        out = m([torch.zeros(1)]).squeeze(1)
    loss = torch.nn.BCELoss()(torch.sigmoid(out), torch.nn.functional.one_hot(labels, 16).float())
    return m, out, loss


def main_vit(write: bool = True):
    warnings.filterwarnings("ignore")
    rec = {}
    for three in (False, True):
        tag = "3task" if three else "2task"
        sd, feats, labels = vit_inputs(three)
        m, out, loss = vit_reference(sd, feats, labels, three)
        params = dict(m.named_parameters())
        names = [k for k in sd if k in params]
        grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        o_out, o_loss = vit_oracle(P, feats, labels, three)
        torch.testing.assert_close(o_out, out.detach(), atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(o_loss.detach(), loss.detach(), atol=2e-5, rtol=1e-4)
        o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
        rec[tag + "/output"] = out.detach().numpy()
        rec[tag + "/loss"] = np.float32(loss.item())
        for k, g_ref, g_o in zip(names, grads, o_grads):
            err = float((g_ref - g_o).abs().max()) / (float(g_ref.abs().max()) + 1e-12)
            assert err < 2e-4, (tag, k, err)
            rec[tag + "/grad/" + k] = grad_digest(g_ref).numpy()
        print(f"next_pnr_vit {tag}  out{tuple(out.shape)} loss={loss.item():.6f} params={len(names)}  oracle==reference OK")
    if write:
        np.savez_compressed(GOLDEN_VIT, **rec)
    return rec


if __name__ == "__main__":
    main()
    main_vit()

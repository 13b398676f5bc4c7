"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the REAL reference classes.

Run in the build container (needs /root/reference):   python -m oracle.make_golden [case names ...]
For each case in `oracle/cases.py` the reference translator class (verbatim code, backbones
stubbed — see ref_shims.py) is constructed, the seeded synthetic state_dict is loaded, and in
eval mode (dropout off; torch MHA fast path disabled so the documented slow-path math runs)
we record: the output, the task loss, and a digest of d(loss)/d(param) for every translator
parameter.  The script also asserts on the spot that the restatement in translator_oracle.py
reproduces those numbers (fp32, atol 2e-5 / rtol 1e-4).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shims as rs                      # noqa: E402
from oracle import translator_oracle as O               # noqa: E402
from oracle.cases import CASES, Case, case_inputs, grad_digest, oracle_forward_loss  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference(case: Case, hhi, hoi):
    sp = case.spec
    if sp.family == "hhi_ttm":
        three = len(sp.segments) == 3
        cls = hhi.ttm.TaskFusionMFTransformer3Task if three else hhi.ttm.TaskFusionMFTransformer2Task
        return cls(rs.hhi_args(sp.hidden, sp.heads, sp.layers, sp.p_layer, three))
    if sp.family == "hhi_asd":
        return hhi.asd.TaskFusionMFTransformer3Task(rs.hhi_args(sp.hidden, sp.heads, sp.layers, sp.p_layer, True))
    if sp.family == "hoi_pnr" and sp.head == "pool_linear":
        cfg = rs.hoi_pnr_cfg(sp.hidden, sp.layers, 0.5, sp.p_layer,
                             "keyframe_localization" if sp.n_out == 16 else "state_change_detection")
        cfg.MODEL.FEAT_DROPOUT_MODE = 0                      # HOI/configs/pnr/defaults.py:240
        cfg.PRETRAIN.PNR_FT = cfg.PRETRAIN.OSCC_FT = True
        return hoi.pnr2.TaskFusionMFTransformerDropout(cfg)
    if sp.family == "hoi_pnr" and sp.encoder == "simple_vit" and len(sp.segments) == 2:
        cfg = rs.hoi_pnr_cfg(256, 3, 0.5, 0.1, "keyframe_localization")
        cfg.PRETRAIN.PNR_FT = cfg.PRETRAIN.OSCC_FT = True
        return hoi.pnr2.TaskFusionMFTransformer(cfg)
    if sp.family == "hoi_pnr" and sp.encoder == "simple_vit":
        return hoi.pnr3.TaskFusionMFTransformer3Task(rs.hoi_pnr_cfg(256, 3, 0.5, 0.1, "keyframe_localization_2loader"))
    if sp.family == "hoi_pnr":
        task = "keyframe_localization_2loader" if sp.n_out == 16 else "state_change_detection"
        m = hoi.pnr3.TaskFusionMFTransformer3TaskDropout(rs.hoi_pnr_cfg(sp.hidden, sp.layers, sp.p_feat, sp.p_layer, task))
        return m
    if sp.family == "hoi_ar" and len(sp.segments) == 3:
        cfg = rs.CfgNode(MODEL=rs.CfgNode(NUM_CLASSES=list(sp.head_groups), TRANSLATION_HEADS=sp.heads,
                                          TRANSLATION_LAYERS=sp.layers, TRANSLATION_INPUT_FEATURES=sp.hidden,
                                          TRANSLATION_DROPOUT=sp.p_layer),
                         FORECASTING=rs.CfgNode(NUM_INPUT_CLIPS=2, INPUT_OFFSET=0),
                         PRETRAIN=rs.CfgNode(ACTION_CFG="x", LTA_CFG="x"))
        # the backbone section of the ctor (:216-227) only needs a config object from the (stubbed) loader
        hoi.lta3.load_lta_config = lambda f: rs.CfgNode(MODEL=rs.CfgNode(), CHECKPOINT_FILE_PATH=None)
        return hoi.lta3.TaskFusionMFTransformer2TaskAR(cfg)
    if sp.family == "hoi_ar":
        cfg = rs.CfgNode(MODEL=rs.CfgNode(NUM_CLASSES=list(sp.head_groups), TRANSLATION_HEADS=sp.heads,
                                          TRANSLATION_LAYERS=sp.layers, TRANSLATION_INPUT_FEATURES=sp.hidden,
                                          TRANSLATION_DROPOUT=sp.p_layer),
                         PRETRAIN=rs.CfgNode(PNR_CFG=None, OSCC_CFG=None, ACTION_CFG=None))
        return hoi.lta3.TaskFusionMFTransformer3Task(cfg)
    if sp.family == "hoi_lta" and len(sp.segments) == 2:
        return hoi.lta4.TaskFusionMFTransformer2Task(
            rs.hoi_lta_cfg(sp.hidden, sp.layers, sp.heads, sp.p_layer, sp.segments[0].tokens, sp.n_heads_out,
                           sp.head_groups, sp.p_head))
    if sp.family == "hoi_lta":
        return hoi.lta4.TaskFusionMFTransformerLTA4Task(
            rs.hoi_lta_cfg(sp.hidden, sp.layers, sp.heads, sp.p_layer, sp.segments[0].tokens, sp.n_heads_out,
                           sp.head_groups, sp.p_head))
    if sp.family == "hhi_g":
        args = rs.hhi_args(sp.hidden, sp.heads, sp.layers, sp.p_layer, True)
        vocab = {'</s>': 0, '<unk>': 1, 'ttm': 2, 'lam': 3, 'asd': 4, '0': 5, '1': 6}      # HHI/utils/utils.py:12-18
        return hhi.multitask.TaskTranslationPromptTransformer(args, vocab)
    if sp.family == "hoi_g":
        from types import SimpleNamespace
        mt = hoi.multitask
        mt.load_lta_config = lambda f: rs.CfgNode(MODEL=rs.CfgNode(), FORECASTING=rs.CfgNode(), CHECKPOINT_FILE_PATH=None,
                                                  CHECKPOINT_FILE_PATH_LTA=None)
        args = SimpleNamespace(hidden_dim=sp.hidden, num_heads=sp.heads, num_layers=sp.layers, dropout=sp.p_layer,
                               pnr_cfg_file=None, oscc_cfg_file=None, action_cfg_file=None, lta_cfg_file=None)
        vocab = {("action" if i == 4 else f"w{i}"): i for i in range(sp.vocab)}      # only len(vocab) and the start word matter
        cls = mt.TaskTranslationPromptTransformer6Task if sp.n_task_embed == 4 else mt.TaskTranslationPromptTransformer
        return cls(args, vocab)
    raise ValueError(sp.family)


class _Const(torch.nn.Module):
    """A frozen backbone whose output is already known."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *a, middle=False, **k):
        return self.fn(x)


def hoi_g_reference_forward(sp, m, feats, target_in):
    """Drive the real HOI EgoT2-g forward() with stand-in backbones that return this case's features."""
    if sp.g_mode == "lta":
        # 6Task.encode 'lta' (:325-331): encode_clips_pnr averages each clip's (B, 16, 8192) map over time, encode_clips
        # stacks one recognition feature per clip, lta_model returns (num_input, B, 2048)
        m.recognition_model = _Const(lambda x: x[0])
        m.lta_model = _Const(lambda x: feats["lta"].transpose(0, 1))
        # video_pnr stands for the PNR clip tensor AND (deep-copied, :327) the OSCC one: give the two stubs different
        # features by stacking them on a trailing axis the stubs pick from
        m.pnr_model = _Const(lambda x: x[0][..., 0].unsqueeze(1))
        m.oscc_model = _Const(lambda x: x[0][..., 1].unsqueeze(1))
        video_pnr = torch.stack([feats["pnr"], feats["oscc"]], dim=-1)                      # (B, n, 8192, 2)
        return m(video_pnr, [feats["action"]], target_in, "lta_verb")
    slow5 = feats["slow"].permute(0, 2, 1)[..., None, None]
    fast5 = feats["fast"].permute(0, 2, 1).repeat_interleave(4, dim=2)[..., None, None]     # (B,256,32,1,1)
    m.pnr_model, m.oscc_model = _Const(lambda x: feats["pnr"]), _Const(lambda x: feats["oscc"])
    m.recognition_model = _Const(lambda x: [slow5, fast5])
    vid = [torch.zeros(feats["pnr"].shape[0], 1)]
    if sp.n_task_embed == 4:
        return m(vid, None, target_in, "action")
    return m(vid, None, target_in)


def reference_forward_loss(case: Case, m, hhi, feats, labels, extra, keep_head_dropout: bool = False):
    """Drive the reference module through its own forward() and the loss its Lightning task uses.
    keep_head_dropout: leave the LTA head's Dropout(0.5) on (benchmark legs in train mode; the goldens switch it off)."""
    sp = case.spec
    if sp.family == "hhi_ttm":
        if len(sp.segments) == 3:
            out = m(*rs.hhi_inputs(feats))
        else:
            out = m(rs._DictVideo(feats), None)
        # HHI/tasks/ttm/video_task.py:23-24,36
        loss = torch.nn.CrossEntropyLoss(weight=torch.tensor([0.266, 0.734], device=out.device))(out, labels)
    elif sp.family == "hhi_asd":
        out = m(*rs.hhi_inputs(feats))
        lav = hhi.asd_loss.lossAV(dim=sp.hidden)
        lav.load_state_dict({"criterion.weight": torch.tensor([1.0, 4.0]), "FC.weight": extra["FC.weight"],
                             "FC.bias": extra["FC.bias"]})
        loss = lav(out, labels)[0]
    elif sp.family == "hhi_g":
        # HHI/tasks/multitask/video_tasktranslation.py:48-61 (the three forwards share one model; one case = one of them)
        out = m(*rs.hhi_inputs(feats), labels[:, :-1], sp.g_mode)                          # (rows, V, 2)
        loss = torch.nn.CrossEntropyLoss()(out, labels[:, 1:])
    elif sp.family == "hoi_g":
        out = hoi_g_reference_forward(sp, m, feats, labels[:, :-1])                         # (B, V, 2)
        loss = torch.nn.CrossEntropyLoss()(out, labels[:, 1:])                              # HOI/tasks/multitask/video_task.py:177,185
    elif sp.family == "hoi_pnr" and len(sp.segments) == 2:
        m.pnr_model = rs.FeatureBackbone(); m.pnr_model.slot = "pnr"
        m.oscc_model = rs.FeatureBackbone(); m.oscc_model.slot = "oscc"
        out = m([{"pnr": feats["pnr"], "oscc": feats["oscc"]}])
        if sp.n_out == 16:
            out = out.squeeze(1)
            with torch.autocast(out.device.type, enabled=False):      # BCELoss refuses to run under autocast (benchmark legs)
                loss = torch.nn.BCELoss()(torch.sigmoid(out.float()), torch.nn.functional.one_hot(labels, 16).float())
        else:
            out = out.squeeze(2)
            loss = torch.nn.functional.cross_entropy(out, labels)
    elif sp.family == "hoi_pnr":
        slow = extra["slow5"] if case.raw_slowfast else feats["slow"].permute(0, 2, 1)[..., None, None]
        fast = extra["fast5"] if case.raw_slowfast else \
            feats["fast"].permute(0, 2, 1).repeat_interleave(4, dim=2)[..., None, None]   # (B,256,32,1,1)
        m.pnr_model = rs.FeatureBackbone(); m.pnr_model.slot = "pnr"
        m.oscc_model = rs.FeatureBackbone(); m.oscc_model.slot = "oscc"

        class _SF(torch.nn.Module):
            def forward(self, x, middle=False):
                return [slow, fast]
        m.recognition_model = _SF()
        out = m([{"pnr": feats["pnr"], "oscc": feats["oscc"]}], None)
        if sp.n_out == 16:
            out = out.squeeze(1)                                   # (B,1,16) -> (B,16)
            # HOI/tasks/pnr/video_taskspecific_pnr.py:29-31
            with torch.autocast(out.device.type, enabled=False):      # BCELoss refuses to run under autocast (benchmark legs)
                loss = torch.nn.BCELoss()(torch.sigmoid(out.float()), torch.nn.functional.one_hot(labels, 16).float())
        else:
            out = out.squeeze(2)                                   # (B,2,1) -> (B,2)
            loss = torch.nn.functional.cross_entropy(out, labels)  # :143-146
    elif sp.family == "hoi_ar" and len(sp.segments) == 3:
        # forward() lines 229-236 run the frozen backbones under no_grad; we enter at the projections (:240-246)
        feat = torch.cat((m.proj_slow(feats["slow"]), m.proj_fast(feats["fast"]), m.proj_lta(feats["lta"])), dim=1)
        feat = m.ln(feat) + m.pe
        out_t = m.transformer(feat).mean(dim=1)
        preds = [m.linear_head1(out_t), m.linear_head2(out_t)]
        out = torch.cat(preds, dim=-1)
        loss = torch.nn.functional.cross_entropy(preds[0], labels[:, 0]) + torch.nn.functional.cross_entropy(preds[1], labels[:, 1])
    elif sp.family == "hoi_ar":
        slow = feats["slow"].permute(0, 2, 1)[..., None, None]
        fast = feats["fast"].permute(0, 2, 1).repeat_interleave(4, dim=2)[..., None, None]   # (B,256,32,1,1)
        m.pnr_model = rs.FeatureBackbone(); m.pnr_model.slot = "pnr"
        m.oscc_model = rs.FeatureBackbone(); m.oscc_model.slot = "oscc"

        class _SF(torch.nn.Module):
            def forward(self, x, middle=False):
                return [slow, fast]
        m.recognition_model = _SF()
        preds = m(None, [{"pnr": feats["pnr"], "oscc": feats["oscc"]}])
        out = torch.cat(preds, dim=-1)
        # HOI/tasks/lta/long_term_anticipation_taskspecfic.py:31-33
        loss = torch.nn.functional.cross_entropy(preds[0], labels[:, 0]) + torch.nn.functional.cross_entropy(preds[1], labels[:, 1])
    elif sp.family == "hoi_lta" and len(sp.segments) == 2:
        # forward() lines 510-512 run the backbones; we enter at the projection (:513-518)
        feat = torch.cat((feats["action"], m.proj_lta(feats["lta"])), dim=1)
        feat = m.ln(feat) + m.pe
        out_t = m.transformer(feat).mean(dim=1)
        m.head.training = True                                     # raw logits (train-mode head), dropout is p=0 below
        if hasattr(m.head, "dropout") and not keep_head_dropout:
            m.head.dropout.p = 0.0
        preds = m.decode(out_t)
        out = torch.cat(preds, dim=-1)
        loss = 0
        for h, head_x in enumerate(preds):
            for z in range(head_x.shape[1]):
                loss = loss + torch.nn.functional.cross_entropy(head_x[:, z], labels[:, z, h])
    elif sp.family == "hoi_lta":
        # forward() lines 355-358 run the backbones; we enter at the projections (359-363)
        feat = torch.cat((m.proj_pnr(feats["pnr"]), m.proj_oscc(feats["oscc"]), feats["action"],
                          m.proj_lta(feats["lta"])), dim=1)
        feat = m.ln(feat) + m.pe
        out_t = m.transformer(feat).mean(dim=1)
        m.head.training = True                                     # raw logits (train-mode head), dropout is p=0 below
        if hasattr(m.head, "dropout") and not keep_head_dropout:
            m.head.dropout.p = 0.0
        preds = m.decode(out_t)                                    # [(B,Z,115),(B,Z,478)]
        out = torch.cat(preds, dim=-1)
        # HOI/tasks/lta/long_term_anticipation_taskspecfic.py:177-183
        loss = 0
        for h, head_x in enumerate(preds):
            for z in range(head_x.shape[1]):
                loss = loss + torch.nn.functional.cross_entropy(head_x[:, z], labels[:, z, h])
    return out, loss


def main(only=()):
    """only: case names to (re)generate; default = every case.  state_dict_keys.json always covers all cases."""
    warnings.filterwarnings("ignore")
    torch.backends.mha.set_fastpath_enabled(False)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    hhi, hoi = rs.load_hhi(), rs.load_hoi()
    ref_keys = {}
    for name, case in CASES.items():
        m0 = build_reference(case, hhi, hoi)
        ref_keys[name] = {k: list(v.shape) for k, v in m0.state_dict().items()}
    import json
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(ref_keys, f, indent=0, sort_keys=True)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        sd, feats, labels, extra = case_inputs(case)
        m = build_reference(case, hhi, hoi)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        # everything we do not set must be a buffer / alias / backbone, never a translator weight
        allowed = ("pos_embed.pe", "linear_head.0.", "linear_head1.0.", "linear_head2.0.", "lam_model", "ttm_model", "asd_model",
                   "action_model", "lta_model", "pnr_model", "oscc_model", "recognition_model")
        assert not unexpected, unexpected
        assert all(k.startswith(allowed) for k in missing), missing
        m.eval()
        out, loss = reference_forward_loss(case, m, hhi, feats, labels, extra)
        params = dict(m.named_parameters())
        names = [k for k in sd if k in params]
        grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        rec = {"output": out.detach().numpy(), "loss": np.float32(loss.item())}
        for k, g in zip(names, grads):
            rec["grad/" + k] = grad_digest(g if g is not None else torch.zeros_like(params[k])).numpy()

        # --- pin the restatement against the reference right here ---
        P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        o_out, o_loss = oracle_forward_loss(case, P, feats, labels, extra)
        torch.testing.assert_close(o_out, out.detach(), atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(o_loss.detach(), loss.detach(), atol=2e-5, rtol=1e-4)
        o_grads = torch.autograd.grad(o_loss, [P[k] for k in names], allow_unused=True)
        for k, g_ref, g_o in zip(names, grads, o_grads):
            if g_ref is None:
                assert g_o is None or float(g_o.abs().max()) == 0.0, k
                continue
            scale = float(g_ref.abs().max()) + 1e-12
            err = float((g_ref - g_o).abs().max()) / scale
            assert err < 2e-4, (name, k, err)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **rec)
        print(f"{name:22s} out{tuple(out.shape)} loss={loss.item():.6f}  params={len(names)}  oracle==reference OK")


if __name__ == "__main__":
    main(tuple(sys.argv[1:]))

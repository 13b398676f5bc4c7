"""Batched, on-device versions of the PNR / OSCC evaluation metrics that directly follow the translator
(HOI/evaluation/pnr/metrics.py:11-80; called from HOI/tasks/pnr/video_taskspecific_pnr.py:46-54,79-86,150).

The reference walks the batch clip by clip and calls `.item()` several times per clip (a device sync each); here one
kernel of libegot2.so (`egot2_pnr_metrics`) computes arg-max, the state-change filter, the frame -> seconds mapping and
the sums for the whole batch, and ONE small device->host copy returns them.  Same function names, arguments and return
values as the reference module, so `from egot2_b200.metrics import ...` is a drop-in for
`from evaluation.pnr.metrics import ...`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L
from .engine import _stream


def _stack(x, dtype, device):
    if isinstance(x, (list, tuple)):
        x = torch.stack([torch.as_tensor(v) for v in x])
    return torch.as_tensor(x).to(device=device, dtype=dtype).contiguous()


def _run(preds, labels, sc_labels=None, fps=None, info=None) -> Tuple[int, int, float]:
    if isinstance(preds, (list, tuple)):
        preds = torch.stack(list(preds))
    if not preds.is_cuda:
        raise L.Egot2Error("egot2_b200.metrics runs on CUDA tensors only (no CPU fallback)")
    dev = preds.device
    logits = preds.reshape(preds.shape[0], -1).to(torch.float32).contiguous()
    B, n = logits.shape
    if B == 0:
        return 0, 0, 0.0
    lab = labels if not isinstance(labels, (list, tuple)) else torch.stack([torch.as_tensor(v) for v in labels])
    lab = torch.as_tensor(lab).to(dev)
    label_idx = label_onehot = None
    if lab.numel() == B:                                    # class indices (state change)
        label_idx = lab.reshape(B).to(torch.int64).contiguous()
    else:                                                   # one-hot rows (keyframe localisation)
        label_onehot = lab.reshape(B, n).to(torch.float32).contiguous()
    sc = None if sc_labels is None else _stack(sc_labels, torch.int64, dev).reshape(B)
    f = s = e = p = None
    if fps is not None:
        f = _stack(fps, torch.float64, dev).reshape(B)
        s = _stack(info["clip_start_frame"], torch.int64, dev).reshape(B)
        e = _stack(info["clip_end_frame"], torch.int64, dev).reshape(B)
        p = _stack(info["pnr_frame"], torch.int64, dev).reshape(B)
    counts = torch.empty(2, device=dev, dtype=torch.int64)
    dist = torch.empty(1, device=dev, dtype=torch.float64)
    ptr = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        L.call("egot2_pnr_metrics", B, n, logits.data_ptr(), ptr(label_idx), ptr(label_onehot), ptr(sc), ptr(f), ptr(s),
               ptr(e), ptr(p), None, counts.data_ptr(), dist.data_ptr(), _stream())
    out = torch.cat([counts.to(torch.float64), dist]).cpu()          # the one device -> host copy
    return int(out[0]), int(out[1]), float(out[2])


def state_change_accuracy(preds, labels) -> float:
    """metrics.py:11-21: fraction of clips whose arg-max class equals the label."""
    correct, total, _ = _run(preds, labels)
    return correct / total


def keyframe_accuracy(preds, labels, sc_labels) -> Tuple[int, int]:
    """metrics.py:24-34: (correct, total) over the clips with a state change."""
    correct, total, _ = _run(preds, labels, sc_labels)
    return correct, total


def keyframe_distance(preds, labels, sc_labels, fps, info, evaluate_trained: bool = False, sum: bool = False) -> Optional[float]:
    """metrics.py:37-80: mean (or sum) temporal error in seconds of the predicted keyframe over the clips with a state change."""
    _, total, dist = _run(preds, labels, sc_labels, fps, info)
    if total == 0:
        return None if evaluate_trained else 0.0
    return dist if sum else dist / total

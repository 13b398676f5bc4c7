"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's PNR / OSCC evaluation metrics
(HOI/evaluation/pnr/metrics.py), clip by clip exactly as the reference loops do.  Pinned against the reference
functions themselves in tests/test_metrics.py (where /root/reference exists)."""
from __future__ import annotations

import numpy as np
import torch


def state_change_accuracy(preds, labels):
    """metrics.py:11-21"""
    correct = total = 0
    for pred, label in zip(preds, labels):
        correct += int(int(torch.argmax(pred)) == int(label))
        total += 1
    return correct / total


def keyframe_accuracy(preds, labels, sc_labels):
    """metrics.py:24-34: only clips whose state-change label is 1 count."""
    correct = total = 0
    for pred, label, sc in zip(preds, labels, sc_labels):
        if int(sc) == 1:
            total += 1
            correct += int(int(torch.argmax(pred)) == int(torch.argmax(label)))
    return correct, total


def keyframe_distance(preds, labels, sc_labels, fps, info, evaluate_trained=False, sum=False):
    """metrics.py:37-80: predicted keyframe index -> frame offset ((end - start) / 16 * index, a float32 tensor
    expression in the reference) -> absolute error against (pnr_frame - start) in frames -> seconds."""
    errs = []
    for pred, sc, f, s, e, p in zip(preds, sc_labels, fps, info["clip_start_frame"], info["clip_end_frame"], info["pnr_frame"]):
        if int(sc) == 1:
            k = int(torch.argmax(pred))
            mapped = float(((e - s) / 16 * k).item())
            errs.append(abs(mapped - (int(p) - int(s))) / float(f))
    if not errs:
        return None if evaluate_trained else 0.0
    return float(np.sum(errs)) if sum else float(np.mean(errs))

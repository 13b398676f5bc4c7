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


# ---------------------------------------------------------------------------------------------- SURVEY 8f-3, second half
def ttm_segment_scores(batches):
    """HHI/utils/ttm/utils.py:57-80 (PostProcessor.update + _merge_output), minibatch by minibatch exactly as the reference
    walks them.  batches: [(logits (rows,2), targets)] with targets = [[uid], ..., label (index 2), ..., start (-3), end (-2),
    index (-1)].  Returns (groundtruth rows, prediction rows)."""
    gt, pred, cur, seg = [], [], None, []

    def merge():
        p = torch.softmax(torch.cat([o for o, _ in seg], dim=0).mean(0), dim=-1)
        start = min(int(t[-3]) for _, t in seg)
        end = max(int(t[-2]) for _, t in seg)
        uid, idx = cur.split(':')
        gt.append([uid, idx, start, end, int(seg[0][1][2])])
        pred.append([uid, idx, start, end, 1, float(p[1])])
    for out, tg in batches:
        sid = tg[0][0] + ':' + str(int(tg[-1]))
        if cur is not None and sid != cur:
            merge()
            seg = []
        cur = sid
        seg.append((out, tg))
    if seg:
        merge()
    return gt, pred


def topks_correct(preds, labels, ks):
    """HOI/evaluation/lta/lta_metrics.py:39-73: rows whose label is among the k largest predictions."""
    top = torch.topk(preds, max(ks), dim=1, largest=True, sorted=True)[1].t()
    hit = top.eq(labels.view(1, -1).expand_as(top))
    return [float(hit[:k].reshape(-1).float().sum()) for k in ks]


def levenshtein(a, b) -> int:
    """The third-party `editdistance` package (imported at lta_metrics.py:13; not vendored in /root/reference and not pinned
    in its environment files - any release, the function has not changed since 0.3): `editdistance.eval` is the Levenshtein
    distance - unit-cost insertions, deletions and substitutions, no transpositions (its README; the "Damerau"
    in the reference's docstring, lta_metrics.py:88-91, is not what the package computes)."""
    a, b = [int(x) for x in a], [int(x) for x in b]
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def edit_distance(preds, labels):
    """lta_metrics.py:87-96"""
    N, Z, K = preds.shape
    return float(np.mean([min(levenshtein(preds[n, :, k], labels[n]) / Z for k in range(K)) for n in range(N)]))


def aued(preds, labels):
    """lta_metrics.py:103-114 (same quirky shapes: every value is an array of shape (1,))"""
    N, Z, K = preds.shape
    preds = np.asarray(preds)
    labels = np.asarray(labels).reshape(N, Z)
    ED = np.vstack([edit_distance(preds[:, :z], labels[:, :z]) for z in range(1, Z + 1)])
    trapz = getattr(np, "trapezoid", None) or np.trapz
    out = {"AUED": trapz(y=ED, axis=0) / (Z - 1)}
    out.update({f"ED_{z}": ED[z] for z in range(Z)})
    return out

"""Batched, on-device versions of the LTA / action-recognition evaluation metrics that directly follow the translators
(HOI/evaluation/lta/lta_metrics.py:23-126; called from HOI/tasks/lta/long_term_anticipation.py:34-39,145-150,242).

Same function names, arguments and return values as the reference module, so `from egot2_b200 import lta_metrics as metrics`
is a drop-in for `from evaluation.lta import lta_metrics as metrics`.  The reference ranks with `torch.topk` and walks the clips
one by one through the `editdistance` package on the host; here one launch of libegot2.so per metric handles the whole batch
(`egot2_topk_correct`, `egot2_edit_distance_prefix`: all Z prefix lengths of AUED from one dynamic-programming table per
sequence pair) and one small device -> host copy returns the integer counts / sums.  CUDA tensors only - no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import _stream


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.Egot2Error(f"egot2_b200.lta_metrics.{what} runs on CUDA tensors only (no CPU fallback)")


def _gather(t: torch.Tensor) -> torch.Tensor:
    """du.all_gather_unaligned + cat: every rank's rows (the row counts may differ)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)], dim=0)


def topks_correct(preds: torch.Tensor, labels: torch.Tensor, ks: Sequence[int]) -> List[torch.Tensor]:
    """lta_metrics.py:39-73: for each k the number of rows whose label is among the k largest predictions (float tensors,
    like the reference's `.float().sum()`)."""
    assert preds.size(0) == labels.size(0), "Batch dim of predictions and labels must match"
    _need_cuda(preds, "topks_correct")
    dev = preds.device
    assert preds.dim() == 2, "preds: (N, classes)"
    p = preds.to(torch.float32).contiguous()
    lab = labels.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    ks = [int(k) for k in ks]
    arr = (C.c_int32 * len(ks))(*ks)
    out = torch.empty(len(ks), device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        L.call("egot2_topk_correct", p.shape[0], p.shape[1], p.data_ptr(), lab.data_ptr(), len(ks), C.cast(arr, C.c_void_p),
               out.data_ptr(), _stream())
    return list(out.to(torch.float32).unbind(0))


def topk_errors(preds, labels, ks):
    """lta_metrics.py:76-85"""
    return [(1.0 - x / preds.size(0)) * 100.0 for x in topks_correct(preds, labels, ks)]


def distributed_topk_errors(preds, labels, ks):
    """lta_metrics.py:23-36"""
    return topk_errors(_gather(preds), _gather(labels), ks)


def _prefix_sums(preds: torch.Tensor, labels: torch.Tensor):
    """(sum over the clips of the min-over-K Levenshtein distance for every prefix length 1..Z (int64 numpy), N, Z)"""
    _need_cuda(preds, "edit_distance")
    dev = preds.device
    N, Z, K = preds.shape
    p = preds.to(torch.int64).contiguous()
    lab = labels.reshape(N, Z).to(device=dev, dtype=torch.int64).contiguous()
    sums = torch.empty(Z, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        L.call("egot2_edit_distance_prefix", N, Z, K, p.data_ptr(), lab.data_ptr(), None, sums.data_ptr(), _stream())
    return sums.cpu().numpy(), N, Z


def edit_distance(preds, labels) -> float:
    """lta_metrics.py:87-96: mean over the clips of the lowest (over the K samples) edit distance / Z."""
    sums, N, Z = _prefix_sums(torch.as_tensor(preds), torch.as_tensor(labels))
    return float(sums[Z - 1]) / (Z * N)


def distributed_edit_distance(preds, labels):
    return edit_distance(_gather(preds), _gather(labels))


def AUED(preds, labels):
    """lta_metrics.py:103-114: edit distance at every prefix length and the area under that curve.  Same (quirky) return
    shapes as the reference: every value is a numpy array of shape (1,)."""
    preds, labels = torch.as_tensor(preds), torch.as_tensor(labels)
    sums, N, Z = _prefix_sums(preds, labels.squeeze(-1) if labels.dim() == 3 else labels)
    ED = (sums.astype(np.float64) / (np.arange(1, Z + 1, dtype=np.float64) * N)).reshape(Z, 1)
    trapz = getattr(np, "trapezoid", None) or np.trapz
    out = {"AUED": trapz(y=ED, axis=0) / (Z - 1)}
    out.update({f"ED_{z}": ED[z] for z in range(Z)})
    return out


def distributed_AUED(preds, labels):
    return AUED(_gather(preds), _gather(labels))

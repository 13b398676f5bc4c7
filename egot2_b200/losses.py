"""Head + loss on plain row vectors through egot2_head_loss_fwd/bwd (used by the lossAV drop-in).

linear_ce(x, W, b, labels, class_weight) = (x @ W^T + b, CrossEntropy(weight)(logits, labels))
with Linear, softmax-CE and every gradient computed by libegot2 (reference: HHI/tasks/asd/loss.py:11-30).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .engine import _dt, _stream, _torch_dt


class _LinearCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, labels, class_weight, dtype):
        if not x.is_cuda:
            raise L.Egot2Error("egot2_b200 losses run on CUDA only (no CPU fallback)")
        rows, H = x.shape
        n_out = W.shape[0]
        dev = x.device
        tdt = _torch_dt(dtype)
        hd = L.HeadDesc()
        hd.dtype, hd.B, hd.T, hd.H, hd.pool, hd.row_tokens, hd.use_ln = _dt(dtype), rows, 1, H, 1, 0, 0
        hd.n_out, hd.ln_eps = n_out, 1e-5
        hd.loss = L.LOSS_CE if labels is not None else L.LOSS_NONE
        xin = x.detach().to(tdt).contiguous()
        Wc = W.detach().to(tdt).contiguous()
        bc = b.detach().float().contiguous()
        hin, hout = L.HeadIn(), L.HeadOut()
        hin.x, hin.w, hin.b = xin.data_ptr(), Wc.data_ptr(), bc.data_ptr()
        keep = [xin, Wc, bc]
        pooled = torch.empty((rows, H), device=dev, dtype=torch.float32)
        g = torch.empty((rows, H), device=dev, dtype=tdt)
        logits = torch.empty((rows, n_out), device=dev, dtype=torch.float32)
        loss2 = torch.zeros(2, device=dev, dtype=torch.float32)
        row_loss = torch.empty((rows, 2), device=dev, dtype=torch.float32)
        argmax = torch.empty((rows,), device=dev, dtype=torch.int32)
        hout.pooled, hout.g, hout.logits = pooled.data_ptr(), g.data_ptr(), logits.data_ptr()
        hout.loss, hout.row_loss, hout.argmax = loss2.data_ptr(), row_loss.data_ptr(), argmax.data_ptr()
        keep += [pooled, g, logits, loss2, row_loss, argmax]
        if labels is not None:
            lab = labels.detach().to(device=dev, dtype=torch.int64).contiguous()
            hin.labels = lab.data_ptr()
            keep.append(lab)
            if class_weight is not None:
                cw = class_weight.detach().to(device=dev, dtype=torch.float32).contiguous()
                hin.class_weight = cw.data_ptr()
                keep.append(cw)
        L.call("egot2_head_loss_fwd", C.byref(hd), C.byref(hin), C.byref(hout), _stream())
        ctx.hd, ctx.hin, ctx.hout, ctx.keep = hd, hin, hout, keep
        ctx.shapes = (rows, H, n_out, dtype)
        ctx.mark_non_differentiable(argmax)
        ctx.set_materialize_grads(False)
        return logits, loss2[0], argmax

    @staticmethod
    def backward(ctx, dlogits, dloss, _dargmax):
        rows, H, n_out, dtype = ctx.shapes
        hd, hin, hout = ctx.hd, ctx.hin, ctx.hout
        dev = ctx.keep[0].device
        tdt = _torch_dt(dtype)
        dx_total = torch.zeros((rows, H), device=dev, dtype=torch.float32)
        dW = torch.zeros((n_out, H), device=dev, dtype=torch.float32)
        db = torch.zeros((n_out,), device=dev, dtype=torch.float32)
        ws = torch.empty(L.load().egot2_head_workspace_bytes(C.byref(hd)), device=dev, dtype=torch.uint8)

        def run(loss_kind, dl, scale_tensor):
            hd.loss = loss_kind
            gW = torch.zeros_like(dW); gb = torch.zeros_like(db)
            hg = L.HeadGrads()
            hg.w, hg.b = gW.data_ptr(), gb.data_ptr()
            dx = torch.empty((rows, 1, H), device=dev, dtype=tdt)
            L.call("egot2_head_loss_bwd", C.byref(hd), C.byref(hin), C.byref(hout), dl.data_ptr(), 1.0, dx.data_ptr(),
                   C.byref(hg), ws.data_ptr(), ws.numel(), _stream())
            s = 1.0 if scale_tensor is None else scale_tensor
            dx_total.add_(dx.view(rows, H).float() * s)
            dW.add_(gW * s)
            db.add_(gb * s)

        has_loss = hin.labels is not None and dloss is not None
        if has_loss:
            run(L.LOSS_CE, torch.empty((rows, n_out), device=dev, dtype=torch.float32), dloss)
        if dlogits is not None:
            run(L.LOSS_NONE, dlogits.float().contiguous().clone(), None)
        hd.loss = L.LOSS_CE if hin.labels is not None else L.LOSS_NONE
        return dx_total, dW, db, None, None, None


def linear_ce(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor, labels: Optional[torch.Tensor],
              class_weight: Optional[torch.Tensor], dtype: str = "fp32"):
    """Returns (logits fp32 (rows,n_out), loss scalar or None)."""
    logits, loss, _ = _LinearCE.apply(x, W, b, labels, class_weight, dtype)
    return logits, (loss if labels is not None else None)

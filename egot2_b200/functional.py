"""torch.autograd glue: one Function wraps the whole translator (features + parameters -> output),
so the reference's Lightning tasks can keep calling `loss.backward()` on what our modules return."""
from __future__ import annotations

from typing import List, Sequence

import torch

from .engine import TranslatorEngine


class _TranslatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: TranslatorEngine, training: bool, seed: int, n_feats: int, names: Sequence[str],
                prompt, *tensors: torch.Tensor):
        feats = [t.detach() for t in tensors[:n_feats]]
        act = engine.forward(feats, training=training, seed=seed, prompt=prompt)
        ctx.engine, ctx.act, ctx.names, ctx.n_feats = engine, act, list(names), n_feats
        ctx.feat_needs = [bool(t.requires_grad) for t in tensors[:n_feats]]
        ctx.param_needs = [bool(t.requires_grad) for t in tensors[n_feats:]]
        return act.t["out"]

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        eng: TranslatorEngine = ctx.engine
        # a fresh flat buffer per backward: returned views never alias a later step's gradients
        grad = torch.zeros_like(eng.arena.grad)
        _, dfeats = eng.backward(ctx.act, dout=dout, grad=grad, zero_grad=False, want_dfeat=ctx.feat_needs)
        out: List = [None, None, None, None, None, None]
        for need, df, f in zip(ctx.feat_needs, dfeats, ctx.act.feats):
            out.append(df.to(f.dtype) if (need and df is not None) else None)
        for need, name in zip(ctx.param_needs, ctx.names):
            out.append(eng.arena.view(name, grad) if need else None)
        ctx.act = None
        return tuple(out)


def translator_apply(engine: TranslatorEngine, feats: Sequence[torch.Tensor], params: Sequence[torch.Tensor],
                     names: Sequence[str], training: bool, seed: int, prompt=None) -> torch.Tensor:
    """prompt: (rows, S) int64 decoder tokens for the EgoT2-g translator; None otherwise."""
    return _TranslatorFn.apply(engine, training, seed, len(feats), tuple(names), prompt, *feats, *params)

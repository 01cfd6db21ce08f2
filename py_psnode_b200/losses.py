"""Fused masked squared-error loss of the training scripts: one CUDA pass for the value, one for the
gradient, instead of the ~10 eager launches of

    torch.sum(Loss_func(x_pred, x, reduction='none') * mask)          # neural_00_ODE_01_no_encode.py:353-354

and of its weighted DAE variant (neural_01_DAE_01_no_encode.py:414-418).  CUDA tensors only: like the integrators this
module has no CPU implementation and raises if the native library is missing."""
from typing import Optional

import ctypes as C
import torch

from . import _native as N


def _series(t: torch.Tensor, outer: int, inner: int) -> N.Series:
    return N.Series(t.data_ptr(), t.stride(outer), t.stride(inner))


def _row_view(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 3:
        raise ValueError(f"{name} must have 3 dimensions (T, B, X) or (B, T, X), got {tuple(t.shape)}")
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name} must be a float32 CUDA tensor (there is no CPU path)")
    return t if (t.shape[-1] == 1 or t.stride(-1) == 1) else t.contiguous()


class _MaskedSSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, feat_weight):
        pred, target = _row_view(pred, "pred"), _row_view(target, "target")
        if target.shape != pred.shape:
            raise ValueError(f"target {tuple(target.shape)} does not match pred {tuple(pred.shape)}")
        if mask.shape[:2] != pred.shape[:2] or not (mask.shape[-1] == 1 or mask.stride(-1) == 0):
            raise ValueError("mask must be (.., .., 1) or a broadcast view of it (one value per trajectory and grid point)")
        mask = _row_view(mask if mask.shape[-1] == 1 else mask[..., :1], "mask")
        X = pred.shape[-1]
        if feat_weight is not None:
            feat_weight = feat_weight.to(device=pred.device, dtype=torch.float32).contiguous()
            if feat_weight.numel() != X:
                raise ValueError(f"feat_weight must have {X} entries")
        # the dimension with the larger stride is the outer loop of the kernel
        outer, inner = (0, 1) if pred.stride(0) >= pred.stride(1) else (1, 0)
        lib = N.lib()
        loss = torch.empty(1, device=pred.device, dtype=torch.float32)
        ws = torch.empty(int(lib.psnode_masked_sse_workspace()), device=pred.device, dtype=torch.uint8)
        args = (_series(pred, outer, inner), _series(target, outer, inner), _series(mask, outer, inner))
        wptr = C.c_void_p(feat_weight.data_ptr()) if feat_weight is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
        with torch.cuda.device(pred.device):
            N.check(lib.psnode_masked_sse(C.byref(args[0]), C.byref(args[1]), C.byref(args[2]), wptr, pred.shape[outer],
                                          pred.shape[inner], X, C.c_void_p(loss.data_ptr()), C.c_void_p(ws.data_ptr()),
                                          ws.numel(), stream), "psnode_masked_sse")
        ctx.save_for_backward(pred, target, mask, feat_weight)
        ctx.dims = (outer, inner)
        return loss[0]

    @staticmethod
    def backward(ctx, go):
        pred, target, mask, feat_weight = ctx.saved_tensors
        outer, inner = ctx.dims
        lib = N.lib()
        grad = torch.empty_like(pred)           # same (dense) layout as pred
        if grad.stride(-1) != 1:
            grad = torch.empty(pred.shape, device=pred.device, dtype=torch.float32)
        up = go.detach().to(torch.float32).reshape(1).contiguous()
        wptr = C.c_void_p(feat_weight.data_ptr()) if feat_weight is not None else None
        g = _series(grad, outer, inner)
        a = (_series(pred, outer, inner), _series(target, outer, inner), _series(mask, outer, inner))
        with torch.cuda.device(pred.device):
            N.check(lib.psnode_masked_sse_grad(C.byref(a[0]), C.byref(a[1]), C.byref(a[2]), wptr, pred.shape[outer],
                                               pred.shape[inner], pred.shape[-1], C.c_void_p(up.data_ptr()), C.byref(g),
                                               C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)),
                    "psnode_masked_sse_grad")
        return grad, (-grad if ctx.needs_input_grad[1] else None), None, None


def masked_sse(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor,
               feat_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_c w_c * sum(mask * (pred - target)^2): the numerator of the scripts' masked MSE (divide by `mask.sum()`).

    pred / target: (T, B, X) or (B, T, X) float32 CUDA tensors (any strides with a unit feature stride are used in
    place); mask: one value per (trajectory, grid point), shape (.., .., 1); feat_weight: optional X per-feature weights
    (the DAE script counts feature 1 ten times: ones(X) with w[1] = 10)."""
    return _MaskedSSE.apply(pred, target, mask, feat_weight)


def masked_sse_grad_into(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, feat_weight: Optional[torch.Tensor],
                         upstream: torch.Tensor) -> torch.Tensor:
    """upstream[0] * d/dpred masked_sse(pred, target, mask, feat_weight) as a new tensor shaped like pred (one native pass;
    used when a reverse sweep cannot form the loss gradient on the fly)."""
    pred, target = _row_view(pred, "pred"), _row_view(target, "target")
    mask = _row_view(mask if mask.shape[-1] == 1 else mask[..., :1], "mask")
    outer, inner = (0, 1) if pred.stride(0) >= pred.stride(1) else (1, 0)
    grad = torch.empty(pred.shape, device=pred.device, dtype=torch.float32)
    lib = N.lib()
    wptr = C.c_void_p(feat_weight.data_ptr()) if feat_weight is not None else None
    g = _series(grad, outer, inner)
    a = (_series(pred, outer, inner), _series(target, outer, inner), _series(mask, outer, inner))
    with torch.cuda.device(pred.device):
        N.check(lib.psnode_masked_sse_grad(C.byref(a[0]), C.byref(a[1]), C.byref(a[2]), wptr, pred.shape[outer], pred.shape[inner],
                                           pred.shape[-1], C.c_void_p(upstream.data_ptr()), C.byref(g),
                                           C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)), "psnode_masked_sse_grad")
    return grad

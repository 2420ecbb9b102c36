"""JacobianTensor: a torch.Tensor subclass that sends batched tiny matmuls to gsb_small_bmm.

The reference's `GSFunction.backward` (gsplat/gsmodel.py:72-85) multiplies the per-Gaussian
Jacobians our operators return with ~12 `@` products such as [N,1,3]@[N,3,3].  torch lowers
each of them to batched GEMV/GEMM library kernels in chunks of 65535 matrices, which costs
13 ms of a 15 ms training step at N = 1M even though the data is read exactly once.  The
operators therefore return their Jacobians (and `splatB` its four gradients) as
`JacobianTensor`s: any `a @ b` / `torch.matmul` / `torch.bmm` that involves one and has the
batched-tiny shape goes to one streaming CUDA kernel; every other operation falls through to
torch unchanged.  `squeeze` / `unsqueeze` / `reshape` -- the last thing the reference applies
to each gradient before returning it to autograd -- hand back plain tensors, so optimizers
never see the subclass.  Set GSB_FAST_BMM=0 to disable (plain tensors everywhere).
"""
import os

import torch

from . import _lib

ENABLED = os.environ.get("GSB_FAST_BMM", "1") != "0"
_MAX_ELEMS = 96  # m*k and k*n per batch element handled by the streaming kernel

_MATMULS = {torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__, torch.bmm, torch.Tensor.bmm}
_UNWRAP_AFTER = {torch.Tensor.squeeze, torch.squeeze, torch.Tensor.unsqueeze, torch.unsqueeze,
                 torch.Tensor.reshape, torch.reshape, torch.Tensor.view, torch.Tensor.flatten, torch.flatten}


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, JacobianTensor) else t


def _fast_matmul(a, b):
    """C = a @ b for a:[B,m,k], b:[B,k,n] or [k,n]; None if the shapes are not the tiny batched kind."""
    if not (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor)):
        return None
    if a.dim() != 3 or b.dim() not in (2, 3) or not (a.is_cuda and b.is_cuda):
        return None
    if a.dtype != torch.float32 or b.dtype != torch.float32 or a.device != b.device:
        return None
    # the streaming kernel has no autograd node: whenever the product would be recorded (grad mode
    # on and an operand requires grad -- double backward, a loss on a Jacobian) torch must do it
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return None
    B, m, k = a.shape
    if b.dim() == 3:
        if b.shape[0] != B or b.shape[1] != k:
            return None
        n, shared = b.shape[2], 0
    else:
        if b.shape[0] != k:
            return None
        n, shared = b.shape[1], 1
    if m * k > _MAX_ELEMS or k * n > _MAX_ELEMS or m * k > 32 or B == 0:
        return None
    a, b = _plain(a).contiguous(), _plain(b).contiguous()
    out = torch.empty((B, m, n), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.gsb_small_bmm(B, m, k, n, a.data_ptr(), b.data_ptr(), shared, out.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), lib)
    return out


class JacobianTensor(torch.Tensor):
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _MATMULS and len(args) == 2 and not kwargs:
            r = _fast_matmul(args[0], args[1])
            if r is not None:
                return r.as_subclass(JacobianTensor)
        elif func is torch.Tensor.__rmatmul__ and len(args) == 2 and not kwargs:
            r = _fast_matmul(args[1], args[0])
            if r is not None:
                return r.as_subclass(JacobianTensor)
        out = super().__torch_function__(func, types, args, kwargs)
        if func in _UNWRAP_AFTER and isinstance(out, JacobianTensor):
            return out.as_subclass(torch.Tensor)
        return out


def wrap(t):
    """Tag an operator output so the reference's Jacobian chain takes the fast path."""
    if not ENABLED or t is None:
        return t
    return t.as_subclass(JacobianTensor)

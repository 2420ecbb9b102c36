"""Fused training loss: a drop-in for the reference's
`gsplat.pytorch_ssim.gau_loss(image, gt_image, loss_lambda=0.2)` (pytorch_ssim.py:64-67, used at
train.py:52).  One call computes the loss and dloss/dimage on the GPU (two HBM-bound kernels,
csrc/loss.cu); autograd only scales the stored gradient by the incoming grad_output.

Swap-in for train.py:  `from easygaussiansplatting_b200.loss import gau_loss`.
"""
import torch

from . import _lib
from .ops import _chk, _L, _ptr, _same_device, _stream


def gau_loss_with_grad(image, gt_image, loss_lambda=0.2, want_grad=True):
    """-> (loss scalar tensor [1], dloss_dimage [3,H,W] or None)"""
    image = _chk(image, "image", ndim=3); gt = _chk(gt_image, "gt_image", ndim=3)
    _same_device(image, gt)
    if image.shape != gt.shape or image.shape[0] != 3:
        raise ValueError("image and gt_image must both be [3,H,W], got %s and %s" % (tuple(image.shape), tuple(gt.shape)))
    H, W = int(image.shape[1]), int(image.shape[2])
    lib = _L()
    loss = torch.empty((1,), dtype=torch.float32, device=image.device)
    grad = torch.empty_like(image) if want_grad else None
    with torch.cuda.device(image.device):
        ws_bytes = lib.gsb_gau_loss_workspace_bytes(H, W)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=image.device)
        _lib.check(lib.gsb_gau_loss(H, W, _ptr(image), _ptr(gt), float(loss_lambda), _ptr(loss), _ptr(grad), _ptr(ws),
                                    ws_bytes, _stream()), lib)
    return loss, grad


class _GauLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt_image, loss_lambda):
        loss, grad = gau_loss_with_grad(image.detach(), gt_image.detach(), loss_lambda, want_grad=True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        return grad * grad_output, None, None


def gau_loss(image, gt_image, loss_lambda=0.2):
    """Same signature and value as the reference's gau_loss; differentiable w.r.t. `image`."""
    return _GauLoss.apply(image, gt_image, loss_lambda)

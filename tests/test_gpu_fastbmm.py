"""JacobianTensor / gsb_small_bmm: the reference's unmodified Jacobian chain
(gsmodel.py:72-85, mirrored in gsfunction.GSFunction) gives the same gradients whether its
`@` products run through torch's batched GEMM or through the streaming kernel."""
import numpy as np
import pytest
import torch

from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_small_bmm_matches_torch():
    from easygaussiansplatting_b200.jacobian_tensor import JacobianTensor, wrap
    g = torch.Generator(device=DEV).manual_seed(0)
    for (m, k, n) in [(1, 3, 3), (1, 3, 6), (1, 6, 4), (1, 6, 3), (1, 2, 3), (3, 1, 16), (2, 5, 7), (4, 8, 12)]:
        A = torch.randn((1000, m, k), device=DEV, generator=g)
        B = torch.randn((1000, k, n), device=DEV, generator=g)
        ref = torch.bmm(A, B)
        got = wrap(A) @ B
        assert isinstance(got, JacobianTensor) and torch.allclose(got.as_subclass(torch.Tensor), ref, rtol=1e-5, atol=1e-5)
        got2 = A @ wrap(B)
        assert torch.allclose(got2.as_subclass(torch.Tensor), ref, rtol=1e-5, atol=1e-5)
        # non-contiguous left operand (the reference permutes dloss_dcolors) and a shared right operand
        At = torch.randn((1000, k, m), device=DEV, generator=g)
        got3 = wrap(At).permute(0, 2, 1) @ B
        assert torch.allclose(got3.as_subclass(torch.Tensor), torch.bmm(At.permute(0, 2, 1), B), rtol=1e-5, atol=1e-5)
        Bs = torch.randn((k, n), device=DEV, generator=g)
        got4 = wrap(A) @ Bs
        assert torch.allclose(got4.as_subclass(torch.Tensor), A @ Bs, rtol=1e-5, atol=1e-5)
    # shapes outside the tiny-batched kind fall through to torch
    big = wrap(torch.randn((4, 64, 64), device=DEV)) @ torch.randn((4, 64, 64), device=DEV)
    assert big.shape == (4, 64, 64)
    # the last ops of the reference chain return plain tensors
    x = wrap(torch.randn((10, 1, 3), device=DEV))
    assert type(x.squeeze()) is torch.Tensor and type(x.reshape(10, -1)) is torch.Tensor
    assert type(x.squeeze().unsqueeze(1)) is torch.Tensor


def test_reference_chain_same_with_and_without_fast_bmm():
    from easygaussiansplatting_b200 import jacobian_tensor as jt
    from easygaussiansplatting_b200.gsfunction import Camera, GSFunction
    N, W, H = 30000, 320, 240
    sc = synthetic_scene(N, W, H, sh_dim=48, seed=4)
    cam = Camera(W, H, sc["fx"], sc["fy"], sc["cx"], sc["cy"], t(sc["Rcw"]), t(sc["tcw"]), t(sc["twc"]))
    dl = t(upstream_gradient(W, H, 4) * (3.0 * W * H))
    res = []
    for enabled in (True, False):
        jt.ENABLED = enabled
        try:
            P = {k: t(sc[k]).requires_grad_() for k in ("pws", "shs", "scales", "rots")}
            al = t(sc["alphas"][:, None]).requires_grad_()
            us0 = torch.zeros((N, 2), device=DEV, requires_grad=True)
            image, _ = GSFunction.apply(P["pws"], P["shs"], al, P["scales"], P["rots"], us0, cam)
            image.backward(dl)
            res.append([P["pws"].grad, P["shs"].grad, al.grad, P["scales"].grad, P["rots"].grad, us0.grad])
        finally:
            jt.ENABLED = True
    for a, b in zip(*res):
        assert type(a) is torch.Tensor and a.shape == b.shape
        s = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-5 * s + 1e-12

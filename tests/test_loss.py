"""Training loss.  CPU: the oracle's gau_loss against the fixture produced by
the reference's own gsplat/pytorch_ssim.py (tests/golden/make_golden_loss.py).  GPU: the fused
kernels against the oracle, the fixture, and a torch transcription of pytorch_ssim.py:26-67."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden", "loss.npz")


def test_oracle_matches_reference_fixture():
    d = dict(np.load(G))
    for k in "ab":
        r = orc.gau_loss(d["img_" + k], d["gt_" + k])
        assert abs(r["loss"] - float(d["loss_" + k])) < 2e-6          # reference runs in float32
        ref = d["grad_" + k]
        assert np.abs(r["dloss_dimage"] - ref).max() / np.abs(ref).max() < 2e-4
    w = orc.ssim_window()
    assert abs(w.sum() - 1) < 1e-6 and w.argmax() == 5 and np.allclose(w, w[::-1])


def test_oracle_identical_images_and_lambda():
    img = np.random.default_rng(0).uniform(0, 1, (3, 20, 24)).astype(np.float32)
    r = orc.gau_loss(img, img.copy())
    assert abs(r["l1"]) < 1e-12 and abs(r["ssim"] - 1) < 1e-9 and abs(r["loss"]) < 1e-9
    gt = np.clip(img + 0.1, 0, 1).astype(np.float32)
    a, b = orc.gau_loss(img, gt, 0.0), orc.gau_loss(img, gt, 1.0)
    assert abs(a["loss"] - a["l1"]) < 1e-12 and abs(b["loss"] - (1 - b["ssim"])) < 1e-12


def test_oracle_gradient_is_numerical_derivative():
    rng = np.random.default_rng(3)
    img = rng.uniform(0.2, 0.8, (3, 14, 13)).astype(np.float32)
    gt = rng.uniform(0.2, 0.8, (3, 14, 13)).astype(np.float32)
    r = orc.gau_loss(img, gt)
    for (c, y, x) in [(0, 0, 0), (1, 7, 6), (2, 13, 12), (0, 5, 11)]:
        h = np.float32(2.0 ** -10)  # exactly representable step
        p, m = img.copy(), img.copy()
        p[c, y, x] += h; m[c, y, x] -= h
        num = (orc.gau_loss(p, gt, want_grad=False)["loss"] - orc.gau_loss(m, gt, want_grad=False)["loss"]) / (
            float(p[c, y, x]) - float(m[c, y, x]))
        assert abs(num - r["dloss_dimage"][c, y, x]) < 2e-6 + 1e-3 * abs(num)


def _torch_gau_loss(image, gt, lam=0.2):
    """transcription of pytorch_ssim.py:26-67 for CUDA tensors (no reference import on the GPU box)"""
    import torch.nn.functional as F
    g = torch.tensor([np.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    win = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous().to(image.device)
    mu1 = F.conv2d(image, win, padding=5, groups=3); mu2 = F.conv2d(gt, win, padding=5, groups=3)
    s11 = F.conv2d(image * image, win, padding=5, groups=3) - mu1.pow(2)
    s22 = F.conv2d(gt * gt, win, padding=5, groups=3) - mu2.pow(2)
    s12 = F.conv2d(image * gt, win, padding=5, groups=3) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1.pow(2) + mu2.pow(2) + C1) * (s11 + s22 + C2))
    return (1 - lam) * torch.abs(image - gt).mean() + lam * (1 - ssim.mean())


@pytest.mark.gpu
@pytest.mark.parametrize("H,W", [(37, 53), (16, 16), (128, 96), (1080, 1920)])
def test_gpu_gau_loss_vs_oracle(H, W):
    from easygaussiansplatting_b200.loss import gau_loss, gau_loss_with_grad
    rng = np.random.default_rng(H * 1000 + W)
    yy, xx = np.mgrid[0:H, 0:W]
    base = np.stack([0.5 + 0.4 * np.sin(xx / 9.0 + c) * np.cos(yy / 6.0 - c) for c in range(3)])
    img = np.clip(base + rng.normal(scale=0.08, size=base.shape), -0.2, 1.3).astype(np.float32)
    gt = np.clip(base + rng.normal(scale=0.02, size=base.shape), 0, 1).astype(np.float32)
    ref = orc.gau_loss(img, gt)
    ti = torch.from_numpy(img).cuda().requires_grad_(); tg = torch.from_numpy(gt).cuda()
    loss = gau_loss(ti, tg)
    loss.backward()
    assert abs(loss.item() - ref["loss"]) < 2e-6
    got = ti.grad.cpu().numpy().astype(np.float64)
    e = np.abs(got - ref["dloss_dimage"]).max() / np.abs(ref["dloss_dimage"]).max()
    assert e < 3e-4, e   # float32 E[x^2] - mu^2 cancellation, the same formulation as the reference
    # the torch transcription of the reference agrees too
    t2 = torch.from_numpy(img).cuda().requires_grad_()
    l2 = _torch_gau_loss(t2, tg)
    l2.backward()
    assert abs(l2.item() - loss.item()) < 2e-6
    assert (t2.grad - ti.grad).abs().max().item() / t2.grad.abs().max().item() < 3e-4
    # loss only (no gradient buffer) and scaling through autograd
    lo, g = gau_loss_with_grad(tg + 0.05, tg, 0.2, want_grad=False)
    assert g is None and lo.item() > 0
    t3 = torch.from_numpy(img).cuda().requires_grad_()
    (3.0 * gau_loss(t3, tg)).backward()
    assert torch.allclose(t3.grad, 3.0 * ti.grad, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_gpu_gau_loss_reference_fixture():
    from easygaussiansplatting_b200.loss import gau_loss
    d = dict(np.load(G))
    for k in "ab":
        ti = torch.from_numpy(d["img_" + k]).cuda().requires_grad_()
        loss = gau_loss(ti, torch.from_numpy(d["gt_" + k]).cuda())
        loss.backward()
        assert abs(loss.item() - float(d["loss_" + k])) < 2e-6
        ref = d["grad_" + k]
        assert np.abs(ti.grad.cpu().numpy() - ref).max() / np.abs(ref).max() < 3e-4

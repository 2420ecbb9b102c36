"""Golden fixture for the training loss: imports the reference's own
gsplat/pytorch_ssim.py (gau_loss, :64-67) on the CPU and stores loss + autograd gradient for
two small image pairs.  Run in the dev container:  python tests/golden/make_golden_loss.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, REF)
sys.modules.setdefault("gsplatcu", types.ModuleType("gsplatcu"))

from gsplat.pytorch_ssim import gau_loss  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    out = {}
    for name, (H, W) in (("a", (37, 53)), ("b", (16, 16))):
        # smooth-ish images in [0,1] plus noise, like a render vs its target
        yy, xx = np.mgrid[0:H, 0:W]
        base = np.stack([0.5 + 0.4 * np.sin(xx / 7.0 + c) * np.cos(yy / 5.0 - c) for c in range(3)])
        img = np.clip(base + rng.normal(scale=0.08, size=base.shape), -0.2, 1.3).astype(np.float32)
        gt = np.clip(base + rng.normal(scale=0.02, size=base.shape), 0, 1).astype(np.float32)
        t = torch.from_numpy(img).requires_grad_()
        loss = gau_loss(t, torch.from_numpy(gt))
        loss.backward()
        out["img_" + name], out["gt_" + name] = img, gt
        out["loss_" + name] = np.float64(loss.item())
        out["grad_" + name] = t.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
    print("loss.npz", out["loss_a"], out["loss_b"])


if __name__ == "__main__":
    main()

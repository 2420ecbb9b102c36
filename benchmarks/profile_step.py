#!/usr/bin/env python
"""A few fwd+bwd steps of BASELINE config 2 (no CPU work) -- the command profiled by ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file L.csv \
      python benchmarks/profile_step.py 2 fused
  ncu --set full --clock-control none --import-source on -k regex:k_draw -s 2 -c 2 -o prof \
      python benchmarks/profile_step.py 2 fused
usage: profile_step.py [steps] [fused|ops]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easygaussiansplatting_b200.gsfunction import Camera, GSFunction, GSFunctionFused  # noqa: E402
from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
F = GSFunction if (len(sys.argv) > 2 and sys.argv[2] == "ops") else GSFunctionFused
N, W, H = 1_000_000, 1920, 1080
dev = "cuda:0"
sc = synthetic_scene(N, W, H, sh_dim=48, seed=0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cam = Camera(W, H, sc["fx"], sc["fy"], sc["cx"], sc["cy"], T(sc["Rcw"]), T(sc["tcw"]), T(sc["twc"]))
P = {k: T(sc[k]).requires_grad_() for k in ("pws", "shs", "scales", "rots")}
al = T(sc["alphas"][:, None]).requires_grad_()
us0 = torch.zeros((N, 2), device=dev, requires_grad=True)
dl = T(upstream_gradient(W, H, 0) * (3.0 * W * H))
for _ in range(steps):
    for p in list(P.values()) + [al]:
        p.grad = None
    image, _ = F.apply(P["pws"], P["shs"], al, P["scales"], P["rots"], us0, cam)
    image.backward(dl)
torch.cuda.synchronize()
print("done", float(image.mean()))

#!/usr/bin/env python
"""BASELINE config 3 stand-in: optimise a random-initialised Gaussian set against images of a
hidden ground-truth scene for a few hundred Adam steps through the B200 rasterizer.

The reference's train.py (train.py:30-83) cannot run on the GPU box (the reference tree is not
there), so this reproduces its loop with the same parameterisation (raw alpha through a
sigmoid, raw scale through exp, quaternion normalised: gsplat/utils.py:121-151,
gsmodel.py:186-212) and Adam groups (gsmodel.py:117-127), an L1 loss, no densification.
Prints the loss curve; returns (first, last) loss.

usage: train_synthetic.py [--n 20000] [--views 4] [--iters 200] [--size 320x240] [--ops]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easygaussiansplatting_b200.gsfunction import Camera, GSFunction, GSFunctionFused  # noqa: E402
from easygaussiansplatting_b200.scene import ring_camera, synthetic_scene  # noqa: E402


def train(n=20000, views=4, iters=200, W=320, H=240, use_ops=False, seed=0, verbose=True, dev="cuda:0"):
    F = GSFunction if use_ops else GSFunctionFused
    gt = synthetic_scene(n, W, H, sh_dim=48, seed=seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cams = []
    for k in range(views):
        Rcw, tcw, twc = ring_camera(k, views, radius=0.3)
        cams.append(Camera(W, H, gt["fx"], gt["fy"], gt["cx"], gt["cy"], T(Rcw), T(tcw), T(twc)))
    us0 = torch.zeros((n, 2), device=dev)
    with torch.no_grad():
        targets = [F.apply(T(gt["pws"]), T(gt["shs"]), T(gt["alphas"][:, None]), T(gt["scales"]), T(gt["rots"]),
                           us0, c)[0].clone() for c in cams]
    # initial guess: jittered positions, grey colours, half opacity, slightly wrong sizes
    rng = np.random.default_rng(seed + 1)
    pws = torch.nn.Parameter(T(gt["pws"] + rng.normal(scale=0.01, size=gt["pws"].shape).astype(np.float32)))
    low = torch.nn.Parameter(torch.zeros((n, 3), device=dev))
    high = torch.nn.Parameter(torch.full((n, 45), 0.001, device=dev))
    alphas_raw = torch.nn.Parameter(torch.zeros((n, 1), device=dev))                      # sigmoid -> 0.5
    scales_raw = torch.nn.Parameter(torch.log(T(gt["scales"] * 1.3)))
    rots_raw = torch.nn.Parameter(T(gt["rots"] + rng.normal(scale=0.05, size=gt["rots"].shape).astype(np.float32)))
    opt = torch.optim.Adam([{"params": [pws], "lr": 0.001}, {"params": [low], "lr": 0.02},
                            {"params": [high], "lr": 0.001}, {"params": [alphas_raw], "lr": 0.05},
                            {"params": [scales_raw], "lr": 0.005}, {"params": [rots_raw], "lr": 0.001}], eps=1e-15)
    us = torch.zeros((n, 2), device=dev, requires_grad=True)
    losses = []
    for it in range(iters):
        k = it % views
        shs = torch.cat((low, high), dim=1)
        image, mask = F.apply(pws, shs, torch.sigmoid(alphas_raw), torch.exp(scales_raw),
                              torch.nn.functional.normalize(rots_raw), us, cams[k])
        loss = torch.nn.functional.l1_loss(image, targets[k])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        if verbose and (it % 20 == 0 or it == iters - 1):
            print("iter %4d  view %d  L1 %.5f" % (it, k, losses[-1]), flush=True)
    first = float(np.mean(losses[:views]))
    last = float(np.mean(losses[-views:]))
    return first, last, losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--size", default="320x240")
    ap.add_argument("--ops", action="store_true", help="seven-operator surface instead of the fused path")
    a = ap.parse_args()
    W, H = map(int, a.size.split("x"))
    first, last, _ = train(a.n, a.views, a.iters, W, H, a.ops)
    print("L1 first %.5f -> last %.5f (x%.2f)" % (first, last, first / max(last, 1e-12)))

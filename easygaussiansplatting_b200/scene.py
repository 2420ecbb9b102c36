"""Deterministic synthetic 3DGS scenes (SURVEY.md 8d / BASELINE.md 3).

There is no dataset on the GPU box, so every benchmark and full-size parity test renders a
seeded random scene of the shape BASELINE.json names.  numpy only; the arrays follow the
reference's Gaussian record (gsplat/gau_io.py:7-12: pw[3], rot[4] (w,x,y,z), scale[3],
alpha, sh[sh_dim] laid out [coef][rgb]) and its Camera fields
(gsplat/gausplat_dataset.py:14-27).
"""
import math

import numpy as np


def synthetic_scene(N, width, height, sh_dim=48, seed=0):
    """camera Rcw=I, tcw=0, fx=fy=W/(2 tan30deg), cx=W/2, cy=H/2.  Per Gaussian:
    z~U(2,12); target pixel u~U(-0.05W,1.05W), v~U(-0.05H,1.05H); pw back-projected;
    per-axis pixel sigma ~ LogU(0.5,4) -> scale = sigma*z/fx; rot = normalised N(0,1)^4;
    alpha~U(0.05,0.95); sh[:, :3]~N(0,1), higher bands ~N(0,0.2^2)."""
    rng = np.random.default_rng(seed)
    W, H = int(width), int(height)
    fx = fy = W / (2.0 * math.tan(math.radians(30.0)))
    cx, cy = W / 2.0, H / 2.0
    z = rng.uniform(2.0, 12.0, N)
    u = rng.uniform(-0.05 * W, 1.05 * W, N)
    v = rng.uniform(-0.05 * H, 1.05 * H, N)
    pws = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], axis=1)
    sig = np.exp(rng.uniform(math.log(0.5), math.log(4.0), (N, 3)))
    scales = sig * z[:, None] / fx
    rots = rng.normal(size=(N, 4))
    rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    alphas = rng.uniform(0.05, 0.95, N)
    shs = rng.normal(size=(N, sh_dim))
    shs[:, 3:] *= 0.2
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(pws=f32(pws), rots=f32(rots), scales=f32(scales), alphas=f32(alphas),
                shs=f32(shs), Rcw=np.eye(3, dtype=np.float32), tcw=np.zeros(3, np.float32),
                twc=np.zeros(3, np.float32), fx=float(fx), fy=float(fy), cx=float(cx),
                cy=float(cy), width=W, height=H)


def upstream_gradient(width, height, seed=0):
    """dloss_dgammas ~ N(0,1)/(3WH), planar [3,H,W] float32."""
    rng = np.random.default_rng(seed + 1000003)
    g = rng.normal(size=(3, int(height), int(width))) / (3.0 * width * height)
    return np.ascontiguousarray(g, dtype=np.float32)


def ring_camera(k, n_views, radius=0.6):
    """View k of n on a small ring around the synthetic camera (config 5: one camera per
    rank over shared Gaussians).  Returns Rcw, tcw, twc float32."""
    a = 2.0 * math.pi * k / max(1, n_views)
    yaw = 0.05 * math.sin(a)
    c, s = math.cos(yaw), math.sin(yaw)
    Rcw = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)
    twc = np.array([radius * math.cos(a), radius * math.sin(a) * 0.5, 0.0])
    tcw = -Rcw @ twc
    return Rcw.astype(np.float32), tcw.astype(np.float32), twc.astype(np.float32)

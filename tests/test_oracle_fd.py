"""BASELINE config 2's "numerical check" (SURVEY 8d): the analytic gradient of the WHOLE chain
(parameters -> per-Gaussian stages -> tile rasterizer -> image -> scalar loss), as the fp64 oracle
computes it (splat_backward + Jacobian chain), against central finite differences of the oracle's
own forward for >= 64 randomly chosen parameters -- the reference does the same on its 4-Gaussian
example (backward_cpu.py:47-58, 545-698) with a forward difference.  CPU only."""
import numpy as np

from oracle import oracle as orc
from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient


def _forward(sc, W, H, dl):
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    us, pcs, depths, Ju = orc.project(sc["pws"], sc["Rcw"], sc["tcw"], sc["fx"], sc["fy"], sc["cx"], sc["cy"])
    d32 = f32(depths)
    c3, J3r, J3s = orc.compute_cov3d(sc["rots"], sc["scales"], d32)
    c2, J2c, J2p = orc.compute_cov2d(c3, pcs, sc["Rcw"], d32, sc["fx"], sc["fy"], W, H)
    col, Jcs, Jcp = orc.sh2color(sc["shs"], sc["pws"], sc["twc"])
    ci, areas, Jci = orc.inverse_cov2d(c2, d32)
    fwd = orc.splat(H, W, us, ci, sc["alphas"], d32, col, areas)
    loss = float((fwd["image"] * dl).sum())
    return loss, (us, ci, col, fwd, Ju, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci)


def test_full_chain_gradient_vs_central_differences():
    N, W, H = 160, 64, 48
    sc = synthetic_scene(N, W, H, sh_dim=12, seed=11)
    sc = {k: (np.asarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    dl = upstream_gradient(W, H, 3).astype(np.float64) * (3.0 * W * H)
    loss, (us, ci, col, fwd, Ju, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci) = _forward(sc, W, H, dl)
    g4 = orc.splat_backward(H, W, us, ci, sc["alphas"], col, fwd, dl)
    chain = orc.chain_backward(sc["Rcw"], *g4, Ju, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci)
    grads = {"pws": chain["pws"], "scales": chain["scales"], "rots": chain["rots"], "shs": chain["shs"],
             "alphas": np.asarray(g4[2]).reshape(-1)}
    rng = np.random.default_rng(5)
    # only Gaussians the rasterizer actually used can have a non-zero gradient
    used = np.unique(fwd["gsid"])
    picks, good, checked = [], 0, 0
    while len(picks) < 80:
        name = ("pws", "scales", "rots", "shs", "alphas")[rng.integers(5)]
        i = int(rng.choice(used))
        j = 0 if name == "alphas" else int(rng.integers(sc[name].shape[1]))
        picks.append((name, i, j))
    for name, i, j in picks:
        x = sc[name]
        idx = (i,) if name == "alphas" else (i, j)
        x0 = float(x[idx])
        an = float(grads[name][i] if name == "alphas" else grads[name].reshape(N, -1)[i, j])
        checked += 1
        # The forward is only piecewise smooth: a perturbation that moves a pixel across the
        # alpha' < 0.002 skip, the tau < 1e-4 stop, the 0.99 clamp or a tile-rectangle boundary
        # (ceil(3 sigma) radii) adds a jump, and every oracle stage rounds its inputs to fp32 like
        # the GPU operators, which bounds the step from below.  So: three step sizes, a parameter
        # passes when one of them reproduces the analytic value.
        for rel in (2e-3, 5e-4, 1.5e-4):
            h = rel * max(abs(x0), 0.05)
            x[idx] = x0 + h
            lp = _forward(sc, W, H, dl)[0]
            x[idx] = x0 - h
            lm = _forward(sc, W, H, dl)[0]
            x[idx] = x0
            fd = (lp - lm) / (2 * h)
            if abs(fd - an) <= 1e-2 * max(abs(fd), abs(an)) + 1e-6 * abs(loss):
                good += 1
                break
    assert checked >= 64 and good >= int(0.85 * checked), "only %d of %d finite differences agree" % (good, checked)

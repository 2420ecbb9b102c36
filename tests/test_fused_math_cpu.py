"""CPU check of the fused per-Gaussian math (csrc/pg_fused_math.h, compiled for the host by
tests/host_shim/fused_host.cpp): its forward equals the five oracle stages and its analytic
vector-Jacobian products equal the oracle's materialised-Jacobian chain (gsmodel.py:72-85)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from easygaussiansplatting_b200.scene import synthetic_scene

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host_shim", "_build", "libfused_host.so")


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "host_shim", "fused_host.cpp")
    hdr = os.path.join(HERE, "..", "easygaussiansplatting_b200", "csrc", "pg_fused_math.h")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, src], check=True)
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scene(N, k3, seed, rotate=True):
    sc = synthetic_scene(N, 320, 240, sh_dim=3 * k3, seed=seed)
    rng = np.random.default_rng(seed + 5)
    if rotate:  # a general camera, so that Rcw really takes part in the chain
        a, b = 0.2, -0.1
        Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        sc["Rcw"] = (Ry @ Rx).astype(np.float32)
        sc["tcw"] = np.array([0.3, -0.2, 0.5], np.float32)
        sc["twc"] = (-(sc["Rcw"].astype(np.float64).T @ sc["tcw"].astype(np.float64))).astype(np.float32)
    sc["rots"] = (sc["rots"] * rng.uniform(0.7, 1.4, (N, 1))).astype(np.float32)  # un-normalised q
    idx = rng.choice(N, size=max(1, N // 40), replace=False)
    sc["pws"][idx, 2] = -1.0   # culled
    wide = rng.choice(N, size=max(1, N // 40), replace=False)
    sc["pws"][wide, 0] *= 5.0  # fov clamp active
    return sc


def _fwd(shim, sc, k3, W, H):
    N = len(sc["pws"])
    us = np.empty((N, 2), np.float32); ci = np.empty((N, 3), np.float32); col = np.empty((N, 3), np.float32)
    d = np.empty(N, np.float32); ar = np.empty((N, 2), np.int32)
    rc = shim.fused_forward_host(N, k3, _p(sc["pws"]), _p(sc["rots"]), _p(sc["scales"]), _p(sc["shs"]),
                                 _p(sc["Rcw"]), _p(sc["tcw"]), _p(sc["twc"]), C.c_float(sc["fx"]),
                                 C.c_float(sc["fy"]), C.c_float(sc["cx"]), C.c_float(sc["cy"]), C.c_float(W),
                                 C.c_float(H), _p(us), _p(ci), _p(col), _p(d), _p(ar))
    assert rc == 0
    return us, ci, col, d, ar


def _oracle_stages(sc, W, H):
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    us, pcs, depths, Ju = orc.project(sc["pws"], sc["Rcw"], sc["tcw"], sc["fx"], sc["fy"], sc["cx"], sc["cy"])
    d32 = f32(depths)
    c3, J3r, J3s = orc.compute_cov3d(sc["rots"], sc["scales"], d32)
    c2, J2c, J2p = orc.compute_cov2d(f32(c3), f32(pcs), sc["Rcw"], d32, sc["fx"], sc["fy"], W, H)
    col, Jcs, Jcp = orc.sh2color(sc["shs"], sc["pws"], sc["twc"])
    ci, areas, Jci = orc.inverse_cov2d(f32(c2), d32)
    return dict(us=us, depths=depths, ci=ci, col=col, areas=areas, c2=c2,
                J=(Ju, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci))


def _close(got, ref, name, rtol=3e-5, atol_scale=6e-6):
    got = np.asarray(got, np.float64).reshape(ref.shape)
    scale = max(np.abs(ref).max(), 1e-30)
    excess = np.abs(got - ref) - (rtol * np.abs(ref) + atol_scale * scale)
    i = np.argmax(excess)
    assert excess.flat[i] <= 0, "%s: got %r want %r (scale %g)" % (name, got.flat[i], ref.flat[i], scale)


@pytest.mark.parametrize("k3", [1, 4, 9, 16])
def test_fused_forward_matches_oracle_stages(shim, k3):
    W, H = 320, 240
    sc = _scene(3000, k3, seed=k3)
    us, ci, col, d, ar = _fwd(shim, sc, k3, W, H)
    o = _oracle_stages(sc, W, H)
    assert np.array_equal(d < 0, o["depths"] < 0)
    _close(us, o["us"], "us"); _close(d, o["depths"], "depths"); _close(col, o["col"], "colors")
    _close(ci, o["ci"], "cinv2ds", rtol=1e-4, atol_scale=2e-5)
    # radius from the fused fp32 cov2d vs the oracle's fp64 cov2d: identical except where
    # 3*sqrt() sits within fp32 rounding of an integer
    assert np.mean(ar != o["areas"]) < 2e-3 and np.abs(ar - o["areas"]).max() <= 1
    assert not us[d < 0].any() and not ci[d < 0].any() and not ar[d < 0].any()


@pytest.mark.parametrize("k3", [1, 4, 9, 16])
def test_fused_backward_matches_jacobian_chain(shim, k3):
    W, H = 320, 240
    N = 3000
    sc = _scene(N, k3, seed=10 + k3)
    o = _oracle_stages(sc, W, H)
    rng = np.random.default_rng(k3)
    gu = rng.normal(size=(N, 1, 2)).astype(np.float32)
    gci = rng.normal(size=(N, 1, 3)).astype(np.float32) * 10
    gcol = rng.normal(size=(N, 1, 3)).astype(np.float32)
    gal = rng.normal(size=(N, 1, 1)).astype(np.float32)
    keep = o["depths"] > 0
    gu[~keep] = 0; gci[~keep] = 0; gcol[~keep] = 0  # culled Gaussians have no patches
    ref = orc.chain_backward(sc["Rcw"], gu, gci, gal, gcol, *o["J"])
    gpw = np.empty((N, 3), np.float32); gsh = np.empty((N, 3 * k3), np.float32)
    gs = np.empty((N, 3), np.float32); gq = np.empty((N, 4), np.float32)
    rc = shim.fused_backward_host(N, k3, _p(sc["pws"]), _p(sc["rots"]), _p(sc["scales"]), _p(sc["shs"]),
                                  _p(sc["Rcw"]), _p(sc["tcw"]), _p(sc["twc"]), C.c_float(sc["fx"]),
                                  C.c_float(sc["fy"]), C.c_float(sc["cx"]), C.c_float(sc["cy"]), C.c_float(W),
                                  C.c_float(H), _p(gu), _p(gci), _p(gcol), _p(gpw), _p(gsh), _p(gs), _p(gq))
    assert rc == 0
    # normalised per tensor (fp32 recomputation vs fp64 chain over fp32-rounded intermediates)
    for got, name in ((gpw, "pws"), (gsh, "shs"), (gs, "scales"), (gq, "rots")):
        e = np.abs(got.astype(np.float64) - ref[name]).max() / np.abs(ref[name]).max()
        assert e < 2e-5, (name, e)
    # and elementwise on well-conditioned rows
    _close(gsh, ref["shs"], "dshs", rtol=1e-5, atol_scale=1e-6)


def test_fused_backward_zero_grads_and_culls(shim):
    k3, N, W, H = 16, 64, 320, 240
    sc = _scene(N, k3, seed=99)
    sc["pws"][:8, 2] = 0.1
    z2 = np.zeros((N, 2), np.float32); z3 = np.zeros((N, 3), np.float32)
    gpw = np.full((N, 3), 7, np.float32); gsh = np.full((N, 48), 7, np.float32)
    gs = np.full((N, 3), 7, np.float32); gq = np.full((N, 4), 7, np.float32)
    shim.fused_backward_host(N, k3, _p(sc["pws"]), _p(sc["rots"]), _p(sc["scales"]), _p(sc["shs"]), _p(sc["Rcw"]),
                             _p(sc["tcw"]), _p(sc["twc"]), C.c_float(sc["fx"]), C.c_float(sc["fy"]),
                             C.c_float(sc["cx"]), C.c_float(sc["cy"]), C.c_float(W), C.c_float(H), _p(z2), _p(z3),
                             _p(z3), _p(gpw), _p(gsh), _p(gs), _p(gq))
    assert not gpw.any() and not gsh.any() and not gs.any() and not gq.any()

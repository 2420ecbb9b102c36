"""Pins the CPU oracle (oracle/gs_oracle.c) to the reference's own Python: every function
is checked against fixtures produced by backward_cpu.py / gsplat/gausplat.py
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


def close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.max(np.abs(a - b)) if a.size else 0.0
    assert err <= tol, err


@pytest.fixture(scope="module")
def st():
    return dict(np.load(os.path.join(G, "stages.npz")))


@pytest.fixture(scope="module")
def bl():
    return dict(np.load(os.path.join(G, "blend.npz")))


def test_project(st):
    fx, fy, cx, cy, W, H = st["cam"]
    us, pcs, depths, J = orc.project(st["pws"], st["Rcw"], st["tcw"], fx, fy, cx, cy)
    close(us, st["us"], 1e-9); close(pcs, st["pcs"], 1e-12)
    close(depths, st["pcs"][:, 2], 1e-12); close(J, st["du_dpcs"], 1e-9)


def test_project_cull():
    pws = np.array([[0, 0, 0.1], [0, 0, 0.2], [0, 0, 0.25], [0, 0, -3]], dtype=np.float32)
    us, pcs, depths, J = orc.project(pws, np.eye(3), np.zeros(3), 10, 10, 5, 5)
    assert depths[0] == -1 and depths[3] == -1 and depths[2] > 0.2
    # float32(0.2) is above the double 0.2 the reference compares against -> kept
    assert depths[1] > 0
    assert np.all(us[0] == 0) and np.all(pcs[3] == 0) and np.all(J[0] == 0)


def test_cov3d(st):
    d = np.ones(len(st["rots"]), dtype=np.float32)
    cov, Jr, Js = orc.compute_cov3d(st["rots"], st["scales"], d)
    close(cov, st["cov3ds"], 1e-12); close(Jr, st["dcov3d_drots"], 1e-12)
    close(Js, st["dcov3d_dscales"], 1e-12)
    d[3] = -1
    cov, Jr, Js = orc.compute_cov3d(st["rots"], st["scales"], d)
    assert np.all(cov[3] == 0) and np.all(Jr[3] == 0) and np.all(Js[3] == 0)


def test_cov2d(st):
    fx, fy, cx, cy, W, H = st["cam"]
    d = np.ones(len(st["rots"]), dtype=np.float32)
    cov, Jc, Jp, cl = orc.compute_cov2d(st["cov3ds"], st["pcs"], st["Rcw"], d, fx, fy, W, H,
                                        return_clamped=True)
    assert not cl.any()
    close(cov, st["cov2ds"], 1e-9); close(Jc, st["dcov2d_dcov3ds"], 1e-9)
    close(Jp, st["dcov2d_dpcs"], 1e-9)


def test_cov2d_clamp_matches_device_rule():
    # kernel.cu:458-461: x/z clamped to 1.3*W/(2fx) and the clamped x enters J.
    pcs = np.array([[9.0, 0.1, 1.0]], dtype=np.float32)
    cov3 = np.array([[0.01, 0, 0, 0.01, 0, 0.01]], dtype=np.float32)
    d = np.ones(1, dtype=np.float32)
    cov, _, _, cl = orc.compute_cov2d(cov3, pcs, np.eye(3), d, 50, 50, 100, 100,
                                      return_clamped=True)
    assert cl[0]
    x = 1.3 * 100 / (2 * 50)
    J = np.array([[50, 0, -50 * x], [0, 50, -50 * 0.1]])
    ref = J @ (0.01 * np.eye(3)) @ J.T
    close(cov[0], [ref[0, 0] + 0.3, ref[0, 1], ref[1, 1] + 0.3], 1e-4)  # fp32 inputs


def test_sh2color(st):
    col, Js, Jp = orc.sh2color(st["shs"], st["pws"], st["twc"])
    close(col, st["colors"], 1e-9); close(Js, st["dcolor_dshs"], 1e-9)
    close(Jp, st["dcolor_dpws"], 1e-9)
    for k in (1, 4, 9):
        col = orc.sh2color(st["shs"][:, : 3 * k], st["pws"], st["twc"], calc_J=False)[0]
        close(col, st["colors_k%d" % k], 1e-9)


def test_inverse_cov2d(st):
    d = np.ones(len(st["rots"]), dtype=np.float32)
    cinv, areas, J = orc.inverse_cov2d(st["cov2ds"], d)
    c32 = st["cov2ds"].astype(np.float32).astype(np.float64)
    close(cinv, st["cinv2ds"], 1e-9); close(J, st["dcinv2d_dcov2ds"], 1e-9)
    assert np.array_equal(areas, np.ceil(3 * np.sqrt(c32[:, [0, 2]].astype(np.float32))).astype(np.int32))
    # NaN determinant -> depth culled in place (kernel.cu:301-305)
    bad = np.array([[np.inf, np.inf, np.inf], [1, 0, 1]], dtype=np.float32)  # inf-inf
    d = np.ones(2, dtype=np.float32)
    cinv, areas, J = orc.inverse_cov2d(bad, d)
    assert d[0] == -1 and d[1] == 1 and np.all(cinv[0] == 0) and np.all(areas[0] == 0)


def _splat_blend(bl):
    fx, fy, cx, cy, W, H = bl["cam"]
    W, H = int(W), int(H)
    d = bl["pws"][:, 2].astype(np.float32).copy()
    cov2 = bl["cov2ds"].astype(np.float32)
    areas = np.ascontiguousarray(np.ceil(3 * np.sqrt(cov2[:, [0, 2]])).astype(np.int32))
    fwd = orc.splat(H, W, bl["us32"], bl["cinv32"], bl["alphas"], d, bl["colors32"], areas)
    return H, W, fwd


def test_splat_matches_calc_gamma(bl):
    """tile rasterizer oracle == backward_cpu.get_image / calc_gamma on the fixture scene"""
    H, W, fwd = _splat_blend(bl)
    assert not fwd["ambiguous"].any()
    close(fwd["image"].transpose(1, 2, 0), bl["image"], 1e-12)
    # the device's contrib is 1 + index within the TILE's list (kernel.cu:239,251);
    # backward_cpu's is 1 + index within the whole (depth-sorted) array: map through gsid.
    gx = (W + 15) // 16
    for y in range(H):
        for x in range(W):
            c = fwd["contrib"][y, x]
            start = fwd["ranges"][(y // 16) * gx + x // 16, 0]
            assert c > 0 and fwd["gsid"][start + c - 1] + 1 == bl["contrib"][y, x]
    # 10 Gaussians sorted by depth, every large one covers both tiles
    assert fwd["P"] == len(fwd["gsid"]) and np.all(np.diff(fwd["keys"].astype(np.int64)) > 0)


def test_splat_backward_matches_calc_loss(bl):
    H, W, fwd = _splat_blend(bl)
    du, dc, da, dcol = orc.splat_backward(H, W, bl["us32"], bl["cinv32"], bl["alphas"],
                                          bl["colors32"], fwd, bl["dloss_dgammas"])
    # dloss_dgammas crosses the op boundary as float32 -> ~1e-7 relative
    for got, name in ((du, "dloss_dus"), (dc, "dloss_dcinv2ds"), (da, "dloss_dalphas"),
                      (dcol, "dloss_dcolors")):
        close(got, bl[name], 1e-6 * np.max(np.abs(bl[name])))


def test_full_chain_matches_backward(bl):
    """params -> stages -> splat -> splatB -> Jacobian chain == backward_cpu.backward().
    Intermediates cross op boundaries as float32 (as on the device), hence 2e-5."""
    fx, fy, cx, cy, W, H = bl["cam"]
    W, H = int(W), int(H)
    twc = np.zeros(3, np.float32)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    us, pcs, depths, du_dpcs = orc.project(bl["pws"], bl["Rcw"], bl["tcw"], fx, fy, cx, cy)
    d32 = f32(depths)
    cov3, J3r, J3s = orc.compute_cov3d(bl["rots"], bl["scales"], d32)
    cov2, J2c, J2p = orc.compute_cov2d(f32(cov3), f32(pcs), bl["Rcw"], d32, fx, fy, W, H)
    col, Jcs, Jcp = orc.sh2color(bl["shs"], bl["pws"], twc)
    cinv, areas, Jci = orc.inverse_cov2d(f32(cov2), d32)
    close(us, bl["us"], 1e-6); close(cinv, bl["cinv2ds"], 1e-5); close(col, bl["colors"], 1e-6)
    fwd = orc.splat(H, W, f32(us), f32(cinv), bl["alphas"], d32, f32(col), areas)
    image = fwd["image"].transpose(1, 2, 0)
    dl = (np.sign(image - bl["image_gt"]) / image.size).transpose(2, 0, 1)  # L1Loss grad
    assert abs(np.mean(np.abs(image - bl["image_gt"])) - float(bl["chain_loss"][0])) < 1e-6
    du, dc, da, dcol = orc.splat_backward(H, W, f32(us), f32(cinv), bl["alphas"], f32(col),
                                          fwd, f32(dl))
    g = orc.chain_backward(bl["Rcw"], du, dc, da, dcol, du_dpcs, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci)
    g_np = orc.chain_backward(bl["Rcw"], du, dc, da, dcol, du_dpcs, J3r, J3s, J2c, J2p, Jcs, Jcp, Jci,
                              use_numpy=True)
    for name in g:  # C chain (orc_chain) == numpy transcription of gsmodel.py:72-85
        close(g[name], g_np[name].reshape(g[name].shape), 1e-12 * max(1.0, np.abs(g_np[name]).max()))
    for name, ref in (("rots", "chain_drots"), ("scales", "chain_dscales"), ("shs", "chain_dshs"),
                      ("alphas", "chain_dalphas"), ("pws", "chain_dpws")):
        scale = max(1e-12, np.max(np.abs(bl[ref])))
        assert np.max(np.abs(g[name] - bl[ref])) / scale < 2e-5, name


def test_forward_cpu_renderer():
    fc = dict(np.load(os.path.join(G, "fwdcpu.npz")))
    img = orc.forward_cpu_splat(int(fc["H"]), int(fc["W"]), fc["us"], fc["cinv2ds"],
                                _scene_alphas(fc), fc["depths"], fc["colors"], fc["areas"])
    close(img, fc["image"], 2e-6)


def _scene_alphas(fc):
    from easygaussiansplatting_b200.scene import synthetic_scene
    sc = synthetic_scene(int(fc["N"]), int(fc["W"]), int(fc["H"]), sh_dim=int(fc["sh_dim"]),
                         seed=int(fc["seed"]))
    return sc["alphas"]


def test_tile_oracle_close_to_forward_cpu():
    """forward_cpu.py agrees with the tile rasterizer only at image level
    (different clamp, radius, thresholds, footprint)."""
    from easygaussiansplatting_b200.scene import synthetic_scene
    fc = dict(np.load(os.path.join(G, "fwdcpu.npz")))
    W, H = int(fc["W"]), int(fc["H"])
    sc = synthetic_scene(int(fc["N"]), W, H, sh_dim=int(fc["sh_dim"]), seed=int(fc["seed"]))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    us, pcs, depths = orc.project(sc["pws"], sc["Rcw"], sc["tcw"], sc["fx"], sc["fy"], sc["cx"],
                                  sc["cy"], calc_J=False)
    d32 = f32(depths)
    cov3, = orc.compute_cov3d(sc["rots"], sc["scales"], d32, calc_J=False)
    cov2, = orc.compute_cov2d(f32(cov3), f32(pcs), sc["Rcw"], d32, sc["fx"], sc["fy"], W, H,
                              calc_J=False)
    col, = orc.sh2color(sc["shs"], sc["pws"], sc["twc"], calc_J=False)
    cinv, areas = orc.inverse_cov2d(f32(cov2), d32, calc_J=False)
    close(cov2, fc["cov2ds"], 1e-3)
    fwd = orc.splat(H, W, f32(us), f32(cinv), sc["alphas"], d32, f32(col), areas)
    img = fwd["image"].transpose(1, 2, 0)
    diff = np.abs(img - fc["image"])
    assert np.mean(diff) < 1.5e-2 and np.quantile(diff, 0.99) < 8e-2, (np.mean(diff), np.quantile(diff, 0.99))


def test_bin_edge_cases():
    H, W = 40, 50  # 4x3 tiles, ragged right/bottom edge
    us = np.array([[-100, -100], [25, 20], [49.5, 39.5], [25, 20], [1e9, 5]], dtype=np.float32)
    areas = np.array([[3, 3], [2, 2], [40, 1], [0, 0], [5, 5]], dtype=np.int32)
    depths = np.array([1, 2, 3, 0.1, 5], dtype=np.float32)
    cinv = np.tile(np.array([[0.5, 0, 0.5]], dtype=np.float32), (5, 1))
    col = np.ones((5, 3), dtype=np.float32)
    al = np.full(5, 0.5, dtype=np.float32)
    fwd = orc.splat(H, W, us, cinv, al, depths, col, areas)
    # 0: off-grid -> culled in place; 3: depth < 0.2 stays as is; 4: off the right edge
    assert depths[0] == -1 and np.all(areas[0] == 0) and depths[3] == np.float32(0.1)
    assert depths[4] == -1
    # 1 covers tile (1,1) only:  (25-2)/16=1.43->1, (25+2+15)/16=2.6->2
    # 2 covers x tiles 0..3 on tile row 2
    assert fwd["P"] == 1 + 4
    T = fwd["ranges"]
    assert tuple(T[1 * 4 + 1]) == (0, 1)
    assert [tuple(T[2 * 4 + x]) for x in range(4)] == [(1, 2), (2, 3), (3, 4), (4, 5)]
    assert fwd["final_tau"][0, 0] == 0 and fwd["contrib"][0, 0] == 0  # empty tile untouched


def test_empty_scene():
    fwd = orc.splat(16, 16, np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32),
                    np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 3), np.float32),
                    np.zeros((0, 2), np.int32))
    assert fwd["P"] == 0 and not fwd["image"].any()

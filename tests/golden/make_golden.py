"""Generates the golden fixtures in tests/golden/ by IMPORTING THE REFERENCE'S OWN PYTHON
(backward_cpu.py and gsplat/gausplat.py from /root/reference) in the dev container.

The reference cannot travel to the GPU box, so its outputs are committed here as small
.npz files together with this script.  Run:  python tests/golden/make_golden.py

Fixtures
  stages.npz   per-Gaussian stages + analytic Jacobians from backward_cpu.py
               (transform/project :68-87, compute_cov_3d :90-151, compute_cov_2d :154-190,
               sh2color :278-385, calc_cinv2d :200-212) on 40 random Gaussians, SH deg 3
               (plus deg 0/1/2 colours).
  blend.npz    per-pixel compositing forward/backward from backward_cpu.py
               (get_image :398-405, calc_loss :408-437) and the full chain `backward`
               (:440-499) on a 10-Gaussian, 32x16 scene built so that the tile rasterizer
               and the whole-image CPU loop coincide (see SURVEY 8a "Divergences").
  fwdcpu.npz   the forward_cpu.py pipeline (gsplat/gausplat.py project/compute_cov_3d/
               compute_cov_2d/sh2color/inverse_cov2d/splat) on 3000 Gaussians at 160x120.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, REF)
sys.modules.setdefault("gsplatcu", types.ModuleType("gsplatcu"))  # gsplat/utils.py:2

import backward_cpu as bc  # noqa: E402
import gsplat.gausplat as gp  # noqa: E402

RCW = np.array([[0.89699204, 0.06525223, 0.43720409],
                [-0.04508268, 0.99739184, -0.05636552],
                [-0.43974177, 0.03084909, 0.89759429]]).T  # backward_cpu.py:518-520
TCW = np.array([1.03796196, 0.42017467, 4.67804612])       # backward_cpu.py:517


def f32(a):
    return np.asarray(a, dtype=np.float32)


def make_stages():
    rng = np.random.default_rng(1234)
    N, sh_dim = 40, 48
    Rcw, tcw = f32(RCW), f32(TCW)
    twc = f32(np.linalg.inv(Rcw.astype(np.float64)) @ (-tcw.astype(np.float64)))
    fx, fy, cx, cy, W, H = 40.0, 42.0, 32.0, 24.0, 64, 48
    pws = f32(rng.uniform(-1.0, 1.0, (N, 3)))
    rots = f32(rng.normal(size=(N, 4)))
    rots[: N // 2] /= np.linalg.norm(rots[: N // 2], axis=1, keepdims=True)  # half un-normalised
    scales = f32(np.exp(rng.uniform(np.log(0.02), np.log(0.4), (N, 3))))
    shs = f32(rng.normal(size=(N, sh_dim)) * 0.5)
    out = dict(pws=pws, rots=rots, scales=scales, shs=shs, Rcw=Rcw, tcw=tcw, twc=twc,
               cam=np.array([fx, fy, cx, cy, W, H], dtype=np.float64))
    R64, t64, twc64 = Rcw.astype(np.float64), tcw.astype(np.float64), twc.astype(np.float64)
    us = np.zeros((N, 2)); pcs = np.zeros((N, 3)); du_dpcs = np.zeros((N, 2, 3))
    cov3ds = np.zeros((N, 6)); J3r = np.zeros((N, 6, 4)); J3s = np.zeros((N, 6, 3))
    cov2ds = np.zeros((N, 3)); J2c = np.zeros((N, 3, 6)); J2p = np.zeros((N, 3, 3))
    colors = np.zeros((N, 3)); Jcs = np.zeros((N, 1, 16)); Jcp = np.zeros((N, 3, 3))
    cinv = np.zeros((N, 3)); Jci = np.zeros((N, 3, 3))
    colors_k = {k: np.zeros((N, 3)) for k in (1, 4, 9)}
    for i in range(N):
        pw, q, s, sh = (pws[i].astype(np.float64), rots[i].astype(np.float64),
                        scales[i].astype(np.float64), shs[i].astype(np.float64))
        pcs[i], _ = bc.transform(pw, R64, t64, True)
        us[i], du_dpcs[i] = bc.project(pcs[i], fx, fy, cx, cy, True)
        cov3ds[i], J3r[i], J3s[i] = bc.compute_cov_3d(q, s, True)
        # the device consumes fp32 intermediates: round them exactly as the op boundary does
        c3 = f32(cov3ds[i]).astype(np.float64)
        pc = f32(pcs[i]).astype(np.float64)
        cov2ds[i], J2c[i], J2p[i] = bc.compute_cov_2d(c3, pc, R64, fx, fy, True)
        colors[i], Jcs[i], Jcp[i] = bc.sh2color(sh, pw, twc64, True)
        for k in colors_k:
            colors_k[k][i] = bc.sh2color(sh[: 3 * k].copy(), pw, twc64, False)
        c2 = f32(cov2ds[i]).astype(np.float64)
        cinv[i], Jci[i] = bc.calc_cinv2d(c2, True)
    assert np.all(pcs[:, 2] > 0.2)
    assert np.all(np.abs(pcs[:, 0] / pcs[:, 2]) < 1.3 * W / (2 * fx))  # clamp inactive
    assert np.all(np.abs(pcs[:, 1] / pcs[:, 2]) < 1.3 * H / (2 * fy))
    out.update(us=us, pcs=pcs, du_dpcs=du_dpcs, cov3ds=cov3ds, dcov3d_drots=J3r,
               dcov3d_dscales=J3s, cov2ds=cov2ds, dcov2d_dcov3ds=J2c, dcov2d_dpcs=J2p,
               colors=colors, dcolor_dshs=Jcs, dcolor_dpws=Jcp, cinv2ds=cinv,
               dcinv2d_dcov2ds=Jci, colors_k1=colors_k[1], colors_k4=colors_k[4],
               colors_k9=colors_k[9])
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **out)
    print("stages.npz", N)


def make_blend():
    """10 Gaussians in front of an identity camera, 32x16 px (2x1 tiles).
    5 large ones (radius >= 32 px -> their tile rect is the whole image) with alpha in
    [0.8, 0.95] (drives tau below 1e-4 on many pixels), 5 small ones with alpha <= 0.15
    (alpha' >= 0.002 only inside 3 sigma -> inside their tile rect).  Sorted by depth with
    distinct millimetres, all alpha' < 0.99, clamp cone inactive."""
    rng = np.random.default_rng(7)
    W, H = 32, 16
    fx = fy = 16.0
    cx, cy = W / 2.0, H / 2.0
    Rcw, tcw = np.eye(3), np.zeros(3)
    N, sh_dim = 10, 48
    z = np.sort(rng.uniform(2.0, 6.0, N))
    assert len(set((z * 1000).astype(int))) == N
    big = np.array([1, 0, 1, 0, 1, 0, 1, 1, 0, 0], dtype=bool)
    upx = np.stack([rng.uniform(4, W - 4, N), rng.uniform(3, H - 3, N)], 1)
    pws = np.stack([(upx[:, 0] - cx) * z / fx, (upx[:, 1] - cy) * z / fy, z], 1)
    sig_px = np.where(big[:, None], rng.uniform(11, 16, (N, 3)), rng.uniform(0.8, 2.0, (N, 3)))
    scales = sig_px * z[:, None] / fx
    rots = rng.normal(size=(N, 4)); rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    alphas = np.where(big, rng.uniform(0.8, 0.95, N), rng.uniform(0.05, 0.15, N))
    shs = rng.normal(size=(N, sh_dim)) * 0.3
    shs[:, :3] = rng.uniform(-1.0, 1.5, (N, 3))
    pws, scales, rots, alphas, shs = map(f32, (pws, scales, rots, alphas, shs))
    image_gt = rng.uniform(0, 1, (H, W, 3))
    p64 = lambda a: a.astype(np.float64)
    bc.sh_dim = sh_dim  # backward() reads this module global (backward_cpu.py:454,487)
    loss, drots, dscales, dshs, dalphas, dpws = bc.backward(
        p64(rots), p64(scales), p64(shs), p64(alphas), p64(pws), Rcw, tcw, fx, fy, cx, cy,
        image_gt, True)
    # stage values on the same scene (fp64, no rounding of intermediates)
    colors = np.zeros((N, 3)); us = np.zeros((N, 2)); cinv = np.zeros((N, 3))
    cov2ds = np.zeros((N, 3))
    twc = np.linalg.inv(Rcw) @ (-tcw)
    for i in range(N):
        pc = bc.transform(p64(pws[i]), Rcw, tcw)
        us[i] = bc.project(pc, fx, fy, cx, cy)
        c3 = bc.compute_cov_3d(p64(rots[i]), p64(scales[i]))
        cov2ds[i] = bc.compute_cov_2d(c3, pc, Rcw, fx, fy)
        colors[i] = bc.sh2color(p64(shs[i]), p64(pws[i]), twc)
        cinv[i] = bc.calc_cinv2d(cov2ds[i])
    # the splat-level fixture uses the fp32-rounded op inputs, exactly what splat() receives
    us32, cinv32, col32 = f32(us), f32(cinv), f32(colors)
    a64 = p64(alphas)
    image = bc.get_image(a64, p64(cinv32).reshape(-1), p64(col32).reshape(-1),
                         p64(us32).reshape(-1), H, W)
    _, dloss_dgammas = bc.get_loss(image, image_gt)
    _, dl_dalphas, dl_dcinv, dl_dcolors, dl_dus = bc.calc_loss(
        a64, p64(cinv32).reshape(-1), p64(col32).reshape(-1), p64(us32).reshape(-1),
        image_gt, True)
    contrib = np.zeros((H, W), dtype=np.int32)
    for y in range(H):
        for x in range(W):
            contrib[y, x] = bc.calc_gamma(a64, p64(cinv32).reshape(-1), p64(col32).reshape(-1),
                                          p64(us32).reshape(-1), np.array([x, y]), True)[-1]
    # the scene must exercise both thresholds
    assert contrib.min() < N and contrib.max() >= 5
    radius = np.ceil(3 * np.sqrt(cov2ds[:, [0, 2]]))
    assert np.all(radius[big] >= 32), radius
    np.savez_compressed(
        os.path.join(HERE, "blend.npz"),
        pws=pws, rots=rots, scales=scales, alphas=alphas, shs=shs, Rcw=f32(Rcw), tcw=f32(tcw),
        cam=np.array([fx, fy, cx, cy, W, H]), image_gt=image_gt,
        us=us, cov2ds=cov2ds, cinv2ds=cinv, colors=colors,
        us32=us32, cinv32=cinv32, colors32=col32,
        image=image, contrib=contrib, dloss_dgammas=dloss_dgammas,
        dloss_dalphas=dl_dalphas.reshape(N, 1, 1), dloss_dcinv2ds=dl_dcinv.reshape(N, 1, 3),
        dloss_dcolors=dl_dcolors.reshape(N, 1, 3), dloss_dus=dl_dus.reshape(N, 1, 2),
        chain_loss=loss, chain_drots=drots.reshape(N, 4), chain_dscales=dscales.reshape(N, 3),
        chain_dshs=dshs.reshape(N, sh_dim), chain_dalphas=dalphas.reshape(N, 1),
        chain_dpws=dpws.reshape(N, 3))
    print("blend.npz contrib range", contrib.min(), contrib.max())


def make_fwdcpu():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from easygaussiansplatting_b200.scene import synthetic_scene
    W, H, N = 160, 120, 3000
    sc = synthetic_scene(N, W, H, sh_dim=12, seed=3)
    p64 = lambda a: np.asarray(a, dtype=np.float64)
    Rcw, tcw = p64(sc["Rcw"]), p64(sc["tcw"])
    fx, fy, cx, cy = sc["fx"], sc["fy"], sc["cx"], sc["cy"]
    twc = np.linalg.inv(Rcw) @ (-tcw)
    us, pcs = gp.project(p64(sc["pws"]), Rcw, tcw, fx, fy, cx, cy)   # forward_cpu.py:43
    depths = pcs[:, 2]
    cov3ds = gp.compute_cov_3d(p64(sc["scales"]), p64(sc["rots"]))   # :48
    cov2ds = gp.compute_cov_2d(pcs, fx, fy, W, H, cov3ds, Rcw)       # :51
    colors = gp.sh2color(p64(sc["shs"]), p64(sc["pws"]), twc)        # :54
    cinv2ds, areas = gp.inverse_cov2d(cov2ds)                        # :57
    image = gp.splat(H, W, us, cinv2ds, p64(sc["alphas"]), depths, colors, areas)  # :59
    np.savez_compressed(os.path.join(HERE, "fwdcpu.npz"), N=N, W=W, H=H, seed=3, sh_dim=12,
                        us=us, depths=depths, cov2ds=cov2ds, colors=colors, cinv2ds=cinv2ds,
                        areas=areas, image=image.astype(np.float32))
    print("fwdcpu.npz", image.shape, float(image.mean()))


if __name__ == "__main__":
    make_stages()
    make_blend()
    make_fwdcpu()

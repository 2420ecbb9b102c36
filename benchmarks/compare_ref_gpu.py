#!/usr/bin/env python
"""B-REF-GPU vs OURS: the reference's UNMODIFIED `gsplatcu` CUDA extension
(built for sm_100a into baseline/_ref/ by baseline/build_ref_gpu.sh) and this repository's
kernels, driven by the SAME autograd wrapper (gsfunction.build_gsfunction) on the SAME
synthetic scenes on one B200.  Each arm runs in its own process because both modules are
called `gsplatcu`.

    python benchmarks/compare_ref_gpu.py            # runs both arms, prints + writes JSON
    python benchmarks/compare_ref_gpu.py --arm ref  # one arm (internal)

t_fwd = GSFunction.apply (6 ops, calc_J=True); t_bwd = image.backward (splatB + Jacobian
chain); CUDA events, 3 warm-ups, median of `--iters`; also per-op splat / splatB times.

    python benchmarks/compare_ref_gpu.py --three-way   # error table of the rasterizer pair
Three-way error table (splat + splatB on IDENTICAL fp32 op inputs, config 2): ours <-> fp64 oracle,
reference-GPU <-> oracle, ours <-> reference-GPU, and -- when a GSB_EXACT_MATH variant library
libgsplat_b200_exact.so exists -- ours-with-IEEE-exp2/rcp <-> oracle; each with and without the
pixels / Gaussians the oracle flags as ambiguous (an alpha' within 2e-5 of the 0.002 threshold).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = [("config2", 1_000_000, 1920, 1080, 48), ("cfg4_50k_512", 50_000, 512, 512, 48),
           ("cfg4_500k_1080p", 500_000, 1920, 1080, 48)]


def run_arm(arm, iters, out_path, configs):
    import torch
    if arm == "ref":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import gsplatcu as gsc  # the reference's compiled extension
        assert gsc.__file__.endswith(".so"), gsc.__file__
        sys.path.insert(1, ROOT)
    else:
        sys.path.insert(0, ROOT)
        import gsplatcu as gsc
    from easygaussiansplatting_b200.gsfunction import Camera, build_gsfunction
    from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient
    GSF = build_gsfunction(gsc)
    dev = "cuda:0"
    res = {}
    for name, N, W, H, sh_dim in configs:
        sc = synthetic_scene(N, W, H, sh_dim=sh_dim, seed=0)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        cam = Camera(W, H, sc["fx"], sc["fy"], sc["cx"], sc["cy"], T(sc["Rcw"]), T(sc["tcw"]), T(sc["twc"]))
        P = {k: T(sc[k]).requires_grad_() for k in ("pws", "shs", "scales", "rots")}
        al = T(sc["alphas"][:, None]).requires_grad_()
        us0 = torch.zeros((N, 2), device=dev, requires_grad=True)
        dl = T(upstream_gradient(W, H, 0) * (3.0 * W * H))
        leaves = [P["pws"], P["shs"], al, P["scales"], P["rots"]]
        tf, tb, ts, tsb = [], [], [], []
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for it in range(3 + iters):
            for p in leaves:
                p.grad = None
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            image, mask = GSF.apply(P["pws"], P["shs"], al, P["scales"], P["rots"], us0, cam)
            e1.record()
            image.backward(dl)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
        # per-op: splat and splatB alone on the same inputs
        with torch.no_grad():
            us, pcs, depths = gsc.project(P["pws"], cam.Rcw, cam.tcw, cam.fx, cam.fy, cam.cx, cam.cy, False)
            c3 = gsc.computeCov3D(P["rots"], P["scales"], depths, False)[0]
            c2 = gsc.computeCov2D(c3, pcs, cam.Rcw, depths, cam.fx, cam.fy, W, H, False)[0]
            col = gsc.sh2Color(P["shs"], P["pws"], cam.twc, False)[0]
            ci, areas = gsc.inverseCov2D(c2, depths, False)
            for it in range(3 + iters):
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                o = gsc.splat(H, W, us, ci, al, depths, col, areas)
                e1.record()
                g = gsc.splatB(H, W, us, ci, al, depths, col, o[1], o[2], o[3], o[4], dl)
                e2.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1)); tsb.append(e1.elapsed_time(e2))
        med = statistics.median
        res[name] = dict(N=N, W=W, H=H, P=int(o[4].numel()), t_fwd_ms=med(tf), t_bwd_ms=med(tb),
                         splat_ms=med(ts), splatB_ms=med(tsb),
                         mpix_per_s=W * H / ((med(tf) + med(tb)) * 1e-3) / 1e6)
        np.savez(out_path + "." + name + ".npz", image=image.detach().cpu().numpy(),
                 **{"g_" + k: v.grad.cpu().numpy() for k, v in P.items()}, g_alphas=al.grad.cpu().numpy(),
                 dus=g[0].cpu().numpy(), dcinv=g[1].cpu().numpy(), dalpha=g[2].cpu().numpy(),
                 dcol=g[3].cpu().numpy())
        del P, al, us0, dl, image, o, g
        torch.cuda.empty_cache()
    json.dump(res, open(out_path, "w"))


def three_way_arm(arm, inp_path, out_path):
    """splat + splatB of one implementation on the shared op inputs"""
    import torch
    if arm == "ref":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import gsplatcu as gsc
        assert gsc.__file__.endswith(".so"), gsc.__file__
    else:
        sys.path.insert(0, ROOT)
        import gsplatcu as gsc
    d = np.load(inp_path)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    H, W = int(d["H"]), int(d["W"])
    us, ci, al, dep, col, ar, dl = (T(d[k]) for k in ("us", "cinv2ds", "alphas", "depths", "colors", "areas", "dl"))
    o = gsc.splat(H, W, us, ci, al, dep, col, ar)
    g = gsc.splatB(H, W, us, ci, al, dep, col, o[1], o[2], o[3], o[4], dl)
    torch.cuda.synchronize()
    np.savez(out_path, image=o[0].cpu().numpy(), dus=g[0].cpu().numpy().reshape(-1, 2),
             dcinv=g[1].cpu().numpy().reshape(-1, 3), dalpha=g[2].cpu().numpy().reshape(-1, 1),
             dcol=g[3].cpu().numpy().reshape(-1, 3))


def three_way(out):
    """-> dict; writes out + '.three_way.json'"""
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient
    N, W, H = 1_000_000, 1920, 1080
    orc.set_num_threads(os.cpu_count() or 1)
    sc = synthetic_scene(N, W, H, sh_dim=48, seed=0)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    us, pcs, depths, _ = orc.project(sc["pws"], sc["Rcw"], sc["tcw"], sc["fx"], sc["fy"], sc["cx"], sc["cy"])
    d32 = f32(depths)
    c3 = orc.compute_cov3d(sc["rots"], sc["scales"], d32, calc_J=False)[0]
    c2 = orc.compute_cov2d(f32(c3), f32(pcs), sc["Rcw"], d32, sc["fx"], sc["fy"], W, H, calc_J=False)[0]
    col = orc.sh2color(sc["shs"], sc["pws"], sc["twc"], calc_J=False)[0]
    ci, areas = orc.inverse_cov2d(f32(c2), d32, calc_J=False)[:2]
    dl = upstream_gradient(W, H, 0) * (3.0 * W * H)
    inp = out + ".three_way_inputs.npz"
    np.savez(inp, H=H, W=W, us=f32(us), cinv2ds=f32(ci), alphas=f32(sc["alphas"]), depths=d32, colors=f32(col),
             areas=np.ascontiguousarray(areas, dtype=np.int32), dl=f32(dl))
    arms = {"ours": {}}
    exact = os.path.join(ROOT, "easygaussiansplatting_b200", "libgsplat_b200_exact.so")
    if os.path.exists(exact):
        arms["ours_exact"] = {"GSB_LIB": exact}
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and any(f.startswith("gsplatcu") and f.endswith(".so") for f in os.listdir(ref_dir)):
        arms["ref"] = {}
    res = {}
    for arm, env in arms.items():
        o = out + ".three_way_%s.npz" % arm
        subprocess.run([sys.executable, os.path.abspath(__file__), "--three-way-arm", "ref" if arm == "ref" else "ours",
                        "--inputs", inp, "--out", o], check=True, env=dict(os.environ, **env))
        res[arm] = dict(np.load(o))
        os.remove(o)
    # fp64 oracle on the same fp32 inputs (the arms' in-place culls do not change depths/areas here:
    # the inputs are already the post-cull values the oracle's own stages produce)
    fwd = orc.splat(H, W, f32(us), f32(ci), sc["alphas"], d32.copy(), f32(col), np.ascontiguousarray(areas, dtype=np.int32))
    du, dc, da, dcol, amb = orc.splat_backward(H, W, f32(us), f32(ci), sc["alphas"], f32(col), fwd, f32(dl),
                                               return_ambiguous=True)
    res["oracle"] = {"image": fwd["image"], "dus": du.reshape(-1, 2), "dcinv": dc.reshape(-1, 3),
                     "dalpha": da.reshape(-1, 1), "dcol": dcol.reshape(-1, 3)}
    okpix, okg = ~fwd["ambiguous"], ~amb

    def err(a, b, name):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        scale = max(np.abs(b).max(), 1e-30)
        if name == "image":
            e = np.abs(a - b).max(axis=0)
            return float(e.max() / scale), float(e[okpix].max() / scale)
        e = np.abs(a - b).max(axis=1)
        return float(e.max() / scale), float(e[okg].max() / scale)
    table = {}
    pairs = [(a, "oracle") for a in arms] + ([("ours", "ref")] if "ref" in arms else [])
    for a, b in pairs:
        table["%s_vs_%s" % (a, b)] = {k: dict(zip(("all", "without_ambiguous"), err(res[a][k], res[b][k], k)))
                                      for k in ("image", "dus", "dcinv", "dalpha", "dcol")}
    outj = {"what": "max |a - b| / max |b| per tensor, splat + splatB on identical fp32 op inputs, config 2 "
                    "(1M Gaussians, 1920x1080)", "ambiguous_pixels": int(fwd["ambiguous"].sum()),
            "ambiguous_gaussians": int(amb.sum()), "table": table}
    os.remove(inp)
    json.dump(outj, open(out + ".three_way.json", "w"), indent=1)
    print(json.dumps(outj, indent=1))
    return outj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--three-way", action="store_true")
    ap.add_argument("--three-way-arm", default=None)
    ap.add_argument("--inputs", default=None)
    ap.add_argument("--arm", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "compare_ref_gpu"))
    ap.add_argument("--configs", default="config2,cfg4_50k_512,cfg4_500k_1080p")
    a = ap.parse_args()
    if a.three_way_arm:
        return three_way_arm(a.three_way_arm, a.inputs, a.out)
    if a.three_way:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        return three_way(a.out)
    cfgs = [c for c in CONFIGS if c[0] in a.configs.split(",")]
    if a.arm:
        return run_arm(a.arm, a.iters, a.out + "." + a.arm + ".json", cfgs)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    have_ref = any(f.endswith(".so") for f in os.listdir(os.path.join(ROOT, "baseline", "_ref"))) \
        if os.path.isdir(os.path.join(ROOT, "baseline", "_ref")) else False
    arms = ["ours"] + (["ref"] if have_ref else [])
    for arm in arms:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--arm", arm, "--iters", str(a.iters),
                        "--out", a.out, "--configs", a.configs], check=True)
    out = {arm: json.load(open(a.out + "." + arm + ".json")) for arm in arms}
    if have_ref:
        out["speedup_fwd_bwd"] = {k: out["ref"][k]["mpix_per_s"] and out["ours"][k]["mpix_per_s"] / out["ref"][k]["mpix_per_s"]
                                  for k in out["ours"]}
        out["speedup_splat_plus_splatB"] = {
            k: (out["ref"][k]["splat_ms"] + out["ref"][k]["splatB_ms"]) / (out["ours"][k]["splat_ms"] + out["ours"][k]["splatB_ms"])
            for k in out["ours"]}
        par = {}
        for k in out["ours"]:
            A, B = np.load(a.out + ".ours.json." + k + ".npz"), np.load(a.out + ".ref.json." + k + ".npz")
            par[k] = {name: float(np.abs(A[name] - B[name]).max() / max(np.abs(B[name]).max(), 1e-30)) for name in A.files}
        out["ours_vs_ref_normalised_max_diff"] = par
    else:
        out["note"] = "baseline/_ref/gsplatcu*.so missing: reference arm skipped"
    for arm in arms:  # the raw tensors are scratch (hundreds of MB)
        for k in list(out[arm]):
            f = a.out + "." + arm + ".json." + k + ".npz"
            if os.path.exists(f):
                os.remove(f)
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

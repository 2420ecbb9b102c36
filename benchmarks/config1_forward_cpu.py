#!/usr/bin/env python
"""BASELINE config 1: the reference's CPU forward (`forward_cpu.py`) on 10k synthetic Gaussians,
256x256, SH degree 0 -- timed on this host's cores, next to the same scene through our GPU path.

The CPU side is the reference's OWN code when its pure-Python package is installed under
baseline/_ref/py (baseline/build_ref_gpu.sh py; git-ignored, travels with the gpurun snapshot):
the call sequence of forward_cpu.py:43-60 over gsplat/gausplat.py (project, compute_cov_3d,
compute_cov_2d, sh2color, inverse_cov2d, splat with im=None).  Without it the C restatement of
that renderer in oracle/ is timed instead (kind "port").  The reference's loop is single-threaded
NumPy; `nproc` is reported because the contract asks for it.

    python benchmarks/config1_forward_cpu.py            # prints one JSON object
Used by bench.py (`cpu_baseline.config1`)."""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N1, W1, H1, SH1 = 10_000, 256, 256, 3


def _reference_modules():
    """gsplat.gausplat of the unmodified reference, or None."""
    py = os.path.join(ROOT, "baseline", "_ref", "py")
    if not os.path.exists(os.path.join(py, "gsplat", "gausplat.py")):
        return None
    for p in (os.path.join(ROOT, "tests", "shims"), py):  # matplotlib / plyfile stand-ins, then the package
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import gsplat.gausplat as gp  # imports matplotlib.pyplot (shim) only
        return gp
    except Exception:  # e.g. the package needs gsplatcu -> libgsplat_b200.so missing
        return None


def run(repeats=3, gpu=True):
    from easygaussiansplatting_b200.scene import synthetic_scene
    sc = synthetic_scene(N1, W1, H1, sh_dim=SH1, seed=0)
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    pws, rots, scales, shs, alphas = map(f64, (sc["pws"], sc["rots"], sc["scales"], sc["shs"], sc["alphas"]))
    Rcw, tcw, twc = f64(sc["Rcw"]), f64(sc["tcw"]), f64(sc["twc"])
    fx, fy, cx, cy = sc["fx"], sc["fy"], sc["cx"], sc["cy"]
    gp = _reference_modules()
    times, image = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        if gp is not None:
            with contextlib.redirect_stdout(io.StringIO()):  # the renderer prints its progress
                us, pcs = gp.project(pws, Rcw, tcw, fx, fy, cx, cy)
                depths = pcs[:, 2]
                cov3ds = gp.compute_cov_3d(scales, rots)
                cov2ds = gp.compute_cov_2d(pcs, fx, fy, W1, H1, cov3ds, Rcw)
                colors = gp.sh2color(shs, pws, twc)
                cinv2ds, areas = gp.inverse_cov2d(cov2ds)
                image = gp.splat(H1, W1, us, cinv2ds, alphas, depths, colors, areas)
        else:
            from oracle import oracle as orc
            us, pcs, depths, _ = orc.project(sc["pws"], sc["Rcw"], sc["tcw"], fx, fy, cx, cy)
            d32 = np.ascontiguousarray(depths, dtype=np.float32)
            c3 = orc.compute_cov3d(sc["rots"], sc["scales"], d32, calc_J=False)[0]
            c2 = orc.compute_cov2d(np.float32(c3), np.float32(pcs), sc["Rcw"], d32, fx, fy, W1, H1, calc_J=False)[0]
            colors = orc.sh2color(sc["shs"], sc["pws"], sc["twc"], calc_J=False)[0]
            ci, areas = orc.inverse_cov2d(np.float32(c2), d32, calc_J=False)[:2]
            image = orc.forward_cpu_splat(H1, W1, us, ci, sc["alphas"], depths, colors, areas)
        times.append(time.perf_counter() - t0)
    sec = min(times)
    out = {"what": "config 1: forward_cpu.py pipeline, 10k synthetic Gaussians, 256x256, SH deg 0, forward only",
           "kind": "reference" if gp is not None else "port", "seconds": sec, "value": W1 * H1 / sec / 1e6,
           "unit": "Mpixels/s", "gaussians_per_s": N1 / sec, "cores": 1, "nproc": os.cpu_count(),
           "note": "the reference renderer is a single-threaded NumPy loop over depth-sorted Gaussians "
                   "(gausplat.py:197-238)"}
    if gpu:
        try:
            import torch
            if torch.cuda.is_available():
                from easygaussiansplatting_b200 import ops
                dev = "cuda:0"
                T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
                args = (T(sc["pws"]), T(sc["rots"]), T(sc["scales"]), T(sc["shs"]), T(sc["Rcw"]), T(sc["tcw"]),
                        T(sc["twc"]), fx, fy, cx, cy, W1, H1)
                al = T(sc["alphas"])

                def fwd():
                    u, ci, col, dep, ar = ops.preprocess(*args)
                    return ops.splat(H1, W1, u, ci, al, dep, col, ar)[0]
                for _ in range(3):
                    img = fwd()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    img = fwd()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                ours = img.permute(1, 2, 0).cpu().numpy().astype(np.float64)
                mse = float(np.mean((ours - image) ** 2))
                out["ours_gpu"] = {"ms": ms, "value": W1 * H1 / (ms * 1e-3) / 1e6, "unit": "Mpixels/s",
                                   "psnr_vs_cpu_db": float(10 * np.log10(1.0 / max(mse, 1e-30))),
                                   "note": "loose check only: forward_cpu.py uses pixel-rectangle footprints, int() "
                                           "radii, det + 1e-6 and no alpha' < 0.002 skip (SURVEY 8a divergences)"}
        except Exception as e:  # the CPU number stands on its own
            out["ours_gpu"] = {"error": repr(e)[:200]}
    return out


if __name__ == "__main__":
    print(json.dumps(run()))

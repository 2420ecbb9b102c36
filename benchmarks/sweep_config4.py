#!/usr/bin/env python
"""BASELINE config 4: tile-occupancy sweep.  N in {50k, 500k, 5M} x {512^2, 1080p, 4K}:
fused forward+backward ms/step, Mpix/s, Gaussians/s, per-kernel CUDA-event times and the HBM
roofline fraction of the two rasterizer kernels on their algorithmic bytes.
Writes gpurun_out/sweep_config4.json.   usage: sweep_config4.py [--iters 10] [--quick]"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easygaussiansplatting_b200 import _lib, ops  # noqa: E402
from easygaussiansplatting_b200.gsfunction import Camera, GSFunctionFused  # noqa: E402
from easygaussiansplatting_b200.scene import synthetic_scene, upstream_gradient  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    lib = _lib.load()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    Ns = [50_000, 500_000] if a.quick else [50_000, 500_000, 5_000_000]
    sizes = [(512, 512), (1920, 1080)] if a.quick else [(512, 512), (1920, 1080), (3840, 2160)]
    out = []
    for N in Ns:
        for W, H in sizes:
            sc = synthetic_scene(N, W, H, sh_dim=48, seed=0)
            T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
            cam = Camera(W, H, sc["fx"], sc["fy"], sc["cx"], sc["cy"], T(sc["Rcw"]), T(sc["tcw"]), T(sc["twc"]))
            P = {k: T(sc[k]).requires_grad_() for k in ("pws", "shs", "scales", "rots")}
            al = T(sc["alphas"][:, None]).requires_grad_()
            us0 = torch.zeros((N, 2), device=dev, requires_grad=True)
            dl = T(upstream_gradient(W, H, 0) * (3.0 * W * H))

            def step():
                for p in list(P.values()) + [al]:
                    p.grad = None
                image, _ = GSFunctionFused.apply(P["pws"], P["shs"], al, P["scales"], P["rots"], us0, cam)
                image.backward(dl)
            ms = []
            for it in range(3 + a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); step(); e1.record(); torch.cuda.synchronize()
                if it >= 3:
                    ms.append(e0.elapsed_time(e1))
            lib.gsb_profile_enable(1)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            lib.gsb_profile_enable(0)
            kern = {}
            for i in range(lib.gsb_profile_kernels()):
                tot, cnt = C.c_double(0), C.c_longlong(0)
                lib.gsb_profile_read(i, C.byref(tot), C.byref(cnt))
                if cnt.value:
                    kern[lib.gsb_profile_kernel_name(i).decode()] = tot.value / 3
            with torch.no_grad():
                us, ci, col, d, ar = ops.preprocess(P["pws"], P["rots"], P["scales"], P["shs"], cam.Rcw, cam.tcw,
                                                    cam.twc, cam.fx, cam.fy, cam.cx, cam.cy, W, H)
                o = ops.splat(H, W, us, ci, al, d, col, ar)
                npatch = o[4].numel()
                gy, gx = (H + 15) // 16, (W + 15) // 16
                pad = torch.zeros((gy * 16, gx * 16), dtype=torch.int32, device=dev)
                pad[:H, :W] = o[1]
                p_eff = int(pad.view(gy, 16, gx, 16).amax(dim=(1, 3)).sum().item())
            t = statistics.median(ms)
            # the same step replayed as two CUDA graphs (graphed.GraphedFusedStep): no launch / wrapper cost
            t_graph = None
            try:
                from easygaussiansplatting_b200.graphed import GraphedFusedStep
                gstep = GraphedFusedStep(P["pws"], P["shs"], al, P["scales"], P["rots"], cam)

                def gs():
                    gstep.forward()
                    gstep.dloss_dimage.copy_(dl)
                    gstep.backward()
                msg = []
                for it in range(3 + a.iters):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); gs(); e1.record(); torch.cuda.synchronize()
                    if it >= 3:
                        msg.append(e0.elapsed_time(e1))
                t_graph = statistics.median(msg)
                del gstep
            except Exception as e:  # noqa: BLE001 - the eager numbers stand on their own
                t_graph = "failed: %r" % (e,)
            alg_f = 44 * p_eff + 8 * gx * gy + 20 * W * H
            alg_b = alg_f + 36 * N
            row = dict(N=N, W=W, H=H, patches=npatch, patches_per_tile=npatch / (gx * gy), p_eff=p_eff,
                       ms_per_step=t, ms_per_step_cuda_graphs=t_graph, mpix_per_s=W * H / (t * 1e-3) / 1e6,
                       gaussians_per_s=N / (t * 1e-3),
                       kernel_ms=kern,
                       draw_hbm_frac=alg_f / (kern["draw"] * 1e-3) / 1e9 / hbm,
                       draw_backward_hbm_frac=alg_b / (kern["draw_backward"] * 1e-3) / 1e9 / hbm,
                       draw_GBps=alg_f / (kern["draw"] * 1e-3) / 1e9,
                       draw_backward_GBps=alg_b / (kern["draw_backward"] * 1e-3) / 1e9)
            out.append(row)
            print(json.dumps(row), flush=True)
            del P, al, us0, dl
            ops.clear_record_cache()
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(hbm_gbs=hbm, rows=out), open(os.path.join(ROOT, "gpurun_out", "sweep_config4.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE config 3: the reference's UNMODIFIED train.py, run once per `gsplatcu`
implementation on the same synthetic COLMAP directory with the same seeds:

  ref      reference train.py + reference gsplat/*.py + the reference's own gsplatcu (CUDA,
           built for sm_100a by baseline/build_ref_gpu.sh)
  ours     reference train.py + reference gsplat/*.py + THIS repository's gsplatcu package
           (the drop-in: nothing else changes)
  ours_n   as `ours`, plus the N1/N2/N3 replacements patched into the reference's modules
           before train.py starts: GSFunction -> fused path, gau_loss -> fused L1+D-SSIM,
           GSModel.update_density_info / update_gaussian_density / reset_alpha -> density.py

and compares what train.py prints: the per-epoch avg_loss (train.py:69) and the density
reports (gsmodel.py:308-315).  100 epochs x --views iterations each (train.py:40).

The reference's Python is read from baseline/_ref/py (installed there, git-ignored, by
baseline/build_ref_gpu.sh); without it this script reports that and exits 0.  matplotlib,
plyfile and faiss are not installed in the image: tests/shims provides import stand-ins
(train.py only draws a preview window with matplotlib).

usage: train_reference.py [--n 20000] [--views 8] [--size 320x240] [--impl ref,ours,ours_n]
"""
import argparse
import json
import os
import re
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFPY = os.path.join(ROOT, "baseline", "_ref", "py")
REFSO = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "tests", "shims")

CHILD = r"""
import os, sys, runpy, time
impl, root, refso, refpy, shims, ds = sys.argv[1:7]
sys.path[:0] = [shims, refpy] + ([refso] if impl == "ref" else [root])
import numpy as np, torch
torch.manual_seed(0); np.random.seed(0)
import gsplatcu
print("GSPLATCU", getattr(gsplatcu, "__file__", "?"), flush=True)
if impl == "ours_n":
    import gsplat.gsmodel as gm, gsplat.pytorch_ssim as ps
    from easygaussiansplatting_b200.gsfunction import GSFunctionFused
    from easygaussiansplatting_b200.loss import gau_loss
    from easygaussiansplatting_b200.density import DensityController
    gm.GSFunction = GSFunctionFused
    ps.gau_loss = gau_loss
    _init = gm.GSModel.__init__
    def init(self, sense_size, max_steps):
        _init(self, sense_size, max_steps)
        self._ctl = DensityController(sense_size)
    def info(self):
        self._ctl.update_density_info(self.us.grad, self.mask)
        del self.us.grad, self.mask
    gm.GSModel.__init__ = init
    gm.GSModel.update_density_info = info
    gm.GSModel.update_gaussian_density = lambda self, params, opt: self._ctl.update_gaussian_density(params, opt)
    gm.GSModel.reset_alpha = lambda self, params, opt: self._ctl.reset_alpha(params, opt)
sys.argv = ["train.py", "--path", ds]
torch.cuda.synchronize(); t0 = time.time()
runpy.run_path(os.path.join(refpy, "train.py"), run_name="__main__")
torch.cuda.synchronize()
print("TRAIN_SECONDS %.3f" % (time.time() - t0), flush=True)
"""


def write_colmap(ds, cams, W, H, fx, fy, cx, cy, gs):
    os.makedirs(os.path.join(ds, "sparse", "0"), exist_ok=True)
    os.makedirs(os.path.join(ds, "images"), exist_ok=True)
    with open(os.path.join(ds, "sparse", "0", "cameras.bin"), "wb") as f:   # read_write_model.py:99-131
        f.write(struct.pack("<Q", 1))
        f.write(struct.pack("<iiQQ", 1, 1, W, H) + struct.pack("<4d", fx, fy, cx, cy))    # PINHOLE
    from easygaussiansplatting_b200.gau_io import matrix_to_quaternion
    with open(os.path.join(ds, "sparse", "0", "images.bin"), "wb") as f:    # read_write_model.py:134-181
        f.write(struct.pack("<Q", len(cams)))
        for i, (Rcw, tcw) in enumerate(cams):
            q = matrix_to_quaternion(Rcw[None].astype(np.float64))[0]
            f.write(struct.pack("<i7di", i + 1, *q, *tcw.astype(np.float64), 1))
            f.write(("%04d.png" % i).encode() + b"\x00" + struct.pack("<Q", 0))
    np.save(os.path.join(ds, "sparse", "0", "points3D.npy"), gs)


def build_dataset(ds, n, views, W, H, seed=0):
    """hidden ground-truth scene -> PNGs rendered with this repository's rasterizer, and a
    perturbed, colourless subset of it as the initial points3D.npy (sh_dim 3)"""
    import torch
    from PIL import Image
    sys.path.insert(0, ROOT)
    from easygaussiansplatting_b200.gsfunction import Camera, GSFunctionFused
    from easygaussiansplatting_b200.scene import ring_camera, synthetic_scene
    from easygaussiansplatting_b200.gau_io import gsdata_type
    gt = synthetic_scene(n, W, H, sh_dim=48, seed=seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cams = []
    us0 = torch.zeros((n, 2), device="cuda")
    for k in range(views):
        Rcw, tcw, twc = ring_camera(k, views, radius=1.0)
        cams.append((Rcw, tcw))
        cam = Camera(W, H, gt["fx"], gt["fy"], gt["cx"], gt["cy"], T(Rcw), T(tcw), T(twc))
        with torch.no_grad():
            img = GSFunctionFused.apply(T(gt["pws"]), T(gt["shs"]), T(gt["alphas"][:, None]), T(gt["scales"]),
                                        T(gt["rots"]), us0, cam)[0]
        arr = (img.clamp(0, 1).permute(1, 2, 0).cpu().numpy() * 255 + 0.5).astype(np.uint8)
        os.makedirs(os.path.join(ds, "images"), exist_ok=True)
        Image.fromarray(arr).save(os.path.join(ds, "images", "%04d.png" % k))
    rng = np.random.default_rng(seed + 1)
    gs = np.zeros(n, dtype=gsdata_type(3))
    gs["pw"] = gt["pws"] + rng.normal(scale=0.02, size=(n, 3))
    gs["rot"] = [1, 0, 0, 0]
    gs["scale"] = np.clip(gt["scales"].mean(axis=1, keepdims=True) * 1.2, 0.002, None).repeat(3, 1)
    gs["alpha"] = 0.5
    gs["sh"] = 0.0
    write_colmap(ds, cams, W, H, gt["fx"], gt["fy"], gt["cx"], gt["cy"], gs)


def run_impl(impl, ds, work):
    os.makedirs(os.path.join(work, impl, "data"), exist_ok=True)
    env = dict(os.environ, PYTHONPATH="", MPLBACKEND="Agg")
    t0 = time.time()
    r = subprocess.run([sys.executable, "-c", CHILD, impl, ROOT, REFSO, REFPY, SHIMS, ds], cwd=os.path.join(work, impl),
                       env=env, capture_output=True, text=True, timeout=3000)
    out = r.stdout
    res = dict(impl=impl, returncode=r.returncode, wall_s=round(time.time() - t0, 2),
               losses=[float(x) for x in re.findall(r"epoch:\d+ avg_loss:([0-9.eE+-]+)", out)],
               reports=[dict(pruned=int(a), cloned=int(b), splited=int(c), total=int(d)) for a, b, c, d in re.findall(
                   r"pruned num:\s+(\d+)\s+cloned num:\s+(\d+)\s+splited num:\s+(\d+)\s+total gaussian number:\s+(\d+)", out)],
               gsplatcu=(re.findall(r"GSPLATCU (\S+)", out) or ["?"])[0],
               train_s=float((re.findall(r"TRAIN_SECONDS ([0-9.]+)", out) or ["nan"])[0]))
    if r.returncode != 0:
        res["stderr_tail"] = r.stderr[-3000:]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--size", default="320x240")
    ap.add_argument("--impl", default="ref,ours,ours_n")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_reference.json"))
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REFPY, "train.py")):
        print(json.dumps({"unavailable": "reference python not installed under baseline/_ref/py "
                                         "(run baseline/build_ref_gpu.sh in the dev container)"}))
        return 0
    W, H = (int(x) for x in a.size.split("x"))
    work = tempfile.mkdtemp(prefix="trainref_")
    ds = os.path.join(work, "dataset")
    build_dataset(ds, a.n, a.views, W, H)
    results = {}
    for impl in a.impl.split(","):
        if impl == "ref" and not any(f.startswith("gsplatcu") and f.endswith(".so") for f in os.listdir(REFSO)):
            results[impl] = {"unavailable": "baseline/_ref/gsplatcu*.so not built"}
            continue
        results[impl] = run_impl(impl, ds, work)
        r = results[impl]
        print("%-7s rc=%d train %.1fs  first/last loss %s  reports %d  (%s)" % (
            impl, r["returncode"], r["train_s"], (r["losses"][:1] + r["losses"][-1:]), len(r["reports"]), r["gsplatcu"]),
            file=sys.stderr)
        if r["returncode"] != 0:
            print(r.get("stderr_tail", ""), file=sys.stderr)
    summary = dict(config=dict(n=a.n, views=a.views, width=W, height=H, epochs=100), results=results)
    base = results.get("ref") if results.get("ref", {}).get("losses") else None
    if base:
        for k, r in results.items():
            if k != "ref" and r.get("losses"):
                m = min(len(base["losses"]), len(r["losses"]))
                d = np.abs(np.array(base["losses"][:m]) - np.array(r["losses"][:m]))
                r["vs_ref"] = dict(max_abs_loss_diff_first5=float(d[:5].max()), max_abs_loss_diff=float(d.max()),
                                   final_loss_ratio=r["losses"][m - 1] / base["losses"][m - 1],
                                   first_report_equal=bool(r["reports"][:1] == base["reports"][:1]),
                                   speedup_train=base["train_s"] / r["train_s"])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ("train_s", "vs_ref", "returncode", "unavailable")}
                      for k, v in results.items()}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

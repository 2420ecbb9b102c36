#!/usr/bin/env python
"""Runs the reference's OWN `backward_gpu.py` and `forward_gpu.py`, unmodified, on top of a `gsplatcu`
module -- ours by default, the reference's compiled extension with --impl ref -- and reports
what they print.

backward_gpu.py (reference backward_gpu.py:81-152) is the reference's GPU-vs-CPU parity script: it
computes every stage with backward_cpu.py (fp64, per Gaussian) and prints one `[OK]` / `[NG]` line
(abs diff < 1e-4, backward_cpu.py:61-65) per operator output / Jacobian, for the image and for
splatB's four gradients: 19 lines.  forward_gpu.py:47-60 is the inference path (calc_J = False).
The scripts come from baseline/_ref/py (baseline/build_ref_gpu.sh py; git-ignored copies of the
reference files); matplotlib / plyfile are the import stubs of tests/shims.

    python benchmarks/run_reference_scripts.py [--impl ours|ref] [--out gpurun_out/reference_scripts.json]
"""
import argparse
import contextlib
import io
import json
import os
import re
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY = os.path.join(ROOT, "baseline", "_ref", "py")


def available():
    return all(os.path.exists(os.path.join(PY, f)) for f in ("backward_gpu.py", "forward_gpu.py", "backward_cpu.py"))


def run(impl="ours"):
    """-> {"backward_gpu": {"ok": n, "ng": n, "lines": [...]}, "forward_gpu": {...}}"""
    paths = [os.path.join(ROOT, "tests", "shims"), PY]
    paths.insert(0, os.path.join(ROOT, "baseline", "_ref") if impl == "ref" else ROOT)
    if impl == "ref":
        paths.append(ROOT)
    old_path, old_argv, old_cwd = list(sys.path), list(sys.argv), os.getcwd()
    sys.path[:0] = paths
    out = {}
    try:
        import gsplatcu
        out["gsplatcu"] = getattr(gsplatcu, "__file__", "?")
        import numpy as np
        np.random.seed(0)  # backward_gpu.py draws its 45 rest-SH coefficients unseeded
        os.chdir(PY)
        for script in ("backward_gpu.py", "forward_gpu.py"):
            sys.argv = [script]
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                g = runpy.run_path(os.path.join(PY, script), run_name="__main__")
            lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
            plain = [re.sub(r"\x1b\[[0-9;]*m", "", ln) for ln in lines]  # check() colours its verdicts
            rec = {"ok": sum(ln.startswith("[OK]") for ln in plain), "ng": sum(ln.startswith("[NG]") for ln in plain),
                   "lines": plain}
            if script == "forward_gpu.py":
                img = g["image"]
                rec["image_shape"] = list(img.shape)
                rec["image_mean"], rec["image_max"] = float(img.mean()), float(img.max())
            out[script[:-3]] = rec
    finally:
        sys.path[:], sys.argv[:] = old_path, old_argv
        os.chdir(old_cwd)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "ref"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_scripts.json"))
    a = ap.parse_args()
    if not available():
        print("baseline/_ref/py has no reference scripts (run baseline/build_ref_gpu.sh py where /root/reference exists)")
        sys.exit(0)
    res = run(a.impl)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out.replace(".json", "_%s.json" % a.impl), "w"), indent=1)
    print(json.dumps(res, indent=1))

"""Golden fixtures for SURVEY 8f row N3 (density control + Gaussian I/O), produced by RUNNING
THE REFERENCE'S OWN PYTHON in the dev container:

  density.npz  gsplat/gsmodel.py: GSModel.update_density_info (:219-234),
               update_gaussian_density (:236-318), reset_alpha (:320-331), with
               prune_params / update_params (:132-166) acting on a real torch.optim.Adam.
               The reference hard-codes device="cuda" in two places and draws the split
               offsets with torch.normal; a thin proxy around the `torch` module seen by
               gsmodel.py drops the device argument and records the unit normals z that
               torch.normal(mean, std) is defined by (out = z * std + mean), so the fixture
               is reproducible from (inputs, z).
  gsio.npz     gsplat/gau_io.py: load_ply (:60-107, through the plyfile stand-in in
               tests/shims), save_training_params (:138-153) and gsmodel.get_training_params
               (:95-129) on a small scene; the .ply bytes are stored in the fixture.

Run:  python tests/golden/make_golden_density.py
"""
import io
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", "shims"))
sys.path.insert(0, REF)
import types  # noqa: E402

sys.modules.setdefault("gsplatcu", types.ModuleType("gsplatcu"))  # gsplat/utils.py:2

import gsplat.gsmodel as gm  # noqa: E402
import gsplat.gau_io as gio  # noqa: E402


class TorchProxy:
    """`torch` as seen by gsplat/gsmodel.py, on a machine without a GPU"""

    def __init__(self):
        self.z_log = []
        self.gen = torch.Generator().manual_seed(77)

    def __getattr__(self, name):
        return getattr(torch, name)

    def zeros(self, *a, **k):
        k.pop("device", None)
        return torch.zeros(*a, **k)

    def normal(self, mean, std):
        z = torch.randn(std.shape, generator=self.gen)
        self.z_log.append(z.numpy().copy())
        return z * std + mean          # aten normal(Tensor mean, Tensor std): normal_(0,1).mul_(std).add_(mean)


def cpu_training_params(gs):
    """get_training_params (gsmodel.py:95-129) moves everything .to('cuda'); run it with that
    call neutralised"""
    orig = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else orig(self, *a, **k)
    try:
        return gm.get_training_params(gs)
    finally:
        torch.Tensor.to = orig


def n(t):
    return t.detach().numpy().copy()


NAMES = ("pws", "low_shs", "high_shs", "alphas_raw", "scales_raw", "rots_raw")


def snapshot(prefix, params, optimizer, out):
    for g in optimizer.param_groups:
        p = g["params"][0]
        out[prefix + g["name"]] = n(p)
        st = optimizer.state.get(p, None)
        if st is not None and "exp_avg" in st:
            out[prefix + "m_" + g["name"]] = n(st["exp_avg"])
            out[prefix + "v_" + g["name"]] = n(st["exp_avg_sq"])
        assert params[g["name"]] is p


def make_scene(rng, N, sense):
    gs = np.zeros(N, dtype=gio.gsdata_type(12))          # SH degree 1 on disk -> padded to 48
    gs["pw"] = rng.uniform(-2, 2, (N, 3))
    q = rng.normal(size=(N, 4))
    gs["rot"] = q / np.linalg.norm(q, axis=1, keepdims=True)
    base = np.exp(rng.uniform(np.log(0.002 * sense), np.log(0.12 * sense), (N, 1)))
    gs["scale"] = base * rng.uniform(0.6, 1.5, (N, 3))
    gs["alpha"] = 1 / (1 + np.exp(-rng.uniform(-7.5, 4, N)))
    gs["sh"] = rng.normal(size=(N, 12)) * 0.5
    return gs


def make_density(with_state):
    rng = np.random.default_rng(31 if with_state else 32)
    N, sense = 500, 5.0
    gs = make_scene(rng, N, sense)
    params, adam_params = cpu_training_params(gs)
    optimizer = torch.optim.Adam(adam_params, lr=0.0, eps=1e-15)
    out = {"sense_size": np.float64(sense)}
    if with_state:                                       # one real Adam step so that exp_avg / exp_avg_sq exist
        for g in optimizer.param_groups:
            p = g["params"][0]
            p.grad = torch.from_numpy(rng.normal(size=tuple(p.shape)).astype(np.float32) * 1e-3)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
    proxy = TorchProxy()
    gm.torch = proxy
    try:
        model = gm.GSModel(sense, 100)
        # three accumulation steps (update_density_info)
        for it in range(3):
            model.us = torch.zeros(N, 2, requires_grad=True)
            g = rng.normal(size=(N, 2)).astype(np.float32) * 3e-7
            mask = rng.uniform(size=N) < (0.7 if it else 0.5)
            if it == 0:
                mask[:40] = False                        # never seen in any step -> cunt 0 -> 0/0 -> 0
            else:
                mask[:20] = False
            model.us.grad = torch.from_numpy(g)
            model.mask = torch.from_numpy(mask)
            out["acc%d_dloss_dus" % it], out["acc%d_mask" % it] = g, mask
            model.update_density_info()
            out["acc%d_grad_accum" % it] = n(model.grad_accum)
            out["acc%d_cunt" % it] = n(model.cunt)
        snapshot("in_", params, optimizer, out)
        with torch.no_grad():
            model.update_gaussian_density(params, optimizer)
        out["z"] = proxy.z_log[0].reshape(-1, 3)
        snapshot("out_", params, optimizer, out)
        assert model.grad_accum is None and model.cunt is None
        if with_state:
            with torch.no_grad():
                model.reset_alpha(params, optimizer)
            snapshot("reset_", params, optimizer, out)
    finally:
        gm.torch = torch
    Nout = out["out_pws"].shape[0]
    print("density(with_state=%s): N %d -> %d, split %d" % (with_state, N, Nout, out["z"].shape[0]))
    return out


def write_official_ply(f, gs_disk):
    """the layout official 3DGS checkpoints use: x y z nx ny nz f_dc_0..2 f_rest_* opacity
    scale_0..2 rot_0..3, all float32, binary little endian; gs_disk holds the ON-DISK values
    (raw opacity, log scales, channel-major f_rest)"""
    N, rest = gs_disk["f_rest"].shape
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + \
        ["f_rest_%d" % i for i in range(rest)] + ["opacity", "scale_0", "scale_1", "scale_2",
                                                  "rot_0", "rot_1", "rot_2", "rot_3"]
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % N
    hdr += "".join("property float %s\n" % s for s in names) + "end_header\n"
    f.write(hdr.encode("ascii"))
    rows = np.concatenate([gs_disk["xyz"], gs_disk["normals"], gs_disk["f_dc"], gs_disk["f_rest"],
                           gs_disk["opacity"][:, None], gs_disk["scale"], gs_disk["rot"]], axis=1)
    f.write(np.ascontiguousarray(rows, dtype="<f4").tobytes())


def make_io():
    rng = np.random.default_rng(41)
    out = {}
    for tag, N, rest in (("deg3", 257, 45), ("deg1", 64, 9)):    # rest = 0 crashes the reference (:91)
        disk = dict(xyz=rng.uniform(-3, 3, (N, 3)), normals=np.zeros((N, 3)), f_dc=rng.normal(size=(N, 3)),
                    f_rest=rng.normal(size=(N, rest)) * 0.2, opacity=rng.uniform(-6, 6, N),
                    scale=rng.uniform(-6, 0, (N, 3)), rot=rng.normal(size=(N, 4)) * rng.uniform(0.2, 3, (N, 1)))
        buf = io.BytesIO()
        write_official_ply(buf, disk)
        raw = buf.getvalue()
        with tempfile.NamedTemporaryFile(suffix=".ply", delete=False) as f:
            f.write(raw)
        try:
            gs = gio.load_ply(f.name)
        finally:
            os.unlink(f.name)
        out[tag + "_ply"] = np.frombuffer(raw, dtype=np.uint8)
        for k in ("pw", "rot", "scale", "alpha", "sh"):
            out[tag + "_" + k] = np.asarray(gs[k]).copy()
        # recarray -> training tensors (get_training_params) -> recarray (save_training_params)
        params, _ = cpu_training_params(gs)
        for k in NAMES:
            out[tag + "_tp_" + k] = n(params[k])
        with tempfile.TemporaryDirectory() as d:
            fn = os.path.join(d, "x.npy")
            gio.save_training_params(fn, params)
            back = np.load(fn)
        for k in ("pw", "rot", "scale", "alpha", "sh"):
            out[tag + "_back_" + k] = np.asarray(back[k]).copy()
        print("gsio", tag, gs.shape, gs.dtype["sh"].shape, back.dtype["sh"].shape)
    # rotate_gaussian + matrix_to_quaternion (gau_io.py:15-57, 110-126)
    q = rng.normal(size=(50, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    gs = np.zeros(50, dtype=gio.gsdata_type(3))
    gs["pw"], gs["rot"] = rng.uniform(-1, 1, (50, 3)), q
    a = 0.7
    T = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ \
        np.array([[1, 0, 0], [0, np.cos(1.9), -np.sin(1.9)], [0, np.sin(1.9), np.cos(1.9)]])
    out["rotg_T"], out["rotg_in_pw"], out["rotg_in_rot"] = T, gs["pw"].copy(), gs["rot"].copy()
    g2 = gio.rotate_gaussian(T, gs.copy())
    out["rotg_out_pw"], out["rotg_out_rot"] = g2["pw"].copy(), g2["rot"].copy()
    ex = gio.get_example_gs()
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        out["example_" + k] = np.asarray(ex[k]).copy()
    return out


def main():
    d = {}
    for ws in (True, False):
        for k, v in make_density(ws).items():
            d[("s_" if ws else "n_") + k] = v
    np.savez_compressed(os.path.join(HERE, "density.npz"), **d)
    np.savez_compressed(os.path.join(HERE, "gsio.npz"), **make_io())
    for fn in ("density.npz", "gsio.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()

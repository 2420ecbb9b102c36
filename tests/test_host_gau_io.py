"""CPU-only checks of the N3 host logic (easygaussiansplatting_b200/gau_io.py, density.py):
PLY header parsing and the column map (against the fixture the reference's load_ply
produced), the small numpy utilities, save_ply round trip, and that the device paths refuse
to run without CUDA instead of falling back."""
import io
import os

import numpy as np
import pytest
import torch

from easygaussiansplatting_b200 import density, gau_io
from oracle import density_oracle as do

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gsio():
    return dict(np.load(os.path.join(G, "gsio.npz")))


@pytest.mark.parametrize("tag", ["deg3", "deg1"])
def test_ply_header_and_column_map(gsio, tag):
    raw = gsio[tag + "_ply"].tobytes()
    count, names, off = gau_io.read_ply_header(io.BytesIO(raw))
    assert (count, off) == do.parse_ply(raw)[::2] and names == do.parse_ply(raw)[1]
    cmap, sh_dim = gau_io.ply_column_map(names)
    assert sh_dim == gsio[tag + "_sh"].shape[1] and len(cmap) == 11 + sh_dim
    rows = np.frombuffer(raw, "<f4", count * len(names), off).reshape(count, len(names))
    picked = rows[:, cmap]
    # un-activated columns must be exactly what the reference's load_ply returned
    assert np.array_equal(picked[:, :3], gsio[tag + "_pw"])
    assert np.array_equal(picked[:, 11:], gsio[tag + "_sh"])
    # and the activated ones are the reference's up to the activation
    assert np.allclose(np.exp(picked[:, 7:10]), gsio[tag + "_scale"], rtol=1e-6)
    assert np.allclose(1 / (1 + np.exp(-picked[:, 10])), gsio[tag + "_alpha"], rtol=1e-6)
    q = picked[:, 3:7]
    assert np.allclose(q / np.linalg.norm(q, axis=1, keepdims=True), gsio[tag + "_rot"], rtol=1e-5, atol=1e-7)


def test_ply_header_rejects_what_the_kernels_cannot_read():
    def hdr(fmt="binary_little_endian", typ="float", n=17):
        names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1",
                 "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"][:n]
        return ("ply\nformat %s 1.0\ncomment x\nelement vertex 2\n" % fmt +
                "".join("property %s %s\n" % (typ, s) for s in names) + "end_header\n").encode()
    count, names, off = gau_io.read_ply_header(io.BytesIO(hdr()))
    assert count == 2 and len(names) == 17 and gau_io.ply_column_map(names)[1] == 3
    with pytest.raises(ValueError, match="binary_little_endian"):
        gau_io.read_ply_header(io.BytesIO(hdr(fmt="ascii")))
    with pytest.raises(ValueError, match="float32"):
        gau_io.read_ply_header(io.BytesIO(hdr(typ="double")))
    with pytest.raises(ValueError, match="sh_dim"):
        gau_io.ply_column_map(gau_io.read_ply_header(io.BytesIO(hdr(n=16)))[1])
    with pytest.raises(ValueError, match="not a PLY"):
        gau_io.read_ply_header(io.BytesIO(b"plx\n"))
    with pytest.raises(ValueError, match="missing property"):
        gau_io.ply_column_map(["x", "y", "z"] + ["p%d" % i for i in range(14)])


def test_small_numpy_utilities_match_reference(gsio):
    ex = gau_io.get_example_gs()
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        assert np.array_equal(np.asarray(ex[k]), gsio["example_" + k]), k
    gs = np.zeros(50, dtype=gau_io.gsdata_type(3))
    gs["pw"], gs["rot"] = gsio["rotg_in_pw"], gsio["rotg_in_rot"]
    out = gau_io.rotate_gaussian(gsio["rotg_T"], gs)
    assert np.allclose(out["pw"], gsio["rotg_out_pw"], atol=1e-6)
    assert np.allclose(out["rot"], gsio["rotg_out_rot"], atol=1e-6)
    # all four branches of matrix_to_quaternion, against q -> R -> q
    rng = np.random.default_rng(3)
    q = rng.normal(size=(400, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[:100, 0] = rng.uniform(-0.05, 0.05, 100); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]).transpose(2, 0, 1)
    back = gau_io.matrix_to_quaternion(R)
    sign = np.sign((back * q).sum(axis=1, keepdims=True))
    assert np.allclose(back * sign, q, atol=1e-7)


def test_save_ply_is_the_inverse_of_the_reference_loader(gsio, tmp_path):
    gs = np.zeros(len(gsio["deg3_pw"]), dtype=gau_io.gsdata_type(48))
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        gs[k] = gsio["deg3_" + k]
    p = str(tmp_path / "x.ply")
    gau_io.save_ply(p, gs)
    back = do.decode_ply(open(p, "rb").read())
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        assert np.allclose(np.asarray(back[k]), gsio["deg3_" + k], rtol=2e-6, atol=1e-7), k


def test_device_paths_refuse_cpu(gsio, tmp_path):
    p = str(tmp_path / "x.ply")
    open(p, "wb").write(gsio["deg1_ply"].tobytes())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            gau_io.load_ply(p)
    with pytest.raises(ValueError, match="CUDA"):
        gau_io.gs_rows_to_params(torch.zeros(4, 14), 3)
    ctl = density.DensityController(5.0, verbose=False)
    with pytest.raises(ValueError, match="CUDA"):
        density.accumulate(torch.zeros(4, 2), torch.zeros(4, dtype=torch.bool), torch.zeros(4, 1),
                           torch.zeros(4, dtype=torch.int32), True)
    with pytest.raises(RuntimeError, match="update_density_info"):
        ctl.update_gaussian_density({}, None)
    th = density.raw_thresholds(5.0)
    want = do.thresholds(5.0)
    assert all(np.float32(th[k]) == want[k] for k in want)

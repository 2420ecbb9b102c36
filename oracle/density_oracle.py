"""CPU oracle for SURVEY 8f row N3: density control and Gaussian I/O.  TEST INFRASTRUCTURE
ONLY -- imported by tests/ (and nothing in the product); a numpy float32 restatement of the
reference's Python, pinned to the reference itself by tests/golden/density.npz and gsio.npz
(tests/test_oracle_golden.py::test_density_* / test_gsio_*).

Each function names the reference lines it follows.  Index work (classes, slots, counts,
row order) is exact; float work is float32 like the reference's torch / numpy calls and agrees
with them to 1 ulp-level differences of exp/log (tolerance 2e-6 relative in the tests).
"""
import numpy as np

F = np.float32
NAMES = ("pws", "low_shs", "high_shs", "alphas_raw", "scales_raw", "rots_raw")
WIDTH = dict(pws=3, low_shs=3, high_shs=45, alphas_raw=1, scales_raw=3, rots_raw=4)
KEEP, CLONE, SPLIT, PRUNE = 0, 1, 2, 3


def thresholds(sense_size):
    """GSModel.__init__ (gsmodel.py:170-181) + the raw-space constants update_gaussian_density
    compares against (:238-239): float32 values of logit(0.005), log(0.1 sense), 4e-7,
    0.01 sense"""
    return dict(alpha_raw_min=F(np.log(0.005 / (1 - 0.005))), scale_raw_max=F(np.log(0.1 * sense_size)),
                grad_min=F(4e-7), scale_clone_max=F(0.01 * sense_size))


def accumulate(dloss_dus, mask, grad_accum, cunt):
    """GSModel.update_density_info (gsmodel.py:219-234).  grad_accum/cunt None = first call:
    the norm of EVERY Gaussian is stored, masked or not (:228-229)"""
    g = np.sqrt((dloss_dus.astype(F) ** 2).sum(axis=1, dtype=F)).astype(F)[:, None]
    if grad_accum is None:
        return g.copy(), mask.astype(np.int32)
    grad_accum = grad_accum.copy()
    grad_accum[mask] += g[mask]
    return grad_accum, cunt + mask.astype(np.int32)


def classify(alphas_raw, scales_raw, grad_accum, cunt, th):
    """update_gaussian_density (gsmodel.py:238-262): class per Gaussian.  The prune test is on
    the raw values (:238-239); the clone/split test on exp(scales_raw) (:257) and on
    grad_accum / cunt with 0/0 -> 0 (:244-245)"""
    a = alphas_raw.reshape(-1).astype(F)
    prune = (a < th["alpha_raw_min"]) | (scales_raw.max(axis=1) > th["scale_raw_max"])
    with np.errstate(divide="ignore", invalid="ignore"):
        g = grad_accum.reshape(-1).astype(F) / cunt.astype(F)
    g[np.isnan(g)] = 0
    by_grad = g >= th["grad_min"]
    small = np.exp(scales_raw.astype(F)).max(axis=1) <= th["scale_clone_max"]
    cls = np.full(a.shape[0], KEEP, np.uint8)
    cls[by_grad & small] = CLONE
    cls[by_grad & ~small] = SPLIT
    cls[prune] = PRUNE
    return cls


def _normalize(q):
    """torch.nn.functional.normalize (utils.py:146-147): q / max(|q|, 1e-12)"""
    nrm = np.sqrt((q * q).sum(axis=1, dtype=F)).astype(F)
    return (q / np.maximum(nrm, F(1e-12))[:, None]).astype(F)


def _rotate(q, v):
    """rotate_vector_by_quaternion (utils.py:46-54), q = (w, x, y, z)"""
    q = _normalize(q)
    u, s = q[:, 1:], q[:, :1]
    uv = (u * v).sum(axis=1, keepdims=True, dtype=F)
    uu = (u * u).sum(axis=1, keepdims=True, dtype=F)
    return (F(2) * u * uv + v * (s * s - uu) + F(2) * np.cross(u, v).astype(F) * s).astype(F)


def densify(params, m, v, cls, z):
    """update_gaussian_density (gsmodel.py:236-318) after classification.  params / m / v:
    dicts over NAMES (m, v may be None = optimizer has no state yet, update_params' else
    branch :150-153).  z: unit normals [n_split, 3]; the reference draws
    torch.normal(0, std=scales[split]) = z * std (:276-277).
    Output rows: survivors in order, then clones, then splits (prune_params :156-166,
    update_params :132-153, torch.cat order :287-292); new rows get zero Adam moments."""
    keep, cl, sp = cls != PRUNE, cls == CLONE, cls == SPLIT
    alphas = (F(1) / (F(1) + np.exp(-params["alphas_raw"].astype(F)))).astype(F)      # get_alphas
    scales = np.exp(params["scales_raw"].astype(F)).astype(F)                        # get_scales
    rots = _normalize(params["rots_raw"].astype(F))                                  # get_rots
    samples = (z.astype(F) * scales[sp]).astype(F)
    new = {
        "pws": np.concatenate([params["pws"][cl], params["pws"][sp] + _rotate(rots[sp], samples)]),
        "low_shs": np.concatenate([params["low_shs"][cl], params["low_shs"][sp]]),
        "high_shs": np.concatenate([params["high_shs"][cl], params["high_shs"][sp]]),
        # the original of a split keeps its size: `scales` is a temporary (:281-282)
        "scales_raw": np.log(np.concatenate([scales[cl], scales[sp] * F(0.6)])),
        "rots_raw": np.concatenate([rots[cl], rots[sp]]),
    }
    a_new = np.concatenate([alphas[cl], alphas[sp]])
    with np.errstate(divide="ignore"):
        new["alphas_raw"] = np.log(a_new / (F(1) - a_new))                            # get_alphas_raw
    out_p, out_m, out_v = {}, {}, {}
    for k in NAMES:
        out_p[k] = np.concatenate([params[k][keep], new[k].astype(F)]).astype(F)
        if m is not None:
            zeros = np.zeros_like(new[k], dtype=F)
            out_m[k] = np.concatenate([m[k][keep], zeros])
            out_v[k] = np.concatenate([v[k][keep], zeros])
    counts = (int(keep.sum()), int(cl.sum()), int(sp.sum()))
    return out_p, (out_m if m is not None else None), (out_v if m is not None else None), counts


def reset_alpha(alphas_raw, reset_alpha_val=0.01):
    """GSModel.reset_alpha (gsmodel.py:320-331): clamp from above at logit(0.01); the alpha
    group's Adam moments become zero"""
    val = F(np.log(reset_alpha_val / (1 - reset_alpha_val)))
    out = alphas_raw.copy()
    out[out > val] = val
    return out, np.zeros_like(out), np.zeros_like(out)


# ----------------------------------------------------------------------------- Gaussian I/O

def gsdata_type(sh_dim):
    """gau_io.py:7-12"""
    return [("pw", "<f4", (3,)), ("rot", "<f4", (4,)), ("scale", "<f4", (3,)), ("alpha", "<f4"),
            ("sh", "<f4", (sh_dim,))]


def parse_ply(raw):
    """header of a binary_little_endian PLY -> (vertex count, property names, numpy dtype,
    data offset); what plyfile.PlyData.read does for load_ply (gau_io.py:61)"""
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    names, count = [], 0
    for line in raw[:end].decode("ascii").splitlines():
        tok = line.split()
        if tok[:2] == ["element", "vertex"]:
            count = int(tok[2])
        elif tok[:1] == ["property"]:
            assert tok[1] in ("float", "float32")
            names.append(tok[2])
    return count, names, end


def decode_ply(raw):
    """load_ply (gau_io.py:60-107): sigmoid(opacity), exp(scale), normalised rot, f_rest
    re-ordered from channel-major [3, k] to coefficient-major [k, 3] (:91); float32 math
    because the columns are float32 arrays"""
    count, names, off = parse_ply(raw)
    rows = np.frombuffer(raw, dtype="<f4", count=count * len(names), offset=off).reshape(count, len(names))
    col = {nm: rows[:, i] for i, nm in enumerate(names)}
    pws = np.stack([col["x"], col["y"], col["z"]], axis=1)
    alphas = (F(1) / (F(1) + np.exp(-col["opacity"]))).astype(F)
    scales = np.exp(np.stack([col["scale_%d" % i] for i in range(3)], axis=1)).astype(F)
    rots = np.stack([col["rot_%d" % i] for i in range(4)], axis=1)
    rots = (rots / np.sqrt((rots * rots).sum(axis=1, dtype=F))[:, None]).astype(F)
    sh_dim = len(names) - 14
    rest = sh_dim - 3
    shs = np.zeros((count, sh_dim), F)
    for i in range(3):
        shs[:, i] = col["f_dc_%d" % i]
    if rest:
        r = np.stack([col["f_rest_%d" % i] for i in range(rest)], axis=1)
        shs[:, 3:] = r.reshape(-1, 3, rest // 3).transpose(0, 2, 1).reshape(-1, rest)
    return np.rec.fromarrays([pws, rots, scales, alphas, shs], dtype=gsdata_type(sh_dim))


def training_params(gs):
    """get_training_params (gsmodel.py:95-113): raw (unactivated) tensors; SH padded to 48
    with 0.001"""
    sh = gs["sh"].astype(F)
    high = np.full((len(gs), 45), F(0.001), F)
    high[:, : sh.shape[1] - 3] = sh[:, 3:]
    a = gs["alpha"].astype(F)[:, None]
    with np.errstate(divide="ignore"):
        return dict(pws=gs["pw"].astype(F), low_shs=sh[:, :3].copy(), high_shs=high,
                    alphas_raw=np.log(a / (F(1) - a)), scales_raw=np.log(gs["scale"].astype(F)),
                    rots_raw=gs["rot"].astype(F))


def params_to_gs(params):
    """save_training_params (gau_io.py:138-153): activated values in the .npy record layout"""
    shs = np.concatenate([params["low_shs"], params["high_shs"]], axis=1)
    alphas = (F(1) / (F(1) + np.exp(-params["alphas_raw"].astype(F)))).reshape(-1)
    return np.rec.fromarrays([params["pws"], _normalize(params["rots_raw"]),
                              np.exp(params["scales_raw"]).astype(F), alphas.astype(F), shs],
                             dtype=gsdata_type(shs.shape[1]))

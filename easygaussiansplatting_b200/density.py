"""Density control on the C ABI: the mirror of the density half of
gsplat.gsmodel.GSModel (reference gsplat/gsmodel.py:168-331) and of prune_params /
update_params (:132-166).

    ctl = DensityController(sense_size)
    ctl.update_density_info(us.grad, mask)            # every iteration, after backward
    ctl.update_gaussian_density(params, optimizer)    # every few epochs
    ctl.reset_alpha(params, optimizer)

`params` is the reference's dict of six leaf tensors, `optimizer` a torch.optim.Adam whose
param groups carry the reference's "name" keys (get_training_params, gsmodel.py:114-127).
The parameters and both Adam moments are rebuilt by ONE kernel pass (gsb_density_apply)
instead of 18 boolean-index + torch.cat round trips.  Split offsets use unit normals drawn
with the SAME torch generator call sequence as the reference (torch.normal(mean, std) is
normal_(0, 1) * std + mean), so a run seeded like the reference splits identically.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GAUSSIAN_TENSORS, GAUSSIAN_WIDTHS, Gaussians


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check(t, name, shape=None, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU fallback)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t.contiguous()


def gaussians_struct(tensors):
    """dict of the six tensors -> (gsb_gaussians, keep-alive list)"""
    g, keep = Gaussians(), []
    for k in GAUSSIAN_TENSORS:
        t = tensors[k]
        keep.append(t)
        setattr(g, k, t.data_ptr())
    return g, keep


def check_params(params, what="params"):
    N = params["pws"].shape[0]
    out = {}
    for k, w in zip(GAUSSIAN_TENSORS, GAUSSIAN_WIDTHS):
        out[k] = _check(params[k].detach(), "%s[%r]" % (what, k), (N, w))
    return N, out


def raw_thresholds(sense_size, alpha_threshold=0.005, grad_threshold=4e-7):
    """GSModel.__init__ (gsmodel.py:170-181) in the units the kernels compare in"""
    f = np.float32
    return dict(alpha_raw_min=float(f(np.log(alpha_threshold / (1 - alpha_threshold)))),
                scale_raw_max=float(f(np.log(0.1 * sense_size))), grad_min=float(f(grad_threshold)),
                scale_clone_max=float(f(0.01 * sense_size)))


def accumulate(dloss_dus, mask, grad_accum, cunt, first):
    """gsb_density_accumulate on caller-owned accumulators"""
    lib = _lib.load()
    N = dloss_dus.shape[0]
    dus = _check(dloss_dus.detach().reshape(N, 2), "dloss_dus", (N, 2))
    m = _check(mask, "mask", (N,), torch.bool)
    _lib.check(lib.gsb_density_accumulate(N, _ptr(dus), _ptr(m), _ptr(grad_accum), _ptr(cunt), int(first), _stream()), lib)


def plan(alphas_raw, scales_raw, grad_accum, cunt, th):
    """-> (cls uint8 [N], slots int32 [N,3], (K, C, S)).  Waits for the counts (one D2H)."""
    lib = _lib.load()
    N = alphas_raw.shape[0]
    dev = alphas_raw.device
    cls = torch.empty(N, dtype=torch.uint8, device=dev)
    slots = torch.empty((N, 3), dtype=torch.int32, device=dev)
    nbytes = lib.gsb_density_workspace_bytes(N)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    counts = (C.c_int64 * 3)()
    _lib.check(lib.gsb_density_plan(N, _ptr(alphas_raw), _ptr(scales_raw), _ptr(grad_accum), _ptr(cunt),
                                    th["alpha_raw_min"], th["scale_raw_max"], th["grad_min"], th["scale_clone_max"],
                                    _ptr(ws), nbytes, _ptr(cls), _ptr(slots), counts, _stream()), lib)
    return cls, slots, (int(counts[0]), int(counts[1]), int(counts[2]))


def apply(cls, slots, counts, src, src_m, src_v, z):
    """gsb_density_apply -> (dst, dst_m, dst_v) dicts of fresh tensors with K + C + S rows"""
    lib = _lib.load()
    K, Cn, S = counts
    N = cls.shape[0]
    dev = cls.device
    rows = K + Cn + S

    def fresh():
        return {k: torch.empty((rows, w), dtype=torch.float32, device=dev)
                for k, w in zip(GAUSSIAN_TENSORS, GAUSSIAN_WIDTHS)}

    dst = fresh()
    state = src_m is not None
    dst_m, dst_v = (fresh(), fresh()) if state else (None, None)
    keep = []
    structs = []
    for d in (src, src_m, src_v, dst, dst_m, dst_v):
        if d is None:
            structs.append(None)
        else:
            g, ka = gaussians_struct(d)
            keep.append(ka)
            structs.append(C.byref(g))
            keep.append(g)
    _lib.check(lib.gsb_density_apply(N, _ptr(cls), _ptr(slots), K, Cn, S, structs[0], structs[1], structs[2], _ptr(z),
                                     structs[3], structs[4], structs[5], _stream()), lib)
    return dst, dst_m, dst_v


class DensityController:
    """the density-control half of gsplat.gsmodel.GSModel (gsmodel.py:168-331), same attribute
    names (`grad_accum`, `cunt`) and the same printed report"""

    def __init__(self, sense_size, verbose=True):
        self.grad_threshold = 4e-7
        self.scale_threshold = 0.01 * sense_size
        self.alpha_threshold = 0.005
        self.big_threshold = 0.1 * sense_size
        self.reset_alpha_val = 0.01
        self.th = raw_thresholds(sense_size, self.alpha_threshold, self.grad_threshold)
        self.grad_accum = None
        self.cunt = None
        self.verbose = verbose

    def update_density_info(self, dloss_dus, mask):
        """gsmodel.py:219-234 (the caller passes us.grad and the GSFunction mask)"""
        N = dloss_dus.shape[0]
        first = self.cunt is None
        if first:
            self.grad_accum = torch.empty((N, 1), dtype=torch.float32, device=dloss_dus.device)
            self.cunt = torch.empty(N, dtype=torch.int32, device=dloss_dus.device)
        accumulate(dloss_dus, mask, self.grad_accum, self.cunt, first)

    @torch.no_grad()
    def update_gaussian_density(self, params, optimizer, generator=None):
        """gsmodel.py:236-318 with prune_params / update_params (:132-166) fused in.  Mutates
        `params` and the optimizer's groups / state like the reference; returns the report."""
        if self.cunt is None:
            raise RuntimeError("update_gaussian_density needs at least one update_density_info call")
        N, src = check_params(params)
        groups = {g["name"]: g for g in optimizer.param_groups}
        states = {k: optimizer.state.get(groups[k]["params"][0], None) for k in GAUSSIAN_TENSORS}
        have = [s is not None and "exp_avg" in s for s in states.values()]
        if any(have) and not all(have):
            raise ValueError("optimizer state must exist for all six groups or for none")
        src_m = src_v = None
        if all(have):
            src_m = {k: _check(states[k]["exp_avg"], "exp_avg[%r]" % k, src[k].shape) for k in GAUSSIAN_TENSORS}
            src_v = {k: _check(states[k]["exp_avg_sq"], "exp_avg_sq[%r]" % k, src[k].shape) for k in GAUSSIAN_TENSORS}
        cls, slots, (K, Cn, S) = plan(src["alphas_raw"], src["scales_raw"], self.grad_accum, self.cunt, self.th)
        # the reference's torch.normal(mean=zeros(S,3), std=scales[split]) (:276-277)
        z = torch.empty((S, 3), dtype=torch.float32, device=src["pws"].device).normal_(0, 1, generator=generator)
        dst, dst_m, dst_v = apply(cls, slots, (K, Cn, S), src, src_m, src_v, z)
        for k in GAUSSIAN_TENSORS:
            group = groups[k]
            old = group["params"][0]
            new = torch.nn.Parameter(dst[k].requires_grad_(True))
            st = optimizer.state.pop(old, None)
            if st is not None and dst_m is not None:
                st["exp_avg"], st["exp_avg_sq"] = dst_m[k], dst_v[k]
                optimizer.state[new] = st
            group["params"][0] = new
            params[k] = new
        report = dict(pruned=N - K, cloned=Cn, splited=S, total=K + Cn + S)
        if self.verbose:
            print("---------------------")
            print("gaussian density update report")
            print("pruned num: ", report["pruned"])
            print("cloned num: ", report["cloned"])
            print("splited num: ", report["splited"])
            print("total gaussian number: ", report["total"])
            print("---------------------")
        self.grad_accum = None
        self.cunt = None
        return report

    @torch.no_grad()
    def reset_alpha(self, params, optimizer):
        """gsmodel.py:320-331"""
        lib = _lib.load()
        val = float(np.float32(np.log(self.reset_alpha_val / (1 - self.reset_alpha_val))))
        p = params["alphas_raw"]
        a = _check(p.detach(), "params['alphas_raw']")
        if a.data_ptr() != p.data_ptr():
            raise ValueError("params['alphas_raw'] must be contiguous")
        group = [g for g in optimizer.param_groups if g["name"] == "alphas_raw"][0]
        st = optimizer.state.get(group["params"][0], None)
        m = v = None
        if st is not None and "exp_avg" in st:
            m, v = _check(st["exp_avg"], "exp_avg"), _check(st["exp_avg_sq"], "exp_avg_sq")
        _lib.check(lib.gsb_reset_alpha(a.shape[0], _ptr(a), _ptr(m), _ptr(v), val, _stream()), lib)

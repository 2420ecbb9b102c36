"""The reference's `gsplatcu` operator surface on top of the C ABI (include/gsplat_b200.h).

Seven functions with the reference's positional signatures, return arity, shapes, dtypes
and in-place side effects (gsplatcu/ext.cpp:10-76, gsplatcu/gausplat.cu): `project`,
`computeCov3D`, `computeCov2D`, `sh2Color`, `inverseCov2D`, `splat`, `splatB`.
torch is used for device memory and the current stream only; all compute is in
libgsplat_b200.so.  Differences from the reference, all deliberate:
  * dtype / device / shape are checked and violations raise (the reference checks nothing);
  * work is enqueued on torch's current stream and nothing device-synchronises except the
    one read of the patch count inside `splat` (the reference syncs the device after every
    kernel, common.cuh:17-25);
  * N == 0 and P <= 1 are handled (SURVEY 8b "degenerate inputs").
"""
import ctypes as C

import torch

from . import _lib
from .jacobian_tensor import wrap as _jac

_lib_handle = None


def _L():
    global _lib_handle
    if _lib_handle is None:
        _lib_handle = _lib.load()
    return _lib_handle


def _chk(t, name, dtype=torch.float32, last=None, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    if last is not None and (t.dim() == 0 or t.shape[-1] != last):
        raise ValueError("%s must have last dim %d, got shape %s" % (name, last, tuple(t.shape)))
    return t.contiguous()


def _same_device(ref, *ts):
    for t in ts:
        if t.device != ref.device:
            raise ValueError("all tensors must be on the same device")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def project(pws, Rcw, tcw, focal_x, focal_y, center_x, center_y, calc_J):
    """ext.cpp:54-61.  -> [us[N,2], pcs[N,3], depths[N]] (+ du_dpcs[N,2,3] if calc_J)"""
    pws = _chk(pws, "pws", last=3, ndim=2)
    Rcw = _chk(Rcw, "Rcw"); tcw = _chk(tcw, "tcw")
    if Rcw.numel() != 9 or tcw.numel() != 3:
        raise ValueError("Rcw must have 9 and tcw 3 elements")
    _same_device(pws, Rcw, tcw)
    N = pws.shape[0]
    o = dict(dtype=torch.float32, device=pws.device)
    us = torch.empty((N, 2), **o); pcs = torch.empty((N, 3), **o); depths = torch.empty((N,), **o)
    J = torch.empty((N, 2, 3), **o) if calc_J else None
    lib = _L()
    with torch.cuda.device(pws.device):
        _lib.check(lib.gsb_project(N, _ptr(pws), _ptr(Rcw), _ptr(tcw), float(focal_x), float(focal_y),
                                   float(center_x), float(center_y), _ptr(us), _ptr(pcs), _ptr(depths),
                                   _ptr(J), _stream()), lib)
    return [us, pcs, depths, _jac(J)] if calc_J else [us, pcs, depths]


def computeCov3D(rots, scales, depths, calc_J):
    """ext.cpp:39-42.  -> [cov3ds[N,6]] (+ dcov3d_drots[N,6,4], dcov3d_dscales[N,6,3])"""
    rots = _chk(rots, "rots", last=4, ndim=2); scales = _chk(scales, "scales", last=3, ndim=2)
    depths = _chk(depths, "depths")
    _same_device(rots, scales, depths)
    N = rots.shape[0]
    if scales.shape[0] != N or depths.numel() != N:
        raise ValueError("rots, scales, depths disagree on N")
    o = dict(dtype=torch.float32, device=rots.device)
    cov = torch.empty((N, 6), **o)
    Jr = torch.empty((N, 6, 4), **o) if calc_J else None
    Js = torch.empty((N, 6, 3), **o) if calc_J else None
    lib = _L()
    with torch.cuda.device(rots.device):
        _lib.check(lib.gsb_compute_cov3d(N, _ptr(rots), _ptr(scales), _ptr(depths), _ptr(cov), _ptr(Jr),
                                         _ptr(Js), _stream()), lib)
    return [cov, _jac(Jr), _jac(Js)] if calc_J else [cov]


def computeCov2D(cov3ds, pcs, Rcw, depths, focal_x, focal_y, width, height, calc_J):
    """ext.cpp:44-52.  -> [cov2ds[N,3]] (+ dcov2d_dcov3ds[N,3,6], dcov2d_dpcs[N,3,3])"""
    cov3ds = _chk(cov3ds, "cov3ds", last=6, ndim=2); pcs = _chk(pcs, "pcs", last=3, ndim=2)
    Rcw = _chk(Rcw, "Rcw"); depths = _chk(depths, "depths")
    _same_device(pcs, cov3ds, Rcw, depths)
    N = pcs.shape[0]
    if cov3ds.shape[0] != N or depths.numel() != N or Rcw.numel() != 9:
        raise ValueError("cov3ds, pcs, depths disagree on N (or Rcw is not 3x3)")
    o = dict(dtype=torch.float32, device=pcs.device)
    cov = torch.empty((N, 3), **o)
    Jc = torch.empty((N, 3, 6), **o) if calc_J else None
    Jp = torch.empty((N, 3, 3), **o) if calc_J else None
    lib = _L()
    with torch.cuda.device(pcs.device):
        _lib.check(lib.gsb_compute_cov2d(N, _ptr(cov3ds), _ptr(pcs), _ptr(Rcw), _ptr(depths),
                                         float(focal_x), float(focal_y), float(width), float(height),
                                         _ptr(cov), _ptr(Jc), _ptr(Jp), _stream()), lib)
    return [cov, _jac(Jc), _jac(Jp)] if calc_J else [cov]


def sh2Color(shs, pws, twc, calc_J):
    """ext.cpp:63-66.  shs[N,3k], k in {1,4,9,16}.
    -> [colors[N,3]] (+ dcolor_dshs[N,1,k], dcolor_dpws[N,3,3])"""
    shs = _chk(shs, "shs", ndim=2); pws = _chk(pws, "pws", last=3, ndim=2); twc = _chk(twc, "twc")
    _same_device(pws, shs, twc)
    N = pws.shape[0]
    if shs.shape[0] != N or shs.shape[1] % 3 != 0 or shs.shape[1] // 3 not in (1, 4, 9, 16):
        raise ValueError("shs must be [N, 3k] with k in {1,4,9,16}, got %s" % (tuple(shs.shape),))
    if twc.numel() != 3:
        raise ValueError("twc must have 3 elements")
    k = shs.shape[1] // 3
    o = dict(dtype=torch.float32, device=pws.device)
    col = torch.empty((N, 3), **o)
    Js = torch.empty((N, 1, k), **o) if calc_J else None
    Jp = torch.empty((N, 3, 3), **o) if calc_J else None
    lib = _L()
    with torch.cuda.device(pws.device):
        _lib.check(lib.gsb_sh2color(N, k, _ptr(shs), _ptr(pws), _ptr(twc), _ptr(col), _ptr(Js), _ptr(Jp),
                                    _stream()), lib)
    return [col, _jac(Js), _jac(Jp)] if calc_J else [col]


def inverseCov2D(cov2ds, depths, calc_J):
    """ext.cpp:34-36.  MUTATES depths (NaN determinant -> -1, kernel.cu:301-305).
    -> [cinv2ds[N,3], areas[N,2] int32] (+ dcinv2d_dcov2ds[N,3,3])"""
    cov2ds = _chk(cov2ds, "cov2ds", last=3, ndim=2); depths = _chk(depths, "depths")
    _same_device(cov2ds, depths)
    N = cov2ds.shape[0]
    if depths.numel() != N:
        raise ValueError("cov2ds and depths disagree on N")
    o = dict(dtype=torch.float32, device=cov2ds.device)
    cinv = torch.empty((N, 3), **o)
    areas = torch.empty((N, 2), dtype=torch.int32, device=cov2ds.device)
    J = torch.empty((N, 3, 3), **o) if calc_J else None
    lib = _L()
    with torch.cuda.device(cov2ds.device):
        _lib.check(lib.gsb_inverse_cov2d(N, _ptr(cov2ds), _ptr(depths), _ptr(cinv), _ptr(areas), _ptr(J),
                                         _stream()), lib)
    return [cinv, areas, _jac(J)] if calc_J else [cinv, areas]


GSB_CAPACITY_EXCEEDED = 2
_CAPACITY = {}  # (device index, H, W, N) -> (P_cap, depth_key_cap) learnt from the previous frame
_STATUS = {}    # device index -> pinned int32[4] for gsb_splat_forward's status read
CAPACITY_STATS = {"capacity_frames": 0, "exact_frames": 0, "outgrown": 0}  # how the frames of `capacity=True` went


def _next_capacity(P, dkmax):
    """Bounds for the next frame of the same shape: 25 % head-room on the patch count, and the key
    width the exact path would pick for this frame's largest depth key."""
    bits = max(int(dkmax), 1).bit_length()
    return int(P * 1.25) + 4096, (1 << bits) - 1


def splat(height, width, us, cinv2ds, alphas, depths, colors, areas, records=None, capacity=False):
    """ext.cpp:10-18.  MUTATES depths / areas for Gaussians that touch no tile
    (kernel.cu:114-119).  -> [image[3,H,W], contrib[H,W] i32, final_tau[H,W],
    patch_range_per_tile[T,2] i32, gsid_per_patch[P] i32]
    records (extension): the packed per-Gaussian records `preprocess(..., alphas=...)` wrote;
    the pack pass is then skipped.
    capacity (extension): size the sort from the previous frame of the same shape
    (gsb_splat_forward) instead of reading the patch count back in the middle of the frame; a
    frame that outgrows the bound is transparently redone the exact way."""
    H, W = int(height), int(width)
    if H <= 0 or W <= 0:
        raise ValueError("height and width must be positive")
    us = _chk(us, "us", last=2, ndim=2); cinv2ds = _chk(cinv2ds, "cinv2ds", last=3, ndim=2)
    alphas = _chk(alphas, "alphas"); depths = _chk(depths, "depths")
    colors = _chk(colors, "colors", last=3, ndim=2)
    areas = _chk(areas, "areas", dtype=torch.int32, last=2, ndim=2)
    _same_device(us, cinv2ds, alphas, depths, colors, areas)
    N = us.shape[0]
    if not (cinv2ds.shape[0] == N and alphas.numel() == N and depths.numel() == N
            and colors.shape[0] == N and areas.shape[0] == N):
        raise ValueError("splat inputs disagree on N")
    dev = us.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    lib = _L()
    if records is not None:
        records = _chk(records, "records", last=12, ndim=2)
        if records.shape[0] != N:
            raise ValueError("records must be [N, 12]")
    cap_key = (dev.index, H, W, N)
    if capacity and cap_key in _CAPACITY and N > 0:
        P_cap, dk_cap = _CAPACITY[cap_key]
        with torch.cuda.device(dev):
            st = _stream()
            status = _STATUS.get(dev.index)
            if status is None:
                status = _STATUS[dev.index] = torch.zeros(4, dtype=torch.int32).pin_memory()
            bin_bytes = lib.gsb_splat_bin_workspace_bytes(N)
            bin_ws = torch.empty((bin_bytes,), dtype=torch.uint8, device=dev)
            ws_bytes = lib.gsb_splat_workspace_bytes(N, H, W, P_cap)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            image = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            contrib = torch.empty((H, W), dtype=torch.int32, device=dev)
            final_tau = torch.empty((H, W), dtype=torch.float32, device=dev)
            ranges = torch.empty((T, 2), dtype=torch.int32, device=dev)
            gsid = torch.empty((P_cap,), dtype=torch.int32, device=dev)
            rc = lib.gsb_splat_forward(H, W, N, _ptr(us), _ptr(cinv2ds), _ptr(alphas), _ptr(depths), _ptr(colors),
                                       _ptr(areas), _ptr(records), P_cap, dk_cap, _ptr(bin_ws), bin_bytes, _ptr(ws),
                                       ws_bytes, _ptr(image), _ptr(contrib), _ptr(final_tau), _ptr(ranges),
                                       _ptr(gsid), status.data_ptr(), st)
            if rc == 0:
                CAPACITY_STATS["capacity_frames"] += 1
                P, dkmax = int(status[0]) & 0xffffffff, int(status[1]) & 0xffffffff
                _CAPACITY[cap_key] = _next_capacity(P, dkmax)
                gsid = gsid[:P]
                if P > 0:
                    if records is not None:
                        _remember_records(gsid, records, 0, (us, cinv2ds, alphas, colors))
                    else:
                        _remember_records(gsid, ws, lib.gsb_splat_records_offset(N, H, W, P_cap),
                                          (us, cinv2ds, alphas, colors))
                return [image, contrib, final_tau, ranges, gsid]
            if rc != GSB_CAPACITY_EXCEEDED:
                _lib.check(rc, lib)
            CAPACITY_STATS["outgrown"] += 1
            del _CAPACITY[cap_key]  # outgrown: this frame the exact way (the in-place culls are idempotent)
    with torch.cuda.device(dev):
        st = _stream()
        bin_bytes = lib.gsb_splat_bin_workspace_bytes(N)
        bin_ws = torch.empty((bin_bytes,), dtype=torch.uint8, device=dev)
        P, dkmax = C.c_int64(0), C.c_uint32(0)
        _lib.check(lib.gsb_splat_bin(H, W, N, _ptr(us), _ptr(depths), _ptr(areas), _ptr(bin_ws), bin_bytes,
                                     C.byref(P), C.byref(dkmax), st), lib)
        P = int(P.value)
        ws_bytes = lib.gsb_splat_workspace_bytes(N, H, W, P)
        ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
        image = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        contrib = torch.empty((H, W), dtype=torch.int32, device=dev)
        final_tau = torch.empty((H, W), dtype=torch.float32, device=dev)
        ranges = torch.empty((T, 2), dtype=torch.int32, device=dev)
        gsid = torch.empty((P,), dtype=torch.int32, device=dev)
        if capacity:
            CAPACITY_STATS["exact_frames"] += 1
            _CAPACITY[cap_key] = _next_capacity(P, dkmax.value)
        _lib.check(lib.gsb_splat_render(H, W, N, P, dkmax.value, _ptr(us), _ptr(cinv2ds), _ptr(alphas),
                                        _ptr(depths), _ptr(colors), _ptr(records), _ptr(bin_ws), _ptr(ws), ws_bytes,
                                        _ptr(image), _ptr(contrib), _ptr(final_tau), _ptr(ranges), _ptr(gsid),
                                        st), lib)
        # the workspaces are consumed by kernels already enqueued on `st`; the caching
        # allocator only reuses them for later work on the same stream
        if P > 0:
            if records is not None:
                _remember_records(gsid, records, 0, (us, cinv2ds, alphas, colors))
            else:
                _remember_records(gsid, ws, lib.gsb_splat_records_offset(N, H, W, P), (us, cinv2ds, alphas, colors))
    return [image, contrib, final_tau, ranges, gsid]


# The forward leaves the packed per-Gaussian records in its workspace (or got them from
# `preprocess`); splatB can reuse them instead of packing the same records again.  The handle
# travels ON the gsid_per_patch tensor `splat` returns (an attribute of that tensor object; it
# survives autograd's save_for_backward), so identity is never inferred from addresses the
# caching allocator may hand out again, and the workspace lives exactly as long as that tensor.
# The inputs must still be the very tensors the forward saw, unmodified (data_ptr + torch's
# in-place version counter); writes that bypass the counter (`.data`, foreign kernels) are the
# caller's to announce with `forget_records(gsid)`.
_RECORDS_ATTR = "_gsb_records"


def _input_key(inputs):
    try:
        return tuple((t.data_ptr(), t._version) for t in inputs)
    except RuntimeError:  # inference tensors have no version counter: never cache
        return None


def _remember_records(gsid, ws, offset, inputs):
    key = _input_key(inputs)
    if key is not None:
        setattr(gsid, _RECORDS_ATTR, (ws, offset, key, gsid.data_ptr(), gsid.numel()))


def _cached_records(gsid, inputs):
    h = getattr(gsid, _RECORDS_ATTR, None)
    if h is None:
        return None
    ws, offset, key, gptr, P = h
    if gptr != gsid.data_ptr() or P != gsid.numel() or key != _input_key(inputs):
        return None
    return ws.data_ptr() + offset


def forget_records(gsid):
    """Drop the packed records `splat` attached to its gsid_per_patch output (frees the forward
    workspace early; splatB then packs the records again)."""
    if hasattr(gsid, _RECORDS_ATTR):
        delattr(gsid, _RECORDS_ATTR)


def clear_record_cache():
    """Kept for callers of the round-1 API: there is no global cache any more (see above)."""


def splatB(height, width, us, cinv2ds, alphas, depths, colors, contrib, final_tau,
           patch_range_per_tile, gsid_per_patch, dloss_dgammas, moments_only=False):
    """ext.cpp:20-32.  -> [dloss_dus[N,1,2], dloss_dcinv2ds[N,1,3], dloss_dalphas[N,1,1],
    dloss_dcolors[N,1,3]]   (`depths` is accepted and unused, as in the reference)
    moments_only (extension): -> the raw moment rows [N,9] for `preprocessB(..., moments=...)`;
    the conversion pass to the four tensors is skipped."""
    H, W = int(height), int(width)
    us = _chk(us, "us", last=2, ndim=2); cinv2ds = _chk(cinv2ds, "cinv2ds", last=3, ndim=2)
    alphas = _chk(alphas, "alphas"); colors = _chk(colors, "colors", last=3, ndim=2)
    contrib = _chk(contrib, "contrib", dtype=torch.int32); final_tau = _chk(final_tau, "final_tau")
    ranges = _chk(patch_range_per_tile, "patch_range_per_tile", dtype=torch.int32, last=2)
    gsid = _chk(gsid_per_patch, "gsid_per_patch", dtype=torch.int32)
    dl = _chk(dloss_dgammas, "dloss_dgammas")
    _same_device(us, cinv2ds, alphas, colors, contrib, final_tau, ranges, gsid, dl)
    N = us.shape[0]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    if contrib.numel() != H * W or final_tau.numel() != H * W or dl.numel() != 3 * H * W:
        raise ValueError("contrib / final_tau / dloss_dgammas do not match height x width")
    if ranges.shape[0] != T:
        raise ValueError("patch_range_per_tile does not match the tile grid")
    if not (cinv2ds.shape[0] == N and alphas.numel() == N and colors.shape[0] == N):
        raise ValueError("splatB inputs disagree on N")
    P = gsid.numel()
    dev = us.device
    o = dict(dtype=torch.float32, device=dev)
    if moments_only:
        du = dc = da = dcol = None
        moments = torch.empty((N, 9), **o)
    else:
        du = torch.empty((N, 1, 2), **o); dc = torch.empty((N, 1, 3), **o)
        da = torch.empty((N, 1, 1), **o); dcol = torch.empty((N, 1, 3), **o)
        moments = None
    lib = _L()
    with torch.cuda.device(dev):
        recs = _cached_records(gsid, (us, cinv2ds, alphas, colors)) if P > 0 else None
        # with cached records only the [N,9] moment rows are needed from the workspace
        ws_bytes = lib.gsb_splat_backward_workspace_bytes(N, H, W, 0 if recs else P)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.gsb_splat_backward(H, W, N, P, _ptr(us), _ptr(cinv2ds), _ptr(alphas), _ptr(colors),
                                          _ptr(contrib), _ptr(final_tau), _ptr(ranges), _ptr(gsid), _ptr(dl),
                                          recs, _ptr(ws), ws_bytes, _ptr(du), _ptr(dc), _ptr(da), _ptr(dcol),
                                          _ptr(moments), _stream()), lib)
    if moments_only:
        return moments
    # tagged so that the reference's `dloss_d* @ jacobian` chain takes the streaming matmul
    return [_jac(du), _jac(dc), _jac(da), _jac(dcol)]


# ---------------------------------------------------------------------------------------
# Extensions (not in the reference module): the fused per-Gaussian path, SURVEY 8f row N1.
def preprocess(pws, rots, scales, shs, Rcw, tcw, twc, focal_x, focal_y, center_x, center_y, width, height,
               alphas=None):
    """project + computeCov3D + computeCov2D + sh2Color + inverseCov2D (calc_J=False) in one
    kernel.  -> [us[N,2], cinv2ds[N,3], colors[N,3], depths[N], areas[N,2] int32], ready for
    `splat`.  With alphas[N] a sixth output: the packed per-Gaussian records [N,12] for
    `splat(..., records=...)`."""
    pws = _chk(pws, "pws", last=3, ndim=2); rots = _chk(rots, "rots", last=4, ndim=2)
    scales = _chk(scales, "scales", last=3, ndim=2); shs = _chk(shs, "shs", ndim=2)
    Rcw = _chk(Rcw, "Rcw"); tcw = _chk(tcw, "tcw"); twc = _chk(twc, "twc")
    _same_device(pws, rots, scales, shs, Rcw, tcw, twc)
    N = pws.shape[0]
    if not (rots.shape[0] == N and scales.shape[0] == N and shs.shape[0] == N):
        raise ValueError("preprocess inputs disagree on N")
    if shs.shape[1] % 3 != 0 or shs.shape[1] // 3 not in (1, 4, 9, 16):
        raise ValueError("shs must be [N, 3k] with k in {1,4,9,16}, got %s" % (tuple(shs.shape),))
    if Rcw.numel() != 9 or tcw.numel() != 3 or twc.numel() != 3:
        raise ValueError("Rcw must have 9, tcw and twc 3 elements")
    k = shs.shape[1] // 3
    o = dict(dtype=torch.float32, device=pws.device)
    us = torch.empty((N, 2), **o); cinv = torch.empty((N, 3), **o); col = torch.empty((N, 3), **o)
    depths = torch.empty((N,), **o); areas = torch.empty((N, 2), dtype=torch.int32, device=pws.device)
    records = None
    if alphas is not None:
        alphas = _chk(alphas, "alphas")
        if alphas.numel() != N:
            raise ValueError("alphas must have N elements")
        records = torch.empty((N, 12), **o)
    lib = _L()
    with torch.cuda.device(pws.device):
        _lib.check(lib.gsb_preprocess_forward(
            N, k, _ptr(pws), _ptr(rots), _ptr(scales), _ptr(shs), _ptr(Rcw), _ptr(tcw), _ptr(twc),
            float(focal_x), float(focal_y), float(center_x), float(center_y), float(width), float(height),
            _ptr(us), _ptr(cinv), _ptr(col), _ptr(depths), _ptr(areas), _ptr(alphas), _ptr(records), _stream()), lib)
    return [us, cinv, col, depths, areas] + ([records] if records is not None else [])


def preprocessB(pws, rots, scales, shs, Rcw, tcw, twc, focal_x, focal_y, center_x, center_y, width, height,
                dloss_dus, dloss_dcinv2ds, dloss_dcolors, moments=None, cinv2ds=None, compact=False):
    """Vector-Jacobian products of the five per-Gaussian stages (== the torch.bmm chain of
    gsmodel.py:72-85).  -> [dloss_dpws[N,3], dloss_dshs[N,3k], dloss_dscales[N,3], dloss_drots[N,4]]
    With moments[N,9] (from `splatB(..., moments_only=True)`) and the forward's cinv2ds[N,3] the
    three dloss_d* arguments are ignored (pass None) and two more outputs follow:
    dloss_dus[N,2], dloss_dalphas[N].
    compact (with moments): dloss_dshs is not produced (None) -- the caller expands it from the
    views' dL/dcolor = moments[:, 6:9] with `sh_grad_expand` -- and the other gradients share one
    flat bucket [dpws | dscales | drots | dalphas] of 11 N floats (one collective sums it)."""
    pws = _chk(pws, "pws", last=3, ndim=2); rots = _chk(rots, "rots", last=4, ndim=2)
    scales = _chk(scales, "scales", last=3, ndim=2); shs = _chk(shs, "shs", ndim=2)
    Rcw = _chk(Rcw, "Rcw"); tcw = _chk(tcw, "tcw"); twc = _chk(twc, "twc")
    N = pws.shape[0]
    if moments is not None:
        moments = _chk(moments, "moments", last=9, ndim=2); cinv2ds = _chk(cinv2ds, "cinv2ds", last=3, ndim=2)
        _same_device(pws, rots, scales, shs, Rcw, tcw, twc, moments, cinv2ds)
        if moments.shape[0] != N or cinv2ds.shape[0] != N:
            raise ValueError("moments / cinv2ds must have N rows")
        gu = gc = gcol = None
    else:
        gu = _chk(dloss_dus, "dloss_dus", last=2); gc = _chk(dloss_dcinv2ds, "dloss_dcinv2ds", last=3)
        gcol = _chk(dloss_dcolors, "dloss_dcolors", last=3)
        _same_device(pws, rots, scales, shs, Rcw, tcw, twc, gu, gc, gcol)
        if not (gu.numel() == 2 * N and gc.numel() == 3 * N and gcol.numel() == 3 * N):
            raise ValueError("preprocessB inputs disagree on N")
    if not (rots.shape[0] == N and scales.shape[0] == N and shs.shape[0] == N):
        raise ValueError("preprocessB inputs disagree on N")
    if shs.shape[1] % 3 != 0 or shs.shape[1] // 3 not in (1, 4, 9, 16):
        raise ValueError("shs must be [N, 3k] with k in {1,4,9,16}, got %s" % (tuple(shs.shape),))
    k = shs.shape[1] // 3
    o = dict(dtype=torch.float32, device=pws.device)
    # one flat bucket [dshs | drots | dpws | dscales]: a multi-GPU caller can all-reduce the four
    # gradients in place with a single collective (parallel.allreduce_grads); every view starts
    # 16-byte aligned (3k*N and 4*N floats are multiples of 4 elements when N % 4 == 0; the
    # kernels fall back to scalar stores otherwise)
    if compact:
        if moments is None:
            raise ValueError("compact=True needs the moment rows")
        bucket = torch.empty((N * 11,), **o)
        gsh = None
        gpw = bucket[:3 * N].view(N, 3)
        gs = bucket[3 * N:6 * N].view(N, 3)
        gq = bucket[6 * N:10 * N].view(N, 4)
        dal = bucket[10 * N:]
    else:
        bucket = torch.empty((N * (3 * k + 10),), **o)
        gsh = bucket[:N * 3 * k].view(N, 3 * k)
        gq = bucket[N * 3 * k:N * (3 * k + 4)].view(N, 4)
        gpw = bucket[N * (3 * k + 4):N * (3 * k + 7)].view(N, 3)
        gs = bucket[N * (3 * k + 7):].view(N, 3)
        dal = torch.empty((N,), **o) if moments is not None else None
    dus = torch.empty((N, 2), **o) if moments is not None else None
    lib = _L()
    with torch.cuda.device(pws.device):
        _lib.check(lib.gsb_preprocess_backward(
            N, k, _ptr(pws), _ptr(rots), _ptr(scales), _ptr(shs), _ptr(Rcw), _ptr(tcw), _ptr(twc),
            float(focal_x), float(focal_y), float(center_x), float(center_y), float(width), float(height),
            _ptr(gu), _ptr(gc), _ptr(gcol), _ptr(gpw), _ptr(gsh), _ptr(gs), _ptr(gq), _ptr(moments), _ptr(cinv2ds),
            _ptr(dus), _ptr(dal), _stream()), lib)
    return [gpw, gsh, gs, gq] + ([dus, dal] if moments is not None else [])


def sh_grad_expand(pws, twcs, dloss_dcolors, sh_dim3, out=None):
    """dloss_dshs[N, 3k] = sum over the V views of Y(dir_v) (x) dloss_dcolors[v]  (gsb_sh_grad_expand;
    colour is linear in sh, kernel.cu:735-774).  pws[N,3], twcs[V,3] (camera centres),
    dloss_dcolors[V,N,3]."""
    pws = _chk(pws, "pws", last=3, ndim=2); twcs = _chk(twcs, "twcs", last=3, ndim=2)
    g = _chk(dloss_dcolors, "dloss_dcolors", last=3, ndim=3)
    _same_device(pws, twcs, g)
    N, V, k = pws.shape[0], twcs.shape[0], int(sh_dim3)
    if g.shape[0] != V or g.shape[1] != N or k not in (1, 4, 9, 16):
        raise ValueError("sh_grad_expand: need twcs[V,3], dloss_dcolors[V,N,3], sh_dim3 in {1,4,9,16}")
    if out is None:
        out = torch.empty((N, 3 * k), dtype=torch.float32, device=pws.device)
    lib = _L()
    with torch.cuda.device(pws.device):
        _lib.check(lib.gsb_sh_grad_expand(N, k, V, _ptr(pws), _ptr(twcs), _ptr(g), _ptr(out), _stream()), lib)
    return out

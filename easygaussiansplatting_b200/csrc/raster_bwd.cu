// Backward tile rasterizer ("drawB").  Replaces reference kernel.cu:809-950.
//
// (The kernel itself is k_draw_bwd2 in raster_bwd2.cu; this file holds the algorithm notes, the
// moment finalisation and the launcher.)
// Same tiling and record pipeline as the forward (raster_fwd2.cu), walked back to front starting at the last
// batch any pixel of the tile needs (max contrib).  Each pixel replays its saved
// (final_tau, contrib) state: tau <- tau / (1 - alpha'), dL/dalpha' = tau (c - gamma_next) . dL/dgamma,
// with gamma_next kept as the scalar s = dL/dgamma . gamma_next (s <- alpha' (dL/dgamma . c) + (1 - alpha') s).
//
// Geometry gradients are linear in per-pixel moments, so a pixel only produces nine numbers
//   w dx, w dy, w dx^2, w dx dy, w dy^2   (w = dL/dalpha' * alpha'),  dL/dalpha' * g,  alpha' tau dL/dgamma_rgb
// which are summed over the 32 pixels of the warp with a SPLIT butterfly: at every level each
// lane keeps half of the values and trades the other half, so the 9 sums cost 12 shuffles
// (5+3+2+1+1) instead of 45, and end up in 9 different lanes.  Those 9 lanes then issue ONE
// predicated red.global.add into the Gaussian's contiguous 9-float moment row (36 B, two
// sectors) -- the reference issues 9 atomics per (pixel, record) into four arrays.
// finalize_splat_grads() turns moment rows into the four gradient tensors:
//   dL/du = -(A Sx + B Sy, B Sx + C Sy),  dL/dconic = (-Sxx/2, -Sxy, -Syy/2).
// Records that cannot touch the warp's rectangle are culled by the same ballot test as the
// forward; records where no pixel is active are skipped before any reduction.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tile_io.cuh"

namespace gsb {

constexpr int MOM = 9;  // floats per moment row

// moment rows [N,9] -> dloss_dus[N,2], dloss_dcinv2ds[N,3], dloss_dalphas[N], dloss_dcolors[N,3]
__global__ void __launch_bounds__(PG) k_finalize_grads(int N, const float *__restrict__ moments,
                                                       const float *__restrict__ cinv2ds,
                                                       float *__restrict__ dus, float *__restrict__ dcinv,
                                                       float *__restrict__ dalphas, float *__restrict__ dcolors) {
  __shared__ float sm[TileT<MOM>::FLOATS];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * PG;
  const int nv = min(PG, (int)(N - base));
  const bool valid = tid < nv;
  float A = 0.f, B = 0.f, Cc = 0.f, m[MOM];
  tile_fetch<3>(cinv2ds, base, nv, sm, tid);
  __syncthreads();
  if (valid) { const float *r = sm + tid * TileT<3>::S; A = r[0]; B = r[1]; Cc = r[2]; }
  __syncthreads();
  tile_fetch<MOM>(moments, base, nv, sm, tid);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MOM; i++) m[i] = valid ? sm[tid * TileT<MOM>::S + i] : 0.f;
  __syncthreads();
  {
    // untouched Gaussians stay exactly 0 even if their conic is inf/NaN (reference: no atomics ran)
    const bool none = (m[0] == 0.f) && (m[1] == 0.f);
    float *o = sm + tid * TileT<2>::S;
    o[0] = none ? 0.f : -(A * m[0] + B * m[1]);
    o[1] = none ? 0.f : -(B * m[0] + Cc * m[1]);
  }
  __syncthreads();
  tile_flush<2>(dus, base, nv, sm, tid);
  __syncthreads();
  { float *o = sm + tid * TileT<3>::S; o[0] = -0.5f * m[2]; o[1] = -m[3]; o[2] = -0.5f * m[4]; }
  __syncthreads();
  tile_flush<3>(dcinv, base, nv, sm, tid);
  __syncthreads();
  { float *o = sm + tid * TileT<3>::S; o[0] = m[6]; o[1] = m[7]; o[2] = m[8]; }
  __syncthreads();
  tile_flush<3>(dcolors, base, nv, sm, tid);
  if (valid) dalphas[base + tid] = m[5];
}

int launch_draw_backward(int H, int W, int N, const int32_t *ranges, const Rec *recs, const int32_t *gsid,
                         const int32_t *contrib,
                         const float *final_tau, const float *dloss_dgammas, const float *cinv2ds,
                         float *moments, int *tile_counter, int *work_counter, float *dloss_dus,
                         float *dloss_dcinv2ds, float *dloss_dalphas, float *dloss_dcolors, bool finalize,
                         cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  if (gx <= 0 || gy <= 0 || N <= 0) return 0;
  GSB_CUDA_TRY(cudaMemsetAsync(moments, 0, sizeof(float) * MOM * (size_t)N, st));
  if (recs != nullptr) {
    // GSB_BWD_VARIANT=2 selects the older warp-reduction kernel (raster_bwd2.cu) for A/B runs
    static const int variant = [] {
      const char *e = getenv("GSB_BWD_VARIANT");
      return e != nullptr ? atoi(e) : 4;
    }();
    int rc = variant == 2   ? launch_draw_bwd2_kernel(H, W, ranges, recs, gsid, contrib, final_tau, dloss_dgammas,
                                                      moments, tile_counter, st)
                            : launch_draw_bwd4_kernel(H, W, ranges, recs, gsid, contrib, final_tau, dloss_dgammas,
                                                      moments, work_counter, st);
    if (rc) return rc;
  }
  GSB_CUDA_TRY(cudaGetLastError());
  if (finalize) {  // otherwise the caller consumes the moment rows (gsb_preprocess_backward)
    ProfScope ps(K_FINALIZE, st);
    k_finalize_grads<<<(N + PG - 1) / PG, PG, 0, st>>>(N, moments, cinv2ds, dloss_dus, dloss_dcinv2ds,
                                                       dloss_dalphas, dloss_dcolors);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

// Backward tile rasterizer ("drawB").  Replaces reference kernel.cu:809-950.
//
// Same tiling as the forward (CTA = 16x16 tile, warp = 8x4 pixels, lane = pixel) and the
// same cp.async.bulk + mbarrier record pipeline, walked back to front starting at the last
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
#include "common.cuh"
#include "kernels.h"
#include "tile_io.cuh"

namespace gsb {

constexpr int BWD_BATCH = 128;
constexpr int MOM = 9;  // floats per moment row

// Sum 9 per-lane values over the warp.  Returns the total of value `slot_of_lane(lane)` in
// every lane (lanes 2k and 2k+1 hold the same one); 12 SHFL.
__device__ __forceinline__ float split_reduce9(const float (&v)[9], int lane) {
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
  const unsigned F = 0xffffffffu;
  float a[5], b[3], c[2], d;
  // level 1: lower half keeps 0..4, upper half keeps 5..8
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = (u16 ? v[i + 5] : v[i]) + __shfl_xor_sync(F, u16 ? v[i] : v[i + 5], 16);
  a[4] = (u16 ? 0.f : v[4]) + __shfl_xor_sync(F, u16 ? v[4] : 0.f, 16);
  // level 2: keeps 0..2 | 3..4
#pragma unroll
  for (int i = 0; i < 2; i++) b[i] = (u8 ? a[i + 3] : a[i]) + __shfl_xor_sync(F, u8 ? a[i] : a[i + 3], 8);
  b[2] = (u8 ? 0.f : a[2]) + __shfl_xor_sync(F, u8 ? a[2] : 0.f, 8);
  // level 3: keeps 0..1 | 2
  c[0] = (u4 ? b[2] : b[0]) + __shfl_xor_sync(F, u4 ? b[0] : b[2], 4);
  c[1] = (u4 ? 0.f : b[1]) + __shfl_xor_sync(F, u4 ? b[1] : 0.f, 4);
  // level 4: keeps 0 | 1
  d = (u2 ? c[1] : c[0]) + __shfl_xor_sync(F, u2 ? c[0] : c[1], 2);
  // level 5
  d += __shfl_xor_sync(F, d, 1);
  return d;
}
// which of the 9 values a lane ends up with (-1: a padding slot, or the odd twin lane)
__device__ __forceinline__ int slot_of_lane(int lane) {
  if (lane & 1) return -1;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
  int local;
  if (!u8) local = u4 ? (u2 ? -1 : 2) : (u2 ? 1 : 0);
  else local = u4 ? -1 : (u2 ? 4 : 3);
  if (local < 0) return -1;
  const int g = local + (u16 ? 5 : 0);
  return g < MOM ? g : -1;
}

__global__ void __launch_bounds__(256) k_draw_bwd(
    int W, int H, int gx, const int2 *__restrict__ ranges, const Rec *__restrict__ recs,
    const int32_t *__restrict__ contrib, const float *__restrict__ final_tau,
    const float *__restrict__ dloss_dgammas, float *__restrict__ moments) {
  __shared__ Rec sbuf[2][BWD_BATCH];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ int s_wmax[8];

  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rx0 = tx * TILE + (warp & 1) * 8, ry0 = ty * TILE + (warp >> 1) * 4;
  const int px = rx0 + (lane & 7), py = ry0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)py * W + px;

  const int2 range = __ldg(ranges + tile);
  const int len = range.y - range.x;
  if (len <= 0) return;

  int cont = 0;
  float tau = 0.f, dlr = 0.f, dlg = 0.f, dlb = 0.f;
  if (inside) {
    cont = min(__ldg(contrib + pix), len);
    tau = __ldg(final_tau + pix);
    dlr = __ldg(dloss_dgammas + pix);
    dlg = __ldg(dloss_dgammas + HW + pix);
    dlb = __ldg(dloss_dgammas + 2 * HW + pix);
  }
  int wmax = cont;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) s_wmax[warp] = wmax;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  int bmax = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) bmax = max(bmax, s_wmax[i]);
  if (bmax <= 0) return;
  const int nbn = (bmax + BWD_BATCH - 1) / BWD_BATCH;  // batches [0, nbn) are needed
  const Rec *src = recs + range.x;
  if (tid == 0) {
    for (int bi = 0; bi < 2 && bi < nbn; bi++) {
      const int b = nbn - 1 - bi;
      const uint32_t bytes = (uint32_t)min(BWD_BATCH, len - b * BWD_BATCH) * (uint32_t)sizeof(Rec);
      mbar_expect_tx(&mbar[bi], bytes);
      bulk_g2s(&sbuf[bi][0], src + (size_t)b * BWD_BATCH, bytes, &mbar[bi]);
    }
  }

  const float fpx = (float)px, fpy = (float)py;
  const float bx0 = (float)rx0, bx1 = (float)(rx0 + 7), by0 = (float)ry0, by1 = (float)(ry0 + 3);
  const int slot = slot_of_lane(lane);
  float *const mom_lane = moments + (slot >= 0 ? slot : 0);
  float sdot = 0.f;  // dL/dgamma . gamma_next

  for (int bi = 0; bi < nbn; bi++) {
    const int b = nbn - 1 - bi;
    const int s = bi & 1;
    mbar_wait(&mbar[s], (bi >> 1) & 1);
    const int nrec = min(BWD_BATCH, len - b * BWD_BATCH);
    if (b * BWD_BATCH < wmax) {
      for (int c0 = ((nrec - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        const int j = c0 + lane;
        bool hit = false;
        if (j < nrec && b * BWD_BATCH + j < wmax)
          hit = rec_can_touch(sbuf[s][j].q0, sbuf[s][j].q1, bx0, bx1, by0, by1);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = 31 - __clz(mask);  // back to front
          mask &= ~(1u << k);
          const Rec *r = &sbuf[s][c0 + k];
          const int idx = b * BWD_BATCH + c0 + k;
          const float4 q0 = r->q0, q1 = r->q1;
          const float dx = q0.x - fpx, dy = q0.y - fpy;
          float g;
          const float ap = alpha_prime(q1, dx, dy, &g);
          const bool active = (idx < cont) && (ap >= ALPHA_SKIP);
          if (!__any_sync(0xffffffffu, active)) continue;
          const float4 q2 = r->q2;
          // inactive lanes contribute exact zeros through the two scale factors
          float dl_dap = 0.f, wc = 0.f;
          if (active) {
            tau = tau * rcp_approx(1.0f - ap);  // 1 - alpha' >= 0.01: no denormal handling needed
            const float dc = fmaf(dlr, q2.x, fmaf(dlg, q2.y, dlb * q2.z));
            const float diff = dc - sdot;
            dl_dap = tau * diff;
            sdot = fmaf(ap, diff, sdot);
            wc = ap * tau;
          }
          const float w = dl_dap * ap;
          const float wdx = w * dx, wdy = w * dy;
          const float v[9] = {wdx, wdy, wdx * dx, wdx * dy, wdy * dy, dl_dap * g, wc * dlr, wc * dlg, wc * dlb};
          const float tot = split_reduce9(v, lane);
          if (slot >= 0) atomicAdd(mom_lane + (size_t)__float_as_int(q2.w) * MOM, tot);
        }
      }
    }
    __syncthreads();  // every warp is done with stage s
    if (tid == 0 && bi + 2 < nbn) {
      const int b2 = nbn - 1 - (bi + 2);
      const uint32_t bytes = (uint32_t)min(BWD_BATCH, len - b2 * BWD_BATCH) * (uint32_t)sizeof(Rec);
      fence_proxy_async();
      mbar_expect_tx(&mbar[s], bytes);
      bulk_g2s(&sbuf[s][0], src + (size_t)b2 * BWD_BATCH, bytes, &mbar[s]);
    }
  }
}

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
                         float *moments, int *tile_counter, float *dloss_dus, float *dloss_dcinv2ds,
                         float *dloss_dalphas, float *dloss_dcolors, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  if (gx <= 0 || gy <= 0 || N <= 0) return 0;
  GSB_CUDA_TRY(cudaMemsetAsync(moments, 0, sizeof(float) * MOM * (size_t)N, st));
  if (recs != nullptr) {
    int rc = launch_draw_bwd2_kernel(H, W, ranges, recs, gsid, contrib, final_tau, dloss_dgammas, moments,
                                     tile_counter, st);
    if (rc) return rc;
  }
  GSB_CUDA_TRY(cudaGetLastError());
  {
    ProfScope ps(K_FINALIZE, st);
    k_finalize_grads<<<(N + PG - 1) / PG, PG, 0, st>>>(N, moments, cinv2ds, dloss_dus, dloss_dcinv2ds,
                                                       dloss_dalphas, dloss_dcolors);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

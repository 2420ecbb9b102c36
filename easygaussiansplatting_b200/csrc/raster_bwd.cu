// Backward tile rasterizer ("drawB").  Replaces reference kernel.cu:809-950.
//
// Same tiling as the forward (CTA = 16x16 tile, warp = 8x4 pixels, lane = pixel) and the
// same cp.async.bulk + mbarrier record pipeline, walked back to front starting at the last
// batch any pixel of the tile needs (max contrib).  Each pixel replays its saved
// (final_tau, contrib) state: tau <- tau / (1 - alpha'), dL/dalpha' = tau (c - gamma_next) . dL/dgamma,
// with gamma_next kept as the scalar s = dL/dgamma . gamma_next (s <- alpha' (dL/dgamma . c) + (1 - alpha') s).
// Geometry gradients are linear in per-pixel moments, so a pixel only produces
//   w dx, w dy, w dx^2, w dx dy, w dy^2   (w = dL/dalpha' * alpha'),  dL/dalpha' * g,  alpha' tau dL/dgamma_rgb
// and the conic is applied once per (warp, record) after the warp reduction.
// The reference issues 9 global atomics per (pixel, record); here the 32 pixels of a warp are
// summed with shuffles first (<= 9 atomics per (warp, record)), records that cannot touch
// the warp's rectangle are culled by the same ballot test as the forward, and records where
// no pixel is active are skipped before any reduction.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

constexpr int BWD_BATCH = 128;

__device__ __forceinline__ float warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(256) k_draw_bwd(
    int W, int H, int gx, const int2 *__restrict__ ranges, const Rec *__restrict__ recs,
    const int32_t *__restrict__ contrib, const float *__restrict__ final_tau,
    const float *__restrict__ dloss_dgammas, float *__restrict__ dloss_dus,
    float *__restrict__ dloss_dcinv2ds, float *__restrict__ dloss_dalphas,
    float *__restrict__ dloss_dcolors) {
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
        if (j < nrec && b * BWD_BATCH + j < wmax) {
          const float4 q0 = sbuf[s][j].q0;
          hit = (q0.x + q0.z >= bx0) && (q0.x - q0.z <= bx1) && (q0.y + q0.w >= by0) && (q0.y - q0.w <= by1);
        }
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
          float v[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (active) {
            tau = __fdividef(tau, 1.0f - ap);
            const float dc = fmaf(dlr, q2.x, fmaf(dlg, q2.y, dlb * q2.z));
            const float diff = dc - sdot;
            const float dl_dap = tau * diff;
            sdot = fmaf(ap, diff, sdot);
            const float w = dl_dap * ap;
            const float wc = ap * tau;
            const float wdx = w * dx, wdy = w * dy;
            v[0] = wdx; v[1] = wdy; v[2] = wdx * dx; v[3] = wdx * dy; v[4] = wdy * dy;
            v[5] = dl_dap * g;
            v[6] = wc * dlr; v[7] = wc * dlg; v[8] = wc * dlb;
          }
#pragma unroll
          for (int i = 0; i < 9; i++) v[i] = warp_sum(v[i]);
          if (lane == 0) {
            const int gid = __float_as_int(q2.w);
            const float A = q1.x * (-2.0f / LOG2E), B = q1.y * (-1.0f / LOG2E), C = q1.z * (-2.0f / LOG2E);
            atomicAdd(dloss_dus + 2 * (size_t)gid + 0, -(A * v[0] + B * v[1]));
            atomicAdd(dloss_dus + 2 * (size_t)gid + 1, -(B * v[0] + C * v[1]));
            atomicAdd(dloss_dcinv2ds + 3 * (size_t)gid + 0, -0.5f * v[2]);
            atomicAdd(dloss_dcinv2ds + 3 * (size_t)gid + 1, -v[3]);
            atomicAdd(dloss_dcinv2ds + 3 * (size_t)gid + 2, -0.5f * v[4]);
            atomicAdd(dloss_dalphas + gid, v[5]);
            atomicAdd(dloss_dcolors + 3 * (size_t)gid + 0, v[6]);
            atomicAdd(dloss_dcolors + 3 * (size_t)gid + 1, v[7]);
            atomicAdd(dloss_dcolors + 3 * (size_t)gid + 2, v[8]);
          }
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

int launch_draw_backward(int H, int W, const int32_t *ranges, const Rec *recs, const int32_t *contrib,
                         const float *final_tau, const float *dloss_dgammas, float *dloss_dus,
                         float *dloss_dcinv2ds, float *dloss_dalphas, float *dloss_dcolors,
                         cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  if (gx <= 0 || gy <= 0) return 0;
  ProfScope ps(K_DRAW_BWD, st);
  k_draw_bwd<<<gx * gy, 256, 0, st>>>(W, H, gx, reinterpret_cast<const int2 *>(ranges), recs, contrib,
                                      final_tau, dloss_dgammas, dloss_dus, dloss_dcinv2ds, dloss_dalphas,
                                      dloss_dcolors);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

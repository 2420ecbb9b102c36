// Forward tile rasterizer, variant 2: persistent CTAs, two pixels per lane, Blackwell
// packed-fp32 math.
//
// Both rasterizer kernels are instruction-issue bound on dense scenes (ncu: 86-91 % issue-slot
// utilisation, DRAM 5 %), so the lever is instructions per evaluated pixel.  sm_100 has packed
// fp32 arithmetic (PTX fma/mul/add .f32x2, SASS FFMA2/FMUL2/FADD2: two independent fp32 results
// per instruction).  A lane owns two horizontally adjacent pixels (same dy), a warp owns an 8x8
// pixel block, a CTA of 4 warps renders one 16x16 tile at a time, and the quadratic form, the
// alpha multiply and the whole compositing update run as f32x2 operations on the pixel pair;
// only min/max, ex2 and the compares stay scalar.
// For sparse frames (few patches per tile, thousands of empty tiles: the HBM-bound corner of
// BASELINE config 4) the grid is persistent -- a few CTAs per SM pulling tile indices from an
// atomic counter -- so the kernel is bounded by HBM writes instead of CTA launch rate; dense
// frames use one CTA per tile (the caller passes tile_counter = nullptr).
// Records: one 48-byte record per Gaussian (binning.cu k_pack); a stage of 128 records is filled
// by a gather -- thread t takes the t-th Gaussian id of the batch from the sorted patch list and
// issues three 16-byte cp.async tracked by the stage's mbarrier (common.cuh gather_record).
// Warp-level exact culling (rec_can_touch) drops records that cannot reach the warp's 8x8 block.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace gsb {

constexpr int DRAW2_BATCH = 128;
#ifndef FWD2_PREFETCH_IDS
#define FWD2_PREFETCH_IDS 1
#endif

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }

template <bool PERSIST>
__global__ void __launch_bounds__(128) k_draw2(int W, int H, int gx, int T, const int2 *__restrict__ ranges,
                                               const Rec *__restrict__ recs, const int32_t *__restrict__ gsid,
                                               float *__restrict__ image,
                                               int32_t *__restrict__ contrib, float *__restrict__ final_tau,
                                               int *__restrict__ tile_counter) {
  __shared__ Rec sbuf[2][DRAW2_BATCH];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ int s_tile[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t HW = (size_t)H * W;
  const bool vec2 = (W & 1) == 0;  // pixel pairs are 8-byte aligned in every plane
  if (tid == 0) {
    mbar_init(&mbar[0], GATHER_ARRIVALS);  // see gather_record (common.cuh)
    mbar_init(&mbar[1], GATHER_ARRIVALS);
    fence_mbar_init();
  }
  uint32_t ph0 = 0, ph1 = 0;  // completed phases of the two stages (block-uniform)

  for (int it = 0;; it++) {
    int tile;
    if (PERSIST) {  // persistent grid: pull the next tile from the queue
      if (tid == 0) s_tile[it & 1] = atomicAdd(tile_counter, 1);
      __syncthreads();
      tile = s_tile[it & 1];
    } else {  // classic grid: one tile per CTA (dense frames: the hardware scheduler overlaps
      if (it > 0) break;  // a CTA's start-up latency with its neighbours' compute)
      __syncthreads();
      tile = blockIdx.x;
    }
    if (tile >= T) break;
    const int tx = tile % gx, ty = tile / gx;
    // warp -> 8x8 block, lane -> (row, pixel pair)
    const int rx0 = tx * TILE + (warp & 1) * 8, ry0 = ty * TILE + (warp >> 1) * 8;
    const int px = rx0 + 2 * (lane & 3), py = ry0 + (lane >> 2);
    const bool in0 = px < W && py < H, in1 = px + 1 < W && py < H;
    const size_t pix = (size_t)py * W + px;

    const int2 range = __ldg(ranges + tile);
    const int len = range.y - range.x;
    float2 tau = f2s(0.f), cr = f2s(0.f), cg = f2s(0.f), cb = f2s(0.f);
    int cont0 = 0, cont1 = 0;
    if (len > 0) {  // (a tile without patches keeps image 0, contrib 0, tau 0: kernel.cu:182-183)
      const int nb = (len + DRAW2_BATCH - 1) / DRAW2_BATCH;
      // stage fill = gather: thread t copies the record of the t-th patch of the batch from the
      // per-Gaussian record array (L2 resident) with three 16-byte cp.async
      const int32_t *ids = gsid + range.x;
      for (int b = 0; b < 2 && b < nb; b++) {
        const bool v = b * DRAW2_BATCH + tid < len;
        gather_record(&sbuf[b][0], recs, v ? __ldg(ids + b * DRAW2_BATCH + tid) : 0, v,
                      min(DRAW2_BATCH, len - b * DRAW2_BATCH), &mbar[b], tid);
      }
#if FWD2_PREFETCH_IDS
      int g_pref = (2 * DRAW2_BATCH + tid < len) ? __ldg(ids + 2 * DRAW2_BATCH + tid) : 0;  // id for batch 2
#endif
      const float2 npx = f2(-(float)px, -(float)(px + 1));
      const float fpy = (float)py;
      const float bx0 = (float)rx0, bx1 = (float)(rx0 + 7), by0 = (float)ry0, by1 = (float)(ry0 + 7);
      // a pixel is finished exactly when tau < 1e-4; pixels outside the image start finished
      tau = f2(in0 ? 1.0f : 0.0f, in1 ? 1.0f : 0.0f);
      bool warp_done = __all_sync(0xffffffffu, tau.x < TAU_STOP && tau.y < TAU_STOP);

      int b = 0;
      for (; b < nb; b++) {
        const int s = b & 1;
        if (!PERSIST) {
          mbar_wait(&mbar[s], (b >> 1) & 1);  // one tile per CTA: the phase follows the batch index
        } else if (s == 0) {
          mbar_wait(&mbar[0], ph0 & 1); ph0++;
        } else {
          mbar_wait(&mbar[1], ph1 & 1); ph1++;
        }
        const int nrec = min(DRAW2_BATCH, len - b * DRAW2_BATCH);
        if (!warp_done) {
          for (int c0 = 0; c0 < nrec; c0 += 32) {
            const int j = c0 + lane;
            bool hit = false;
            if (j < nrec) hit = rec_can_touch(sbuf[s][j].q0, sbuf[s][j].q1, bx0, bx1, by0, by1);
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
              const int k = __ffs(mask) - 1;
              mask &= mask - 1;
              const Rec *r = &sbuf[s][c0 + k];
              const float4 q0 = r->q0, q1 = r->q1;
              // log2 g = (a dx + b dy) dx + (c dy) dy   -- same operation sequence as alpha_prime()
              const float2 dx = __fadd2_rn(f2s(q0.x), npx);
              const float dy = q0.y - fpy;
              const float cdy2 = (q1.z * dy) * dy;
              const float2 t = __ffma2_rn(f2s(q1.y), f2s(dy), __fmul2_rn(f2s(q1.x), dx));
              const float2 p = __ffma2_rn(t, dx, f2s(cdy2));
              const float g0 = ex2_approx(fminf(p.x, 0.0f)), g1 = ex2_approx(fminf(p.y, 0.0f));
              const float2 ag = __fmul2_rn(f2s(q0.w), f2(g0, g1));
              const float ap0 = fminf(ALPHA_CLAMP, ag.x), ap1 = fminf(ALPHA_CLAMP, ag.y);
              const bool c0p = (tau.x >= TAU_STOP) && (ap0 >= ALPHA_SKIP);
              const bool c1p = (tau.y >= TAU_STOP) && (ap1 >= ALPHA_SKIP);
              if (c0p || c1p) {
                const float4 q2 = r->q2;
                // a pixel that skips this record composites alpha' = 0: w = 0, tau unchanged
                const float2 e = f2(c0p ? ap0 : 0.0f, c1p ? ap1 : 0.0f);
                const float2 w = __fmul2_rn(tau, e);
                cr = __ffma2_rn(w, f2s(q2.x), cr);
                cg = __ffma2_rn(w, f2s(q2.y), cg);
                cb = __ffma2_rn(w, f2s(q2.z), cb);
                tau = __fmul2_rn(tau, __fadd2_rn(f2s(1.0f), f2(-e.x, -e.y)));
                const int idx = b * DRAW2_BATCH + c0 + k + 1;
                if (c0p) cont0 = idx;
                if (c1p) cont1 = idx;
              }
            }
            warp_done = __all_sync(0xffffffffu, tau.x < TAU_STOP && tau.y < TAU_STOP);
            if (warp_done) break;
          }
        }
        // all warps are past stage s -> it may be refilled; also the tile-wide early out
        const int all_done = __syncthreads_and(warp_done ? 1 : 0);
        if (all_done) break;
        if (b + 2 < nb) {
#if FWD2_PREFETCH_IDS
          gather_record(&sbuf[s][0], recs, g_pref, (b + 2) * DRAW2_BATCH + tid < len,
                        min(DRAW2_BATCH, len - (b + 2) * DRAW2_BATCH), &mbar[s], tid);
          g_pref = ((b + 3) * DRAW2_BATCH + tid < len) ? __ldg(ids + (b + 3) * DRAW2_BATCH + tid) : 0;
#else
          const bool v2 = (b + 2) * DRAW2_BATCH + tid < len;
          gather_record(&sbuf[s][0], recs, v2 ? __ldg(ids + (b + 2) * DRAW2_BATCH + tid) : 0, v2,
                        min(DRAW2_BATCH, len - (b + 2) * DRAW2_BATCH), &mbar[s], tid);
#endif
        }
      }
      // early exit: the next stage's gather is in flight into shared memory -- consume it
      if (b + 1 < nb) {
        if (!PERSIST) {
          mbar_wait(&mbar[(b + 1) & 1], ((b + 1) >> 1) & 1);
        } else if (((b + 1) & 1) == 0) {
          mbar_wait(&mbar[0], ph0 & 1); ph0++;
        } else {
          mbar_wait(&mbar[1], ph1 & 1); ph1++;
        }
      }
    }
    if (in0 && in1 && vec2) {
      *reinterpret_cast<float2 *>(image + pix) = cr;
      *reinterpret_cast<float2 *>(image + HW + pix) = cg;
      *reinterpret_cast<float2 *>(image + 2 * HW + pix) = cb;
      *reinterpret_cast<int2 *>(contrib + pix) = make_int2(cont0, cont1);
      *reinterpret_cast<float2 *>(final_tau + pix) = tau;
    } else {
      if (in0) {
        image[pix] = cr.x; image[HW + pix] = cg.x; image[2 * HW + pix] = cb.x;
        contrib[pix] = cont0; final_tau[pix] = tau.x;
      }
      if (in1) {
        image[pix + 1] = cr.y; image[HW + pix + 1] = cg.y; image[2 * HW + pix + 1] = cb.y;
        contrib[pix + 1] = cont1; final_tau[pix + 1] = tau.y;
      }
    }
  }
}

int persistent_grid(int T, int ctas_per_sm) {
  int dev = 0, sms = 148;  // queried per call: per current device, thread-safe
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    sms = 148;
  const long long g = (long long)sms * ctas_per_sm;
  return (int)(T < g ? T : g);
}

int launch_draw(int H, int W, const int32_t *ranges, const Rec *recs, const int32_t *gsid, float *image,
                 int32_t *contrib, float *final_tau, int *tile_counter, int *work_counter, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  if (gx <= 0 || gy <= 0) return 0;
  // GSB_FWD_VARIANT=2 selects this file's CTA-per-tile kernel for A/B runs; default: raster_fwd3.cu
  static const int variant = [] {
    const char *e = getenv("GSB_FWD_VARIANT");
    return e != nullptr ? atoi(e) : 3;
  }();
  if (variant != 2) return launch_draw3(H, W, ranges, recs, gsid, image, contrib, final_tau, work_counter, st);
  const int T = gx * gy;
  if (tile_counter != nullptr) GSB_CUDA_TRY(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
  ProfScope ps(K_DRAW, st);
  if (tile_counter != nullptr)
    k_draw2<true><<<persistent_grid(T, 12), 128, 0, st>>>(W, H, gx, T, reinterpret_cast<const int2 *>(ranges), recs,
                                                          gsid, image, contrib, final_tau, tile_counter);
  else
    k_draw2<false><<<T, 128, 0, st>>>(W, H, gx, T, reinterpret_cast<const int2 *>(ranges), recs, gsid, image,
                                      contrib, final_tau, tile_counter);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

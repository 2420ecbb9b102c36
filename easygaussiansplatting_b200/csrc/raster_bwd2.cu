// Backward tile rasterizer, variant 2: two pixels per lane + packed-fp32 (f32x2) math.
// See raster_fwd2.cu for the mapping (warp = 8x8 pixels, lane = one row position x 2 adjacent
// pixels, CTA = 4 warps = one 16x16 tile) and the record gather, raster_bwd.cu for the algorithm (back-to-front
// replay from (final_tau, contrib), nine per-pixel contributions, split-butterfly warp
// reduction into one RED per (warp, record) on the Gaussian's 36-byte moment row).
// The two pixels of a lane are added before the warp reduction, so the reduction and the
// atomic are paid once per 64 pixels instead of once per 32.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

constexpr int BWD2_BATCH = 128;
// A/B (benchmarks/ab_variants.py, config 2): CTAs/SM @ registers -> kernel ms:
// 12 @ 40 (spills) 0.693, 10 @ 47 0.669, 9 @ 55 0.661, 8 @ 61 0.667; and loading a batch's Gaussian
// ids right before the gather beats carrying them in a register across the batch (0.671 vs 0.678).
#ifndef BWD2_MINBLOCKS
#define BWD2_MINBLOCKS 9
#endif
#ifndef BWD2_PREFETCH_IDS
#define BWD2_PREFETCH_IDS 0
#endif
constexpr int MOM2 = 9;

__device__ __forceinline__ float2 g2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 g2s(float a) { return make_float2(a, a); }

// identical to raster_bwd.cu (kept local so both variants stay self-contained)
__device__ __forceinline__ float split_reduce9_v2(const float (&v)[9], int lane) {
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
  const unsigned F = 0xffffffffu;
  float a[5], b[3], c[2], d;
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = (u16 ? v[i + 5] : v[i]) + __shfl_xor_sync(F, u16 ? v[i] : v[i + 5], 16);
  a[4] = (u16 ? 0.f : v[4]) + __shfl_xor_sync(F, u16 ? v[4] : 0.f, 16);
#pragma unroll
  for (int i = 0; i < 2; i++) b[i] = (u8 ? a[i + 3] : a[i]) + __shfl_xor_sync(F, u8 ? a[i] : a[i + 3], 8);
  b[2] = (u8 ? 0.f : a[2]) + __shfl_xor_sync(F, u8 ? a[2] : 0.f, 8);
  c[0] = (u4 ? b[2] : b[0]) + __shfl_xor_sync(F, u4 ? b[0] : b[2], 4);
  c[1] = (u4 ? 0.f : b[1]) + __shfl_xor_sync(F, u4 ? b[1] : 0.f, 4);
  d = (u2 ? c[1] : c[0]) + __shfl_xor_sync(F, u2 ? c[0] : c[1], 2);
  d += __shfl_xor_sync(F, d, 1);
  return d;
}
__device__ __forceinline__ int slot_of_lane_v2(int lane) {
  if (lane & 1) return -1;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
  int local;
  if (!u8) local = u4 ? (u2 ? -1 : 2) : (u2 ? 1 : 0);
  else local = u4 ? -1 : (u2 ? 4 : 3);
  if (local < 0) return -1;
  const int g = local + (u16 ? 5 : 0);
  return g < MOM2 ? g : -1;
}

__global__ void __launch_bounds__(128, BWD2_MINBLOCKS) k_draw_bwd2(
    int W, int H, int gx, int T, const int2 *__restrict__ ranges, const Rec *__restrict__ recs,
    const int32_t *__restrict__ gsid, const int32_t *__restrict__ contrib, const float *__restrict__ final_tau,
    const float *__restrict__ dloss_dgammas, float *__restrict__ moments, int *__restrict__ tile_counter) {
  __shared__ Rec sbuf[2][BWD2_BATCH];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ int s_wmax[4];
  __shared__ int s_tile[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t HW = (size_t)H * W;
  const int slot = slot_of_lane_v2(lane);
  float *const mom_lane = moments + (slot >= 0 ? slot : 0);
  if (tid == 0) {
    mbar_init(&mbar[0], GATHER_ARRIVALS);  // see gather_record (common.cuh)
    mbar_init(&mbar[1], GATHER_ARRIVALS);
    fence_mbar_init();
  }
  uint32_t ph0 = 0, ph1 = 0;  // completed phases of the two stages (block-uniform)

  // persistent CTA: tiles are pulled from an atomic counter (see raster_fwd2.cu)
  for (int it = 0;; it++) {
    int tile;
    if (tile_counter != nullptr) {  // persistent grid: pull the next tile from the queue
      if (tid == 0) s_tile[it & 1] = atomicAdd(tile_counter, 1);
      __syncthreads();
      tile = s_tile[it & 1];
    } else {  // classic grid: one tile per CTA (dense frames: the hardware scheduler overlaps
      if (it > 0) break;  // a CTA's start-up latency with its neighbours' compute)
      __syncthreads();
      tile = blockIdx.x;
    }
    if (tile >= T) break;
    const int2 range = __ldg(ranges + tile);
    const int len = range.y - range.x;
    if (len <= 0) continue;
    const int tx = tile % gx, ty = tile / gx;
    const int rx0 = tx * TILE + (warp & 1) * 8, ry0 = ty * TILE + (warp >> 1) * 8;
    const int px = rx0 + 2 * (lane & 3), py = ry0 + (lane >> 2);
    const bool in0 = px < W && py < H, in1 = px + 1 < W && py < H;
    const size_t pix = (size_t)py * W + px;

    int cont0 = 0, cont1 = 0;
    float2 tau = g2s(0.f), dlr = g2s(0.f), dlg = g2s(0.f), dlb = g2s(0.f);
    if (in0) {
      cont0 = min(__ldg(contrib + pix), len);
      tau.x = __ldg(final_tau + pix);
      dlr.x = __ldg(dloss_dgammas + pix);
      dlg.x = __ldg(dloss_dgammas + HW + pix);
      dlb.x = __ldg(dloss_dgammas + 2 * HW + pix);
    }
    if (in1) {
      cont1 = min(__ldg(contrib + pix + 1), len);
      tau.y = __ldg(final_tau + pix + 1);
      dlr.y = __ldg(dloss_dgammas + pix + 1);
      dlg.y = __ldg(dloss_dgammas + HW + pix + 1);
      dlb.y = __ldg(dloss_dgammas + 2 * HW + pix + 1);
    }
    int wmax = max(cont0, cont1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_wmax[warp] = wmax;
    __syncthreads();
    const int bmax = max(max(s_wmax[0], s_wmax[1]), max(s_wmax[2], s_wmax[3]));
    if (bmax <= 0) continue;
    const int nbn = (bmax + BWD2_BATCH - 1) / BWD2_BATCH;  // batches [0, nbn) are needed
    const int32_t *ids = gsid + range.x;
    for (int bi = 0; bi < 2 && bi < nbn; bi++) {  // stage fill = per-thread gather, see raster_fwd2.cu
      const int o = (nbn - 1 - bi) * BWD2_BATCH + tid;
      gather_record(&sbuf[bi][0], recs, o < len ? __ldg(ids + o) : 0, o < len,
                    min(BWD2_BATCH, len - (nbn - 1 - bi) * BWD2_BATCH), &mbar[bi], tid);
    }
#if BWD2_PREFETCH_IDS
    int g_pref = (nbn > 2) ? __ldg(ids + (nbn - 3) * BWD2_BATCH + tid) : 0;  // batch nbn-3 is full
#endif

    const float2 npx = g2(-(float)px, -(float)(px + 1));
    const float fpy = (float)py;
    const float bx0 = (float)rx0, bx1 = (float)(rx0 + 7), by0 = (float)ry0, by1 = (float)(ry0 + 7);
    float2 sdot = g2s(0.f);  // dL/dgamma . gamma_next, per pixel

    for (int bi = 0; bi < nbn; bi++) {
      const int b = nbn - 1 - bi;
      const int s = bi & 1;
      if (s == 0) { mbar_wait(&mbar[0], ph0 & 1); ph0++; } else { mbar_wait(&mbar[1], ph1 & 1); ph1++; }
      const int nrec = min(BWD2_BATCH, len - b * BWD2_BATCH);
      if (b * BWD2_BATCH < wmax) {
        for (int c0 = ((nrec - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
          const int j = c0 + lane;
          bool hit = false;
          if (j < nrec && b * BWD2_BATCH + j < wmax)
            hit = rec_can_touch(sbuf[s][j].q0, sbuf[s][j].q1, bx0, bx1, by0, by1);
          unsigned mask = __ballot_sync(0xffffffffu, hit);
          while (mask) {
            const int k = 31 - __clz(mask);  // back to front
            mask &= ~(1u << k);
            const Rec *r = &sbuf[s][c0 + k];
            const int idx = b * BWD2_BATCH + c0 + k;
            const float4 q0 = r->q0, q1 = r->q1;
            const float2 dx = __fadd2_rn(g2s(q0.x), npx);
            const float dy = q0.y - fpy;
            const float cdy2 = (q1.z * dy) * dy;
            const float2 t = __ffma2_rn(g2s(q1.y), g2s(dy), __fmul2_rn(g2s(q1.x), dx));
            const float2 p = __ffma2_rn(t, dx, g2s(cdy2));
            const float2 gg = g2(ex2_approx(fminf(p.x, 0.0f)), ex2_approx(fminf(p.y, 0.0f)));
            const float2 ag = __fmul2_rn(g2s(q0.w), gg);
            const float ap0 = fminf(ALPHA_CLAMP, ag.x), ap1 = fminf(ALPHA_CLAMP, ag.y);
            const bool a0 = (idx < cont0) && (ap0 >= ALPHA_SKIP);
            const bool a1 = (idx < cont1) && (ap1 >= ALPHA_SKIP);
            if (!__any_sync(0xffffffffu, a0 || a1)) continue;
            const float4 q2 = r->q2;
            // an inactive pixel replays alpha' = 0: tau / (1 - 0) = tau, all nine terms exactly 0
            const float2 e = g2(a0 ? ap0 : 0.0f, a1 ? ap1 : 0.0f);
            const float2 om = __fadd2_rn(g2s(1.0f), g2(-e.x, -e.y));
            tau = __fmul2_rn(tau, g2(a0 ? rcp_approx(om.x) : 1.0f, a1 ? rcp_approx(om.y) : 1.0f));
            const float2 dc = __ffma2_rn(dlr, g2s(q2.x), __ffma2_rn(dlg, g2s(q2.y), __fmul2_rn(dlb, g2s(q2.z))));
            const float2 diff = __fadd2_rn(dc, g2(-sdot.x, -sdot.y));
            sdot = __ffma2_rn(e, diff, sdot);
            const float2 dl_dap = __fmul2_rn(g2(a0 ? tau.x : 0.0f, a1 ? tau.y : 0.0f), diff);
            const float2 wc = __fmul2_rn(e, tau);
            const float2 w = __fmul2_rn(dl_dap, e);
            const float2 wdx = __fmul2_rn(w, dx), wdy = __fmul2_rn(w, g2s(dy));
            const float2 m2 = __fmul2_rn(wdx, dx), m3 = __fmul2_rn(wdx, g2s(dy)), m4 = __fmul2_rn(wdy, g2s(dy));
            const float2 m5 = __fmul2_rn(dl_dap, gg);
            const float2 m6 = __fmul2_rn(wc, dlr), m7 = __fmul2_rn(wc, dlg), m8 = __fmul2_rn(wc, dlb);
            const float v[9] = {wdx.x + wdx.y, wdy.x + wdy.y, m2.x + m2.y, m3.x + m3.y, m4.x + m4.y,
                                m5.x + m5.y, m6.x + m6.y, m7.x + m7.y, m8.x + m8.y};
            const float tot = split_reduce9_v2(v, lane);
            if (slot >= 0) atomicAdd(mom_lane + (size_t)__float_as_int(q0.z) * MOM2, tot);
          }
        }
      }
      __syncthreads();  // every warp is done with stage s
      if (bi + 2 < nbn) {  // batches below the last one are always full
#if BWD2_PREFETCH_IDS
        gather_record(&sbuf[s][0], recs, g_pref, true, BWD2_BATCH, &mbar[s], tid);
        g_pref = (bi + 3 < nbn) ? __ldg(ids + (nbn - 1 - (bi + 3)) * BWD2_BATCH + tid) : 0;
#else
        gather_record(&sbuf[s][0], recs, __ldg(ids + (nbn - 1 - (bi + 2)) * BWD2_BATCH + tid), true, BWD2_BATCH,
                      &mbar[s], tid);
#endif
      }
    }
  }
}

int persistent_grid(int T, int ctas_per_sm);  // raster_fwd2.cu

int launch_draw_bwd2_kernel(int H, int W, const int32_t *ranges, const Rec *recs, const int32_t *gsid,
                            const int32_t *contrib,
                            const float *final_tau, const float *dloss_dgammas, float *moments,
                            int *tile_counter, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int T = gx * gy;
  if (tile_counter != nullptr) GSB_CUDA_TRY(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
  ProfScope ps(K_DRAW_BWD, st);
  k_draw_bwd2<<<tile_counter != nullptr ? persistent_grid(T, 12) : T, 128, 0, st>>>(W, H, gx, T, reinterpret_cast<const int2 *>(ranges), recs,
                                                      gsid, contrib, final_tau, dloss_dgammas, moments, tile_counter);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

// Backward tile rasterizer, variant 3: pixel-parallel replay + record-parallel moment sums,
// transposed through shared memory (no warp reduction per record).
//
// Variant 2 (raster_bwd2.cu) is instruction-issue bound, and ~60 of its 131 instructions per
// (warp, record) are the reduction of the nine per-pixel moment terms over the warp's 64 pixels
// (pair add, 12-shuffle split butterfly with its selects, the RED).  The nine sums are linear in
// two per-pixel weights,
//     w  = dL/dalpha' * alpha'     -> sum w dx, w dy, w dx^2, w dx dy, w dy^2, and
//                                     sum dL/dalpha' g = (sum w) / alpha   (alpha' = alpha g below the
//                                     0.99 clamp; records with alpha > 0.99 take a warp reduction instead)
//     wc = alpha' * tau            -> sum wc dL/dgamma_{r,g,b}
// so the kernel is split into two phases inside each warp:
//   phase A (lane = 2 adjacent pixels of the warp's 8x8 block, as in variant 2): replay the
//     saved (final_tau, contrib) state back to front, and for every record that is active at some
//     pixel write the two weight pairs of the lane into a shared-memory row [slot][2][64 px];
//   phase B (every 8 such records; lane = (slot, quarter of the block)): each lane reads 16
//     pixels of its record's two rows with 128-bit loads and accumulates all nine moments in
//     registers against pixel offsets that are compile-time constants; a two-level split
//     exchange between the four lanes of a slot (8 shuffles per EIGHT records) and three RED
//     instructions per eight records finish the job.
// Per-pixel arithmetic of phase A is the instruction sequence of variant 2 (identical skip
// decisions: alpha' < 0.002, index >= contrib); dx = u_x - px is exact in both formulations
// (integer-valued px), so only the summation order of the moments differs.
// Replaces reference kernel.cu:809-950 (drawB) together with raster_bwd.cu's launcher.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

#ifndef BWD3_BATCH
#define BWD3_BATCH 128
#endif
constexpr int B3_BATCH = BWD3_BATCH;  // records per stage (a multiple of 32, <= 128: one gather per thread)
constexpr int B3_SLOTS = 8;
// slot stride = 2 rows of 64 floats + 16 floats of padding: stride = 16 (mod 32) floats makes the
// phase-B 128-bit reads (lanes = 2 slots x 4 quarter-block parts per 8-lane wavefront) conflict-free
constexpr int B3_SLOT_BYTES = (2 * 64 + 16) * 4;      // 576
constexpr int B3_WARP_W_BYTES = B3_SLOTS * B3_SLOT_BYTES;  // 4608
#ifndef BWD3_STATS
#define BWD3_STATS 0
#endif
#ifndef BWD3_MINBLOCKS
#define BWD3_MINBLOCKS 6  // A/B at config 2 (benchmarks/ab_variants.py): 5 CTAs/SM 0.609 ms, 6 (79 regs) 0.581
#endif

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {  // keep an address in a register
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

#if BWD3_STATS  // variant builds only (benchmarks/ab_variants.py): workload counters of the last launches
__device__ unsigned long long g_b3_stats[8];  // tested, survivors, active survivors, flushes, flushed slots
#define B3_COUNT(i, n) do { if (lane == 0) atomicAdd(&g_b3_stats[i], (unsigned long long)(n)); } while (0)
extern "C" int gsb_debug_bwd3_stats(unsigned long long *out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, g_b3_stats, sizeof(g_b3_stats));
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0};
    e = cudaMemcpyToSymbol(g_b3_stats, z, sizeof(z));
  }
  return (int)e;
}
#else
#define B3_COUNT(i, n) do { } while (0)
#endif

__device__ __forceinline__ float2 p2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 p2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 lo(const float4 &v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi(const float4 &v) { return make_float2(v.z, v.w); }

// Phase B: moment sums of up to 8 records over the warp's 8x8 block.
//   lane = slot * 4 + part; part covers x in [4 (part & 1), +4) of rows (part >> 1) + {0, 2, 4, 6}
//   (pixel chunk c = part + 4 j of the 16 four-pixel chunks -> 16-byte column c of every row).
__device__ __forceinline__ void b3_flush(uint32_t w_addr, uint32_t info_addr, uint32_t dl_addr, int nslots,
                                         float fbx0, float fby0, float *__restrict__ moments, int lane) {
  const int slot = lane >> 2, part = lane & 3;
  // (u_x, u_y, gaussian id, alpha) of the slot's record.  (A/B: keeping these in registers of the
  // slot's lanes instead -- 5 more ALU instructions per record, no 16-byte store -- was 2 % slower.)
  const float4 inf = lds128(info_addr + slot * 16);
  const float U = inf.x - (fbx0 + (float)((part & 1) * 4));
  const float V = inf.y - (fby0 + (float)(part >> 1));
  const float2 dx01 = p2(U, U - 1.0f), dx23 = p2(U - 2.0f, U - 3.0f);
  const uint32_t a = w_addr + slot * B3_SLOT_BYTES + part * 16;
  const uint32_t d = dl_addr + part * 16;
  float2 m0 = p2s(0.f), m1 = p2s(0.f), m2 = p2s(0.f), m3 = p2s(0.f), m4 = p2s(0.f), m5 = p2s(0.f),
         m6 = p2s(0.f), m7 = p2s(0.f), m8 = p2s(0.f);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float dy = V - (float)(2 * j);
    const float dy2 = dy * dy;
    const float4 w = lds128(a + j * 64);
    const float4 wc = lds128(a + 256 + j * 64);
    const float4 dr = lds128(d + j * 64);
    const float4 dg = lds128(d + 256 + j * 64);
    const float4 db = lds128(d + 512 + j * 64);
    const float2 t01 = __fmul2_rn(lo(w), dx01), t23 = __fmul2_rn(hi(w), dx23);
    const float2 r1 = __fadd2_rn(t01, t23), r0 = __fadd2_rn(lo(w), hi(w));
    m0 = __fadd2_rn(m0, r1);
    m2 = __ffma2_rn(t01, dx01, m2);
    m2 = __ffma2_rn(t23, dx23, m2);
    m1 = __ffma2_rn(p2s(dy), r0, m1);
    m3 = __ffma2_rn(p2s(dy), r1, m3);
    m4 = __ffma2_rn(p2s(dy2), r0, m4);
    m5 = __fadd2_rn(m5, r0);
    m6 = __ffma2_rn(lo(wc), lo(dr), m6);
    m6 = __ffma2_rn(hi(wc), hi(dr), m6);
    m7 = __ffma2_rn(lo(wc), lo(dg), m7);
    m7 = __ffma2_rn(hi(wc), hi(dg), m7);
    m8 = __ffma2_rn(lo(wc), lo(db), m8);
    m8 = __ffma2_rn(hi(wc), hi(db), m8);
  }
  // sum dL/dalpha' g = (sum w) / alpha; records above the clamp were reduced in phase A
  const float inv_alpha = inf.w <= ALPHA_CLAMP ? __fdividef(1.0f, inf.w) : 0.0f;
  const float v[9] = {m0.x + m0.y, m1.x + m1.y, m2.x + m2.y, m3.x + m3.y, m4.x + m4.y,
                      (m5.x + m5.y) * inv_alpha, m6.x + m6.y, m7.x + m7.y, m8.x + m8.y};
  // split exchange between the four parts of a slot: after xor 1 the even parts hold the sums of
  // v0..v4 and the odd ones those of v5..v8; after xor 2: part 0 -> v0 v1 v2, part 2 -> v3 v4,
  // part 1 -> v5 v6 v7, part 3 -> v8.
  const unsigned F = 0xffffffffu;
  const bool u1 = part & 1, u2 = part & 2;
  float s[5], b[3];
#pragma unroll
  for (int i = 0; i < 4; i++) s[i] = (u1 ? v[i + 5] : v[i]) + __shfl_xor_sync(F, u1 ? v[i] : v[i + 5], 1);
  s[4] = (u1 ? 0.f : v[4]) + __shfl_xor_sync(F, u1 ? v[4] : 0.f, 1);
  b[0] = (u2 ? s[3] : s[0]) + __shfl_xor_sync(F, u2 ? s[0] : s[3], 2);
  b[1] = (u2 ? s[4] : s[1]) + __shfl_xor_sync(F, u2 ? s[1] : s[4], 2);
  b[2] = (u2 ? 0.f : s[2]) + __shfl_xor_sync(F, u2 ? s[2] : 0.f, 2);
  if (slot < nslots) {
    const int first = u1 ? (u2 ? 8 : 5) : (u2 ? 3 : 0);
    float *row = moments + (size_t)__float_as_int(inf.z) * 9 + first;
    atomicAdd(row, b[0]);
    if (part != 3) atomicAdd(row + 1, b[1]);
    if (!u2) atomicAdd(row + 2, b[2]);
  }
}

__global__ void __launch_bounds__(128, BWD3_MINBLOCKS) k_draw_bwd3(
    int W, int H, int gx, int T, const int2 *__restrict__ ranges, const Rec *__restrict__ recs,
    const int32_t *__restrict__ gsid, const int32_t *__restrict__ contrib, const float *__restrict__ final_tau,
    const float *__restrict__ dloss_dgammas, float *__restrict__ moments, int *__restrict__ tile_counter) {
  __shared__ Rec sbuf[2][B3_BATCH];
  __shared__ __align__(16) unsigned char s_w[4 * B3_WARP_W_BYTES];  // weight rows, per warp
  __shared__ __align__(16) float s_dl[4][3][64];                    // dL/dgamma of the block's pixels
  __shared__ __align__(16) float4 s_info[4][B3_SLOTS];              // (u_x, u_y, id, alpha) of each slot
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ int s_wmax[4];
  __shared__ int s_tile[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t HW = (size_t)H * W;
  if (tid == 0) {
    mbar_init(&mbar[0], GATHER_ARRIVALS);
    mbar_init(&mbar[1], GATHER_ARRIVALS);
    fence_mbar_init();
  }
  uint32_t ph0 = 0, ph1 = 0;
  const uint32_t sbuf_addr = opaque_u32(smem_u32(&sbuf[0][0]));
  const uint32_t w_addr = opaque_u32(smem_u32(s_w + warp * B3_WARP_W_BYTES));
  const uint32_t dl_addr = opaque_u32(smem_u32(&s_dl[warp][0][0]));
  const uint32_t info_addr = opaque_u32(smem_u32(&s_info[warp][0]));

  for (int it = 0;; it++) {
    int tile;
    if (tile_counter != nullptr) {  // persistent grid (sparse frames), see raster_fwd2.cu
      if (tid == 0) s_tile[it & 1] = atomicAdd(tile_counter, 1);
      __syncthreads();
      tile = s_tile[it & 1];
    } else {
      if (it > 0) break;
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
    float2 tau = p2s(0.f), dlr = p2s(0.f), dlg = p2s(0.f), dlb = p2s(0.f);
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
    // the block's dL/dgamma table for phase B (pixel p = 8 row + x = 2 lane + {0, 1})
    sts64(dl_addr + lane * 8, dlr);
    sts64(dl_addr + 256 + lane * 8, dlg);
    sts64(dl_addr + 512 + lane * 8, dlb);
    __syncthreads();
    const int bmax = max(max(s_wmax[0], s_wmax[1]), max(s_wmax[2], s_wmax[3]));
    if (bmax <= 0) continue;
    const int nbn = (bmax + B3_BATCH - 1) / B3_BATCH;
    const int32_t *ids = gsid + range.x;
    for (int bi = 0; bi < 2 && bi < nbn; bi++) {
      const int o = (nbn - 1 - bi) * B3_BATCH + tid;
      const bool v = tid < B3_BATCH && o < len;
      gather_record(&sbuf[bi][0], recs, v ? __ldg(ids + o) : 0, v,
                    min(B3_BATCH, len - (nbn - 1 - bi) * B3_BATCH), &mbar[bi], tid);
    }

    const float2 npx = p2(-(float)px, -(float)(px + 1));
    const float fpy = (float)py;
    const float bx0 = (float)rx0, bx1 = (float)(rx0 + 7), by0 = (float)ry0, by1 = (float)(ry0 + 7);
    float2 sdot = p2s(0.f);
    int nslots = 0;

    for (int bi = 0; bi < nbn; bi++) {
      const int b = nbn - 1 - bi;
      const int s = bi & 1;
      if (s == 0) { mbar_wait(&mbar[0], ph0 & 1); ph0++; } else { mbar_wait(&mbar[1], ph1 & 1); ph1++; }
      const int nrec = min(B3_BATCH, len - b * B3_BATCH);
      if (b * B3_BATCH < wmax) {
        const uint32_t stage_addr = sbuf_addr + s * (B3_BATCH * 48);
        for (int c0 = ((nrec - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
          const int j = c0 + lane;
          bool hit = false;
          if (j < nrec && b * B3_BATCH + j < wmax)
            hit = rec_can_touch(sbuf[s][j].q0, sbuf[s][j].q1, bx0, bx1, by0, by1);
          unsigned mask = __ballot_sync(0xffffffffu, hit);
          B3_COUNT(0, min(32, nrec - c0));
          B3_COUNT(1, __popc(mask));
          const uint32_t chunk_addr = stage_addr + c0 * 48;
          const int chunk_idx = b * B3_BATCH + c0;
          while (mask) {
            const int k = 31 - __clz(mask);  // back to front
            mask ^= (1u << k);
            const uint32_t ra = chunk_addr + k * 48;
            const int idx = chunk_idx + k;
            const float2 u = lds64(ra);
            const float4 q1 = lds128(ra + 16);
            const float2 dx = __fadd2_rn(p2s(u.x), npx);
            const float dy = u.y - fpy;
            const float cdy2 = (q1.z * dy) * dy;
            const float2 t = __ffma2_rn(p2s(q1.y), p2s(dy), __fmul2_rn(p2s(q1.x), dx));
            const float2 p = __ffma2_rn(t, dx, p2s(cdy2));
            const float2 gg = p2(ex2_approx(fminf(p.x, 0.0f)), ex2_approx(fminf(p.y, 0.0f)));
            const float2 ag = __fmul2_rn(p2s(q1.w), gg);
            const float ap0 = fminf(ALPHA_CLAMP, ag.x), ap1 = fminf(ALPHA_CLAMP, ag.y);
            const bool a0 = (idx < cont0) && (ap0 >= ALPHA_SKIP);
            const bool a1 = (idx < cont1) && (ap1 >= ALPHA_SKIP);
            // (99.9 % of the records that pass the rectangle test are active at some pixel of
            // the block: no early-out vote -- an all-inactive record just writes a row of zeros)
            B3_COUNT(2, 1);
            const float4 q2 = lds128(ra + 32);
            // an inactive pixel replays alpha' = 0: tau / (1 - 0) = tau, all three weights exactly 0
            const float2 e = p2(a0 ? ap0 : 0.0f, a1 ? ap1 : 0.0f);
            const float2 om = __fadd2_rn(p2s(1.0f), p2(-e.x, -e.y));
            tau = __fmul2_rn(tau, p2(a0 ? rcp_approx(om.x) : 1.0f, a1 ? rcp_approx(om.y) : 1.0f));
            const float2 dc = __ffma2_rn(dlr, p2s(q2.x), __ffma2_rn(dlg, p2s(q2.y), __fmul2_rn(dlb, p2s(q2.z))));
            const float2 diff = __fadd2_rn(dc, p2(-sdot.x, -sdot.y));
            sdot = __ffma2_rn(e, diff, sdot);
            const float2 dl_dap = __fmul2_rn(p2(a0 ? tau.x : 0.0f, a1 ? tau.y : 0.0f), diff);
            const uint32_t wa = w_addr + nslots * B3_SLOT_BYTES + lane * 8;
            sts64(wa, __fmul2_rn(dl_dap, e));    // w
            sts64(wa + 256, __fmul2_rn(e, tau));  // wc
            sts128(info_addr + nslots * 16, u.x, u.y, q2.w, q1.w);  // (same value from every lane)
            if (q1.w > ALPHA_CLAMP) {  // opacity above the clamp (rare): sum dL/dalpha' g by shuffles
              float v5 = fmaf(dl_dap.x, gg.x, dl_dap.y * gg.y);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v5 += __shfl_xor_sync(0xffffffffu, v5, o);
              if (lane == 0) atomicAdd(moments + (size_t)__float_as_int(q2.w) * 9 + 5, v5);
            }
            nslots++;
            if (nslots == B3_SLOTS) {
              __syncwarp();
              B3_COUNT(3, 1);
              B3_COUNT(4, B3_SLOTS);
              b3_flush(w_addr, info_addr, dl_addr, B3_SLOTS, bx0, by0, moments, lane);
              __syncwarp();
              nslots = 0;
            }
          }
        }
      }
      __syncthreads();  // every warp is done with stage s
      if (bi + 2 < nbn) {
        const bool v = tid < B3_BATCH;  // batches below the last one are always full
        gather_record(&sbuf[s][0], recs, v ? __ldg(ids + (nbn - 1 - (bi + 2)) * B3_BATCH + tid) : 0, v, B3_BATCH,
                      &mbar[s], tid);
      }
    }
    if (nslots > 0) {  // partial group at the end of the tile
      __syncwarp();
      B3_COUNT(3, 1);
      B3_COUNT(4, nslots);
      b3_flush(w_addr, info_addr, dl_addr, nslots, bx0, by0, moments, lane);
      __syncwarp();
    }
  }
}

int persistent_grid(int T, int ctas_per_sm);  // raster_fwd2.cu

int launch_draw_bwd3_kernel(int H, int W, const int32_t *ranges, const Rec *recs, const int32_t *gsid,
                            const int32_t *contrib, const float *final_tau, const float *dloss_dgammas,
                            float *moments, int *tile_counter, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int T = gx * gy;
  if (tile_counter != nullptr) GSB_CUDA_TRY(cudaMemsetAsync(tile_counter, 0, sizeof(int), st));
  ProfScope ps(K_DRAW_BWD, st);
  k_draw_bwd3<<<tile_counter != nullptr ? persistent_grid(T, BWD3_MINBLOCKS) : T, 128, 0, st>>>(
      W, H, gx, T, reinterpret_cast<const int2 *>(ranges), recs, gsid, contrib, final_tau, dloss_dgammas, moments,
      tile_counter);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

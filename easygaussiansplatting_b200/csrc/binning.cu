// Tile binning: per-Gaussian tile rectangles + exclusive offsets (one fused kernel with a
// decoupled look-back scan), (tile | depth-mm) keys, a hand-written LSD radix sort (per pass: a
// column scan of the per-tile digit counts + one rank-and-scatter kernel that also counts the
// next pass's digits), per-tile ranges and the packed per-Gaussian record stream.
//
// Replaces getRects / thrust::inclusive_scan / createKeys / thrust::sort_by_key / getRanges
// (reference kernel.cu:46-150, gausplat.cu:50-91).  No library kernels: round 1 used CUB's
// DeviceScan / DeviceRadixSort, whose tile sizes gave 1.05-wave launches at this problem size
// (2.5 M pairs: 29 us per pass, 16 us for a 1 M-element scan).  Everything is stream-ordered in
// caller-provided workspace and every kernel takes the patch count P from DEVICE memory, so the
// same kernels serve the synchronous operator surface (P read back once, like the reference,
// gausplat.cu:67) and the capacity-based path without a host round trip (api.cu).
//
// Sort width.  The reference sorts all 64 bits of (tile << 32 | (uint32)(depth*1000)).  Only
// ceil(log2(tiles)) + ceil(log2(max depth key + 1)) of them can differ; when the two fields
// fit 32 bits the keys are packed as (tile << dbits | depth key) in 32-bit words -- the same
// (tile, depth-mm, id) order with half the key traffic -- and only the digits that can be
// non-zero are sorted (9-bit digits when that saves a pass: 27 bits at 1080p = 3 passes).
// Every pass is stable, so ties keep ascending Gaussian id exactly like thrust::sort_by_key.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int bits_for(uint64_t n_values) {  // bits needed to represent 0 .. n_values-1
  int b = 0;
  while (((uint64_t)1 << b) < n_values) b++;
  return b;
}
static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    sms = 148;
  return sms;
}

// ---- relaxed / acquire-release device-scope accesses for the look-back descriptors
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// =================================================================== phase 1: rects + scan
// One CTA = 1024 consecutive Gaussians (4 per thread, strided so every access is coalesced).
// Rectangle arithmetic: kernel.cu:105-110; DIV_ROUND_UP(X,16) on floats is ((X) + 16 - 1) / 16
// (common.cuh:15).  Explicit _rn intrinsics: these decide the patch list, no contraction.
// The inclusive prefix sum of the patch counts uses the single-pass decoupled look-back of
// Merrill & Garland: a CTA publishes its aggregate, then its inclusive prefix, in one 64-bit
// word (2 flag bits | value), and a warp walks back over its predecessors 32 at a time.
constexpr int RS_CTA = 256, RS_ITEMS = 4, RS_SPAN = RS_CTA * RS_ITEMS;
constexpr unsigned long long DESC_AGG = 1ull << 62, DESC_INC = 2ull << 62, DESC_VAL = (1ull << 62) - 1;

__global__ void __launch_bounds__(RS_CTA) k_rects_scan(int N, const float2 *__restrict__ us,
                                                       int2 *__restrict__ areas, float *__restrict__ depths,
                                                       int gx, int gy, uint2 *__restrict__ rects,
                                                       uint32_t *__restrict__ incl, uint32_t *__restrict__ total,
                                                       unsigned long long *__restrict__ desc,
                                                       uint32_t *__restrict__ claim) {
  __shared__ uint32_t s_warp[RS_CTA / 32];
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_max[RS_CTA / 32];
  __shared__ int s_blk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_blk = (int)atomicAdd(claim, 1u);  // claim order = data order: look-back cannot deadlock
  __syncthreads();
  const int blk = s_blk;
  const int nblk = (N + RS_SPAN - 1) / RS_SPAN;
  uint32_t cnt[RS_ITEMS], dkmax = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const int i = blk * RS_SPAN + k * RS_CTA + tid;
    uint32_t n = 0;
    if (i < N) {
      uint2 rect = make_uint2(0u, 0u);
      const float d = depths[i];
      if (!(d < MIN_DEPTH)) {
        const float2 u = __ldg(us + i);
        const int2 ar = areas[i];
        const float xs = (float)ar.x, ys = (float)ar.y;
        const int x0 = min(gx, max(0, (int)__fdiv_rn(__fsub_rn(u.x, xs), 16.0f)));
        const int y0 = min(gy, max(0, (int)__fdiv_rn(__fsub_rn(u.y, ys), 16.0f)));
        const int x1 = min(gx, max(0, (int)__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(u.x, xs), 16.0f), 1.0f), 16.0f)));
        const int y1 = min(gy, max(0, (int)__fdiv_rn(__fsub_rn(__fadd_rn(__fadd_rn(u.y, ys), 16.0f), 1.0f), 16.0f)));
        n = (uint32_t)(y1 - y0) * (uint32_t)(x1 - x0);
        if (n == 0) {  // kernel.cu:114-119: in-place cull
          depths[i] = BAD_MARKER;
          areas[i] = make_int2(0, 0);
        } else {
          rect = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)y0 | ((uint32_t)y1 << 16));
          dkmax = max(dkmax, __float2uint_rz(__fmul_rn(d, 1000.0f)));  // kernel.cu:73
        }
      }
      rects[i] = rect;  // rect packed as (x0 | x1 << 16, y0 | y1 << 16); grids up to 65535 x 65535
    }
    cnt[k] = n;
  }
  // CTA-local inclusive scan in data order (sub-tile k holds elements k*256 .. k*256+255)
  uint32_t loc[RS_ITEMS], run = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    uint32_t v = cnt[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    uint32_t wbase = 0, sub = 0;
#pragma unroll
    for (int w = 0; w < RS_CTA / 32; w++) {
      const uint32_t t = s_warp[w];
      if (w < warp) wbase += t;
      sub += t;
    }
    loc[k] = run + wbase + v;
    run += sub;
    __syncthreads();
  }
  // look-back (warp 0): exclusive prefix of this CTA over all earlier CTAs
  if (warp == 0) {
    unsigned long long excl = 0;
    if (blk > 0) {
      if (lane == 0) st_relaxed_u64(desc + blk, DESC_AGG | run);
      int p = blk - 1;  // lane l inspects predecessor p - l
      for (;;) {
        const int q = p - lane;
        unsigned long long v = DESC_INC;  // virtual predecessor before block 0: inclusive 0
        if (q >= 0) {
          do { v = ld_relaxed_u64(desc + q); } while ((v >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;  // nearest predecessor with an inclusive value
        unsigned long long part = (lane <= first_inc) ? (v & DESC_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        excl += part;
        if (inc_mask) break;
        p -= 32;
      }
    }
    if (lane == 0) {
      st_relaxed_u64(desc + blk, DESC_INC | (excl + run));
      s_prefix = excl;
      if (blk == nblk - 1) total[0] = (uint32_t)min(excl + run, 0xffffffffull);  // P (saturating)
    }
  }
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, dkmax);
  if (lane == 0) s_max[warp] = wmax;
  __syncthreads();
  const uint32_t prefix = (uint32_t)s_prefix;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const int i = blk * RS_SPAN + k * RS_CTA + tid;
    if (i < N) incl[i] = prefix + loc[k];
  }
  if (tid == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < RS_CTA / 32; w++) m = max(m, s_max[w]);
    if (m > 0) atomicMax(total + 1, m);  // largest depth key of a binned Gaussian
    if (s_prefix + run > 0xffffffffull) atomicOr(total + 2, 1u);  // P does not fit 32 bits
  }
}

// =================================================================== phase 2: keys
// Sort plan: which bit ranges of the key are sorted, in which order (LSD).
constexpr int MAX_PASSES = 8;
struct PassPlan {
  int n, bits;  // passes, digit width (8 or 9)
  int shift[MAX_PASSES];
};
struct KeyPlan {
  bool narrow;  // 32-bit keys
  int shift;    // tile field starts at this bit
  PassPlan pp;
};
static KeyPlan key_plan(int H, int W, uint32_t depth_key_max) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tbits = bits_for((uint64_t)gx * (uint64_t)gy);
  const int dbits = depth_key_max == 0xffffffffu ? 32 : bits_for((uint64_t)depth_key_max + 1);
  KeyPlan p{};
  p.narrow = dbits < 32 && tbits + dbits <= 32;  // (dbits == 32 would need a shift by the word width)
  p.shift = p.narrow ? dbits : 32;
  PassPlan &pp = p.pp;
  if (p.narrow) {
    const int end = dbits + tbits;
    pp.bits = (end > 0 && (end + 8) / 9 < (end + 7) / 8) ? 9 : 8;
    for (int s = 0; s < end; s += pp.bits) pp.shift[pp.n++] = s;
  } else {  // 64-bit layout: depth bits [0, dbits), tile bits [32, 32 + tbits); the gap is all zero
    pp.bits = 8;
    for (int s = 0; s < dbits; s += 8) pp.shift[pp.n++] = s;
    for (int s = 0; s < tbits; s += 8) pp.shift[pp.n++] = 32 + s;
  }
  return p;
}

// One thread per Gaussian (grid-stride), serial over its rectangle (2.5 patches/Gaussian on the
// benchmark scenes); rectangles of more than 32 tiles are emitted by the whole warp.  64-bit:
// tile << 32 | dk (kernel.cu:71-74); 32-bit: tile << dbits | dk.  Also counts, for every sort
// pass, the digit histogram of all keys (shared-memory counters, flushed once per CTA).
// *status |= 2 when a Gaussian's patches do not fit the capacity, |= 4 when a depth key needs more
// bits than the plan has (both only possible on the capacity-based path, gsb_splat_forward).
constexpr int RX_THREADS = 256, RX_WARPS = RX_THREADS / 32, RX_IPT = 16, RX_TILE = RX_THREADS * RX_IPT;
template <typename KeyT>
__global__ void __launch_bounds__(256) k_keys(int N, const float *__restrict__ depths,
                                              const uint32_t *__restrict__ incl, const uint2 *__restrict__ rects,
                                              int gx, int shift, const __grid_constant__ PassPlan pp,
                                              uint32_t capacity,
                                              KeyT *__restrict__ keys, int32_t *__restrict__ vals,
                                              uint32_t *__restrict__ hist, uint32_t *__restrict__ tile_hist,
                                              uint32_t *__restrict__ status) {
  extern __shared__ uint32_t s_hist[];  // [pp.n][1 << pp.bits]
  const int bins = 1 << pp.bits;
  for (int i = threadIdx.x; i < pp.n * bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t dmask = (uint32_t)bins - 1u;
  auto emit = [&](uint32_t pos, KeyT key, int id) {
    keys[pos] = key;
    vals[pos] = id;
    for (int p = 0; p < pp.n; p++) atomicAdd(&s_hist[p * bins + ((uint32_t)(key >> pp.shift[p]) & dmask)], 1u);
    // digit counts of the first pass per sort tile (the passes count those of their successor)
    if (pp.n > 0) atomicAdd(tile_hist + (size_t)(pos / RX_TILE) * bins + ((uint32_t)(key >> pp.shift[0]) & dmask), 1u);
  };
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x; base < N; base += stride) {  // (warp-uniform trip count)
    const int i = base + threadIdx.x;
    uint32_t off = 0, x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    KeyT dk = 0;
    if (i < N) {
      const float d = __ldg(depths + i);
      if (!(d < MIN_DEPTH)) {
        off = (i == 0) ? 0u : __ldg(incl + i - 1);
        const uint2 r = __ldg(rects + i);
        x0 = r.x & 0xffffu; x1 = r.x >> 16; y0 = r.y & 0xffffu; y1 = r.y >> 16;
        dk = (KeyT)__float2uint_rz(__fmul_rn(d, 1000.0f));
        if (sizeof(KeyT) == 4 && (dk >> shift) != 0) {  // deeper than the planned key width (capacity path)
          atomicOr(status, 4u);
          dk &= ((KeyT)1 << shift) - 1;
        }
      }
    }
    uint32_t w = x1 - x0, n = w * (y1 - y0);
    if (n > 0 && (off > capacity || n > capacity - off)) {  // does not fit: drop, flag
      atomicOr(status, 2u);
      n = 0; w = 0; y1 = y0;
    }
    if (n <= 32) {
      for (uint32_t y = y0; y < y1; y++)
        for (uint32_t x = x0; x < x0 + w; x++) emit(off++, ((KeyT)(y * (uint32_t)gx + x) << shift) | dk, i);
    }
    unsigned big = __ballot_sync(0xffffffffu, n > 32);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t bn = __shfl_sync(0xffffffffu, n, src), bw = __shfl_sync(0xffffffffu, w, src);
      const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
      const uint32_t boff = __shfl_sync(0xffffffffu, off, src);
      const KeyT bdk = (KeyT)__shfl_sync(0xffffffffu, (unsigned long long)dk, src);
      const int bi = __shfl_sync(0xffffffffu, i, src);
      for (uint32_t t = lane; t < bn; t += 32) {  // row-major inside the rectangle == reference order
        const uint32_t y = by0 + t / bw, x = bx0 + t % bw;
        emit(boff + t, ((KeyT)(y * (uint32_t)gx + x) << shift) | bdk, bi);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < pp.n * bins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + i, s_hist[i]);
}

// =================================================================== phase 2: radix sort
// LSD radix sort, one pass = two kernels, no inter-CTA waiting:
//   k_colscan   turns the per-tile digit counts C[tile][digit] of the pass into scatter bases
//               B[tile][digit] = (keys with a smaller digit) + (keys with this digit in earlier tiles);
//   k_radix_pass  a CTA takes a tile of 4096 pairs, ranks its keys (stable: warps own contiguous
//               512-element spans, the lanes of a round are ordered by ballots over the digit
//               bits), scatters them to B + rank, and counts the NEXT pass's digit of every key
//               into that pass's C at the key's new tile (one RED per key).
// The first pass's C comes from k_keys.  (A single-kernel pass with decoupled look-back per digit
// was tried first: with every CTA of a 2.5 M-pair pass starting in lockstep the look-back walked
// ~40 tiles back on average -- 50 us per pass, slower than the library it replaced.)
template <int BITS>
__global__ void __launch_bounds__(1024) k_colscan(const uint32_t *__restrict__ P_dev, uint32_t capacity,
                                                  const uint32_t *__restrict__ hist, uint32_t *__restrict__ C) {
  constexpr int BINS = 1 << BITS, WARPS = 32;
  __shared__ uint32_t s_g[BINS];
  __shared__ uint32_t s_part[WARPS][32];
  __shared__ uint32_t s_scan[WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (P_dev[2] != 0) return;  // [P, max key, flags]: capacity overflow, the frame is abandoned (see k_ranges)
  const uint32_t P = min(*P_dev, capacity);
  const int T = (int)((P + RX_TILE - 1) / RX_TILE);
  {  // exclusive scan of the global digit histogram (BINS <= 1024 = one value per thread)
    const uint32_t v = tid < BINS ? hist[tid] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    uint32_t run = inc - v;
    for (int w = 0; w < warp; w++) run += s_scan[w];
    if (tid < BINS) s_g[tid] = run;
  }
  // this CTA: digits [32 blockIdx.x, +32); warp w: tiles [w R, (w + 1) R); lane = digit
  const int d = blockIdx.x * 32 + lane;
  const int R = (T + WARPS - 1) / WARPS;
  const int t0 = min(T, warp * R), t1 = min(T, t0 + R);
  uint32_t *col = C + d;
  uint32_t sum = 0;
#pragma unroll 8
  for (int t = t0; t < t1; t++) sum += col[(size_t)t * BINS];
  s_part[warp][lane] = sum;
  __syncthreads();
  uint32_t run = s_g[d];
  for (int w = 0; w < warp; w++) run += s_part[w][lane];
#pragma unroll 8
  for (int t = t0; t < t1; t++) {
    const uint32_t c = col[(size_t)t * BINS];
    col[(size_t)t * BINS] = run;
    run += c;
  }
}

#ifndef RX_UNROLL_OUT
#define RX_UNROLL_OUT 0
#endif
#ifndef RX_MINBLOCKS
#define RX_MINBLOCKS 3
#endif
template <typename KeyT, int BITS>
__global__ void __launch_bounds__(RX_THREADS, sizeof(KeyT) == 8 ? 2 : RX_MINBLOCKS) k_radix_pass(const KeyT *__restrict__ keys_in,
                                                              const int32_t *__restrict__ vals_in,
                                                              KeyT *__restrict__ keys_out,
                                                              int32_t *__restrict__ vals_out,
                                                              const uint32_t *__restrict__ P_dev,
                                                              uint32_t capacity, int shift, int next_shift,
                                                              const uint32_t *__restrict__ base,
                                                              uint32_t *__restrict__ next_C) {
  constexpr int BINS = 1 << BITS, PER = BINS / RX_THREADS;
  __shared__ uint32_t s_whist[RX_WARPS][BINS];  // per-warp digit counts -> tile-local offsets
  __shared__ uint32_t s_gofs[BINS];             // global position of a digit's run minus its local start
  __shared__ uint32_t s_scan[RX_WARPS];
  extern __shared__ __align__(16) unsigned char s_dyn[];  // the tile in sorted order: coalesced write-out
  KeyT *const s_keys = reinterpret_cast<KeyT *>(s_dyn);
  int32_t *const s_vals = reinterpret_cast<int32_t *>(s_dyn + RX_TILE * sizeof(KeyT));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Capacity overflow: k_keys left holes with stale keys that no histogram counted, so scatter
  // positions could run past the buffers.  Nothing is sorted; k_ranges leaves every range empty.
  if (P_dev[2] != 0) return;
  const uint32_t P = min(*P_dev, capacity);
  const int ntiles = (int)((P + RX_TILE - 1) / RX_TILE);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();  // previous tile's shared state is no longer read
    for (int i = tid; i < RX_WARPS * BINS; i += RX_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t first = (uint32_t)tile * RX_TILE + warp * (32 * RX_IPT) + lane;
    const uint32_t tile_n = min((uint32_t)RX_TILE, P - (uint32_t)tile * RX_TILE);
    KeyT key[RX_IPT];
    uint32_t rank[RX_IPT];
#pragma unroll
    for (int r = 0; r < RX_IPT; r++) {
      const uint32_t idx = first + r * 32;
      key[r] = idx < P ? keys_in[idx] : (KeyT)0;
    }
#pragma unroll
    for (int r = 0; r < RX_IPT; r++) {
      const bool valid = first + r * 32 < P;
      const uint32_t d = (uint32_t)(key[r] >> shift) & (BINS - 1);
      // lanes of this round with the same digit (ballot per digit bit: __match_any_sync is a
      // data-dependent loop in hardware)
      unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int b = 0; b < BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = s_whist[warp][d];
        s_whist[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader < 0 ? 0 : leader);
      rank[r] = old + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
    }
    __syncthreads();
    {  // per digit (thread t owns digits [PER t, PER t + PER)): offsets of the warps, local start
      uint32_t cnt[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int d = tid * PER + k;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; w++) {
          const uint32_t t = s_whist[w][d];
          s_whist[w][d] = run;
          run += t;
        }
        cnt[k] = run;
        sum += run;
      }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_scan[warp] = inc;
      __syncthreads();
      uint32_t start = inc - sum;  // tile-local start of this thread's first digit
      for (int w = 0; w < warp; w++) start += s_scan[w];
#pragma unroll
      for (int k = 0; k < PER; k++) {
        const int d = tid * PER + k;
#pragma unroll
        for (int w = 0; w < RX_WARPS; w++) s_whist[w][d] += start;
        s_gofs[d] = base[(size_t)tile * BINS + d] - start;
        start += cnt[k];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RX_IPT; r++) {
      if (first + r * 32 < P) {
        const uint32_t lp = s_whist[warp][(uint32_t)(key[r] >> shift) & (BINS - 1)] + rank[r];
        s_keys[lp] = key[r];
        s_vals[lp] = vals_in[first + r * 32];  // (read here, not held in registers through the ranking)
      }
    }
    __syncthreads();
#if !RX_UNROLL_OUT
    for (uint32_t j = tid; j < tile_n; j += RX_THREADS) {  // sorted order: runs of equal digits are contiguous
      const KeyT k = s_keys[j];
      const uint32_t pos = s_gofs[(uint32_t)(k >> shift) & (BINS - 1)] + j;
      keys_out[pos] = k;
      vals_out[pos] = s_vals[j];
      if (next_C != nullptr)
        atomicAdd(next_C + (size_t)(pos / RX_TILE) * BINS + ((uint32_t)(k >> next_shift) & (BINS - 1)), 1u);
    }
#else
    {  // sorted order: runs of equal digits are contiguous.  Unrolled: 16 independent LDS -> STG chains
      KeyT ko[RX_IPT];
      int32_t vo[RX_IPT];
#pragma unroll
      for (int r = 0; r < RX_IPT; r++) {
        const uint32_t j = tid + r * RX_THREADS;
        ko[r] = j < tile_n ? s_keys[j] : (KeyT)0;
        vo[r] = j < tile_n ? s_vals[j] : 0;
      }
#pragma unroll
      for (int r = 0; r < RX_IPT; r++) {
        const uint32_t j = tid + r * RX_THREADS;
        if (j < tile_n) {
          const uint32_t pos = s_gofs[(uint32_t)(ko[r] >> shift) & (BINS - 1)] + j;
          keys_out[pos] = ko[r];
          vals_out[pos] = vo[r];
          if (next_C != nullptr)
            atomicAdd(next_C + (size_t)(pos / RX_TILE) * BINS + ((uint32_t)(ko[r] >> next_shift) & (BINS - 1)), 1u);
        }
      }
    }
#endif
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(256) k_ranges(const uint32_t *__restrict__ P_dev, uint32_t capacity,
                                                const uint32_t *__restrict__ flags,
                                                const KeyT *__restrict__ keys, int shift, uint32_t T,
                                                int2 *__restrict__ ranges) {
  // After a capacity overflow (k_keys dropped the Gaussians that did not fit) the key / id arrays
  // have holes with stale contents and were not sorted: every range stays empty (the memset before
  // the sort), so that the rasterizer enqueued behind this kernel touches no patch and gathers no
  // record through a stale id.  The frame is discarded and redone by the caller anyway.
  if (*flags != 0) return;
  const uint32_t P = min(*P_dev, capacity);
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const uint32_t t = (uint32_t)(keys[p] >> shift);
    if (t >= T) continue;  // (only after a capacity overflow, when the result is discarded anyway)
    if (p == 0 || (uint32_t)(keys[p - 1] >> shift) != t) ranges[t].x = (int)p;
    if (p == P - 1 || (uint32_t)(keys[p + 1] >> shift) != t) ranges[t].y = (int)(p + 1);
  }
}

// Gathers the four per-Gaussian attribute arrays (all L2 resident: 36 B/Gaussian) into the
// 48-B records the rasterizers gather with cp.async (gsid == nullptr: one record per Gaussian).
__global__ void __launch_bounds__(256) k_pack(int64_t P, const int32_t *__restrict__ gsid,
                                              const float2 *__restrict__ us,
                                              const float *__restrict__ cinv2ds,
                                              const float *__restrict__ alphas,
                                              const float *__restrict__ colors, Rec *__restrict__ recs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int g = gsid != nullptr ? __ldg(gsid + p) : (int)p;  // nullptr: one record per Gaussian
  const float2 u = __ldg(us + g);
  recs[p] = build_record(u.x, u.y, __ldg(cinv2ds + 3 * (size_t)g), __ldg(cinv2ds + 3 * (size_t)g + 1),
                         __ldg(cinv2ds + 3 * (size_t)g + 2), __ldg(alphas + g), __ldg(colors + 3 * (size_t)g),
                         __ldg(colors + 3 * (size_t)g + 1), __ldg(colors + 3 * (size_t)g + 2), g);
}

// ---------------------------------------------------------------- phase 1 host side
BinLayout bin_layout(int N) {
  BinLayout L{};
  const size_t n = (size_t)(N > 0 ? N : 1);
  size_t o = 0;
  L.rects = o;   o = align_up(o + n * sizeof(uint2), 256);
  L.offsets = o; o = align_up(o + n * sizeof(uint32_t), 256);
  L.total = o;   o = align_up(o + 4 * sizeof(uint32_t), 256);  // [P, max depth key, flags, scan claim]
  L.scan_desc = o;
  o = align_up(o + ((n + RS_SPAN - 1) / RS_SPAN) * sizeof(unsigned long long), 256);
  L.bytes = o;
  return L;
}

int launch_bin(int H, int W, int N, const float *us, float *depths, int32_t *areas, void *ws,
               const BinLayout &L, cudaStream_t st) {
  char *b = static_cast<char *>(ws);
  uint32_t *total = reinterpret_cast<uint32_t *>(b + L.total);
  // [P, max key, flags, claim] and the look-back descriptors start at zero
  GSB_CUDA_TRY(cudaMemsetAsync(total, 0, L.bytes - L.total, st));
  if (N <= 0) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  ProfScope ps(K_RECTS, st);
  k_rects_scan<<<(N + RS_SPAN - 1) / RS_SPAN, RS_CTA, 0, st>>>(
      N, reinterpret_cast<const float2 *>(us), reinterpret_cast<int2 *>(areas), depths, gx, gy,
      reinterpret_cast<uint2 *>(b + L.rects), reinterpret_cast<uint32_t *>(b + L.offsets), total,
      reinterpret_cast<unsigned long long *>(b + L.scan_desc), total + 3);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- phase 2 host side
int sort_layout(int N, int H, int W, int64_t P, SortLayout *out) {
  SortLayout L{};
  const size_t T = (size_t)((W + TILE - 1) / TILE) * (size_t)((H + TILE - 1) / TILE);
  const size_t n = (size_t)(P > 0 ? P : 1);
  const size_t ng = (size_t)(N > 0 ? N : 1);
  const size_t tiles = (n + RX_TILE - 1) / RX_TILE;
  size_t o = 0;
  L.keys_a = o; o = align_up(o + n * sizeof(uint64_t), 256);
  L.keys_b = o; o = align_up(o + n * sizeof(uint64_t), 256);
  L.vals_a = o; o = align_up(o + n * sizeof(int32_t), 256);
  L.vals_b = o; o = align_up(o + n * sizeof(int32_t), 256);
  L.recs = o;   o = align_up(o + ng * sizeof(Rec), 256);  // one record per Gaussian
  L.counters = o; o = align_up(o + (4 * T + 2) * sizeof(int), 256);  // rasterizer work area (launch_tile_list)
  // zeroed per call: [status | global digit histograms | per-tile digit counts of every pass]
  L.sort_state = o;
  L.hist = o + 64;
  L.desc = L.hist + (size_t)MAX_PASSES * 512 * sizeof(uint32_t);
  L.desc_stride = tiles * 512 * sizeof(uint32_t);
  o = align_up(L.desc + (size_t)MAX_PASSES * L.desc_stride, 256);
  L.sort_state_bytes = o - L.sort_state;
  L.bytes = o;
  *out = L;
  return 0;
}

int launch_pack_only(int64_t P, const int32_t *gsid_per_patch, const float *us, const float *cinv2ds,
                     const float *alphas, const float *colors, Rec *recs, cudaStream_t st) {
  if (P <= 0) return 0;
  ProfScope ps(K_PACK, st);
  k_pack<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, gsid_per_patch, reinterpret_cast<const float2 *>(us),
                                                      cinv2ds, alphas, colors, recs);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

template <typename KeyT>
static int keys_sort_ranges(int N, int64_t P_cap, const float *depths, const uint32_t *incl, const uint2 *rects,
                            uint32_t *total, int gx, int T, const KeyPlan &kp, char *b, const SortLayout &SL,
                            int32_t *ranges, int32_t *gsid_per_patch, const StatusRead *sr, cudaStream_t st) {
  const uint32_t *P_dev = total;
  KeyT *ka = reinterpret_cast<KeyT *>(b + SL.keys_a), *kb = reinterpret_cast<KeyT *>(b + SL.keys_b);
  int32_t *va = reinterpret_cast<int32_t *>(b + SL.vals_a), *vb = reinterpret_cast<int32_t *>(b + SL.vals_b);
  uint32_t *state = reinterpret_cast<uint32_t *>(b + SL.sort_state);
  uint32_t *hist = reinterpret_cast<uint32_t *>(b + SL.hist);
  const PassPlan &pp = kp.pp;
  const int bins = 1 << pp.bits;
  const int sms = sm_count();
  // only the part of the state the passes of this call use is cleared
  const size_t tiles = (size_t)((P_cap + RX_TILE - 1) / RX_TILE);
  const size_t pass_bytes = tiles * bins * sizeof(uint32_t);  // (<= SL.desc_stride: laid out contiguously)
  GSB_CUDA_TRY(cudaMemsetAsync(state, 0, (SL.desc - SL.sort_state) + (size_t)pp.n * pass_bytes, st));
  auto tile_counts = [&](int p) { return reinterpret_cast<uint32_t *>(b + SL.desc + (size_t)p * pass_bytes); };
  // k_keys writes (ka, va); pass p reads what pass p-1 wrote; keys alternate ka <-> kb, values
  // va <-> vb except that the last pass (or k_keys itself when nothing has to be sorted) writes
  // the caller's gsid_per_patch
  KeyT *kin = ka;
  int32_t *vin = pp.n == 0 ? gsid_per_patch : va;
  {
    ProfScope ps(K_KEYS, st);
    const int grid = min((N + 255) / 256, sms * 4);
    k_keys<KeyT><<<grid, 256, (size_t)(pp.n > 0 ? pp.n : 1) * bins * sizeof(uint32_t), st>>>(
        N, depths, incl, rects, gx, kp.shift, pp, (uint32_t)P_cap, kin, vin, hist, tile_counts(0), total + 2);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  if (sr != nullptr) {  // everything the host has to validate is known now; the sort and the
    // rasterizer are enqueued behind this copy and run while the host waits for it
    GSB_CUDA_TRY(cudaMemcpyAsync(sr->host, total, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (sr->ready != nullptr) GSB_CUDA_TRY(cudaEventRecord(sr->ready, st));
  }
  {
    ProfScope ps(K_SORT, st);
    const int grid = (int)(tiles < (size_t)sms * 4 ? tiles : (size_t)sms * 4);
    for (int p = 0; p < pp.n; p++) {
      KeyT *kout = (kin == ka) ? kb : ka;
      int32_t *vout = (p == pp.n - 1) ? gsid_per_patch : ((vin == va) ? vb : va);
      uint32_t *next_C = p + 1 < pp.n ? tile_counts(p + 1) : nullptr;
      const int next_shift = p + 1 < pp.n ? pp.shift[p + 1] : 0;
      constexpr size_t dyn = RX_TILE * (sizeof(KeyT) + sizeof(int32_t));  // staging of the sorted tile
      if (pp.bits == 9) {
        static const cudaError_t attr = cudaFuncSetAttribute(k_radix_pass<KeyT, 9>,
                                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        GSB_CUDA_TRY(attr);
        k_colscan<9><<<512 / 32, 1024, 0, st>>>(P_dev, (uint32_t)P_cap, hist + p * bins, tile_counts(p));
        k_radix_pass<KeyT, 9><<<grid, RX_THREADS, dyn, st>>>(kin, vin, kout, vout, P_dev, (uint32_t)P_cap,
                                                             pp.shift[p], next_shift, tile_counts(p), next_C);
      } else {
        static const cudaError_t attr = cudaFuncSetAttribute(k_radix_pass<KeyT, 8>,
                                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        GSB_CUDA_TRY(attr);
        k_colscan<8><<<256 / 32, 1024, 0, st>>>(P_dev, (uint32_t)P_cap, hist + p * bins, tile_counts(p));
        k_radix_pass<KeyT, 8><<<grid, RX_THREADS, dyn, st>>>(kin, vin, kout, vout, P_dev, (uint32_t)P_cap,
                                                             pp.shift[p], next_shift, tile_counts(p), next_C);
      }
      kin = kout;
      vin = vout;
    }
  }
  GSB_CUDA_TRY(cudaGetLastError());
  {
    ProfScope ps(K_RANGES, st);
    const size_t want = (size_t)((P_cap + 255) / 256);
    const int grid = (int)(want < (size_t)sms * 8 ? want : (size_t)sms * 8);
    k_ranges<KeyT><<<grid, 256, 0, st>>>(P_dev, (uint32_t)P_cap, total + 2, kin, kp.shift, (uint32_t)T,
                                         reinterpret_cast<int2 *>(ranges));
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_sort_and_pack(int H, int W, int N, int64_t P, uint32_t depth_key_max, const float *us,
                         const float *cinv2ds, const float *alphas, const float *depths, const float *colors,
                         const void *bin_ws, const BinLayout &BL, void *ws, const SortLayout &SL,
                         int32_t *ranges, int32_t *gsid_per_patch, bool pack, const StatusRead *sr,
                         cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  GSB_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)gx * gy, st));
  if (P <= 0 || N <= 0) return 0;
  if (P >= ((int64_t)1 << 30)) return set_arg_error("splat: more than 2^30 patches");
  const char *bb = static_cast<const char *>(bin_ws);
  char *b = static_cast<char *>(ws);
  const uint32_t *incl = reinterpret_cast<const uint32_t *>(bb + BL.offsets);
  const uint2 *rects = reinterpret_cast<const uint2 *>(bb + BL.rects);
  uint32_t *total = reinterpret_cast<uint32_t *>(const_cast<char *>(bb) + BL.total);  // [P, max key, flags]
  const KeyPlan kp = key_plan(H, W, depth_key_max);
  int rc = kp.narrow ? keys_sort_ranges<uint32_t>(N, P, depths, incl, rects, total, gx, gx * gy, kp, b, SL, ranges,
                                                  gsid_per_patch, sr, st)
                     : keys_sort_ranges<uint64_t>(N, P, depths, incl, rects, total, gx, gx * gy, kp, b, SL, ranges,
                                                  gsid_per_patch, sr, st);
  if (rc || !pack) return rc;  // !pack: the caller already holds the per-Gaussian records
  return launch_pack_only(N, nullptr, us, cinv2ds, alphas, colors, reinterpret_cast<Rec *>(b + SL.recs), st);
}

}  // namespace gsb

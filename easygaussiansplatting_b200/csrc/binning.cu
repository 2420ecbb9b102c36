// Tile binning: per-Gaussian tile rectangles, exclusive offsets, (tile | depth-mm) keys,
// duplicate-key radix sort, per-tile ranges and the packed per-patch record stream.
//
// Replaces getRects / thrust::inclusive_scan / createKeys / thrust::sort_by_key / getRanges
// (reference kernel.cu:46-150, gausplat.cu:50-91).  Scan and sort use CUB from the CUDA
// toolkit (the reference uses the same toolkit's Thrust); everything is stream-ordered in
// caller-provided workspace (the reference cudaMallocs five device_vectors per call).
//
// Sort width.  The reference sorts all 64 bits of (tile << 32 | (uint32)(depth*1000)).  Only
// ceil(log2(tiles)) + ceil(log2(max depth key + 1)) of them can differ, and phase 1 already
// returns to the host for the patch count, so it also returns the largest depth key; when the
// two fields fit 32 bits the keys are packed as (tile << dbits | depth key) in 32-bit words
// -- the same (tile, depth-mm, id) order with half the key traffic and 4 instead of 6-8
// radix passes at 1080p.  Otherwise the 64-bit layout is used.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "kernels.h"

namespace gsb {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int bits_for(uint64_t n_values) {  // bits needed to represent 0 .. n_values-1
  int b = 0;
  while (((uint64_t)1 << b) < n_values) b++;
  return b;
}

// rect packed as (x0 | x1 << 16, y0 | y1 << 16); tile grids up to 65535 x 65535
__global__ void __launch_bounds__(256) k_rects(int N, const float2 *__restrict__ us,
                                               int2 *__restrict__ areas, float *__restrict__ depths,
                                               int gx, int gy, uint2 *__restrict__ rects,
                                               uint32_t *__restrict__ counts, uint32_t *__restrict__ max_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = 0, dk = 0;
  uint2 rect = make_uint2(0u, 0u);
  if (i < N) {
    const float d = depths[i];
    if (!(d < MIN_DEPTH)) {
      const float2 u = __ldg(us + i);
      const int2 ar = areas[i];
      const float xs = (float)ar.x, ys = (float)ar.y;
      // kernel.cu:105-110; DIV_ROUND_UP(X,16) on floats is ((X) + 16 - 1) / 16 (common.cuh:15).
      // Explicit _rn intrinsics: these decide the patch list, no contraction / reassociation.
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
        dk = __float2uint_rz(__fmul_rn(d, 1000.0f));  // kernel.cu:73
      }
    }
    rects[i] = rect;
    counts[i] = n;
  }
  // one atomicMax per CTA (a per-warp atomic on a single address serialises 31k warps)
  __shared__ uint32_t s_max[8];
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, dk);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = wmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) m = max(m, s_max[w]);
    if (m > 0) atomicMax(max_key, m);
  }
}

__global__ void k_total(int N, const uint32_t *__restrict__ incl, uint32_t *__restrict__ total) {
  total[0] = N > 0 ? incl[N - 1] : 0u;
}

// one thread per Gaussian, serial over its rectangle (2.5 patches/Gaussian on the
// benchmark scenes).  64-bit: tile << 32 | dk (kernel.cu:71-74); 32-bit: tile << dbits | dk.
template <typename KeyT>
__global__ void __launch_bounds__(256) k_keys(int N, const float *__restrict__ depths,
                                              const uint32_t *__restrict__ incl,
                                              const uint2 *__restrict__ rects, int gx, int shift,
                                              KeyT *__restrict__ keys, int32_t *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t off = 0, x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  KeyT dk = 0;
  if (i < N) {
    const float d = __ldg(depths + i);
    if (!(d < MIN_DEPTH)) {
      off = (i == 0) ? 0u : __ldg(incl + i - 1);
      const uint2 r = __ldg(rects + i);
      x0 = r.x & 0xffffu; x1 = r.x >> 16; y0 = r.y & 0xffffu; y1 = r.y >> 16;
      dk = (KeyT)__float2uint_rz(__fmul_rn(d, 1000.0f));
    }
  }
  const uint32_t w = x1 - x0, n = w * (y1 - y0);
  // small rectangles: the owning lane emits them; large ones (a Gaussian covering dozens or
  // hundreds of tiles) are emitted by the whole warp, 32 patches per step, so one huge
  // footprint does not serialise its warp
  if (n <= 32) {
    for (uint32_t y = y0; y < y1; y++)
      for (uint32_t x = x0; x < x1; x++) {
        keys[off] = ((KeyT)(y * (uint32_t)gx + x) << shift) | dk;
        vals[off] = i;
        off++;
      }
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
      keys[boff + t] = ((KeyT)(y * (uint32_t)gx + x) << shift) | bdk;
      vals[boff + t] = bi;
    }
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(256) k_ranges(int64_t P, const KeyT *__restrict__ keys, int shift,
                                                int2 *__restrict__ ranges) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const uint32_t t = (uint32_t)(keys[p] >> shift);
  if (p == 0 || (uint32_t)(keys[p - 1] >> shift) != t) ranges[t].x = (int)p;
  if (p == P - 1 || (uint32_t)(keys[p + 1] >> shift) != t) ranges[t].y = (int)(p + 1);
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

// ---------------------------------------------------------------- phase 1
BinLayout bin_layout(int N) {
  BinLayout L{};
  const size_t n = (size_t)(N > 0 ? N : 1);
  size_t o = 0;
  L.rects = o;   o = align_up(o + n * sizeof(uint2), 256);
  L.counts = o;  o = align_up(o + n * sizeof(uint32_t), 256);
  L.offsets = o; o = align_up(o + n * sizeof(uint32_t), 256);
  L.total = o;   o = align_up(o + 2 * sizeof(uint32_t), 256);  // [P, max depth key]
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  L.scan_tmp = o;
  L.scan_tmp_bytes = tmp > 0 ? tmp : 256;
  o = align_up(o + L.scan_tmp_bytes, 256);
  L.bytes = o;
  return L;
}

int launch_bin(int H, int W, int N, const float *us, float *depths, int32_t *areas, void *ws,
               const BinLayout &L, cudaStream_t st) {
  char *b = static_cast<char *>(ws);
  uint32_t *total = reinterpret_cast<uint32_t *>(b + L.total);
  GSB_CUDA_TRY(cudaMemsetAsync(total, 0, 2 * sizeof(uint32_t), st));
  if (N <= 0) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  uint2 *rects = reinterpret_cast<uint2 *>(b + L.rects);
  uint32_t *counts = reinterpret_cast<uint32_t *>(b + L.counts);
  uint32_t *incl = reinterpret_cast<uint32_t *>(b + L.offsets);
  {
    ProfScope ps(K_RECTS, st);
    k_rects<<<(N + 255) / 256, 256, 0, st>>>(N, reinterpret_cast<const float2 *>(us),
                                             reinterpret_cast<int2 *>(areas), depths, gx, gy, rects, counts,
                                             total + 1);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  size_t tmp = L.scan_tmp_bytes;
  {
    ProfScope ps(K_SCAN, st);
    GSB_CUDA_TRY(cub::DeviceScan::InclusiveSum(b + L.scan_tmp, tmp, counts, incl, N, st));
    k_total<<<1, 1, 0, st>>>(N, incl, total);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- phase 2
struct KeyPlan {
  bool narrow;   // 32-bit keys
  int shift;     // tile field starts at this bit
  int end_bit;   // sort bits [0, end_bit)
};
static KeyPlan key_plan(int H, int W, uint32_t depth_key_max) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int tbits = bits_for((uint64_t)gx * (uint64_t)gy);
  const int dbits = depth_key_max == 0xffffffffu ? 32 : bits_for((uint64_t)depth_key_max + 1);
  KeyPlan p;
  p.narrow = tbits + dbits <= 32;
  p.shift = p.narrow ? dbits : 32;
  p.end_bit = p.shift + tbits;
  return p;
}

int sort_layout(int N, int H, int W, int64_t P, SortLayout *out) {
  SortLayout L{};
  const size_t n = (size_t)(P > 0 ? P : 1);
  const size_t ng = (size_t)(N > 0 ? N : 1);
  size_t o = 0;
  L.keys_a = o; o = align_up(o + n * sizeof(uint64_t), 256);
  L.keys_b = o; o = align_up(o + n * sizeof(uint64_t), 256);
  L.vals_a = o; o = align_up(o + n * sizeof(int32_t), 256);
  L.recs = o;   o = align_up(o + ng * sizeof(Rec), 256);  // one record per Gaussian
  L.counters = o; o = align_up(o + 64, 256);  // persistent-kernel tile counter
  size_t tmp64 = 0, tmp32 = 0;
  const KeyPlan wide = key_plan(H, W, 0xffffffffu);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp64, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                                  (int32_t *)nullptr, (int32_t *)nullptr, (int64_t)n, 0,
                                                  wide.end_bit);
  if (e == cudaSuccess)
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp32, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (int32_t *)nullptr, (int32_t *)nullptr, (int64_t)n, 0, 32);
  if (e != cudaSuccess) return set_cuda_error(e, "cub::DeviceRadixSort size query", __FILE__, __LINE__);
  L.sort_tmp = o;
  L.sort_tmp_bytes = tmp64 > tmp32 ? tmp64 : tmp32;
  if (L.sort_tmp_bytes == 0) L.sort_tmp_bytes = 256;
  o = align_up(o + L.sort_tmp_bytes, 256);
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
static int keys_sort_ranges(int N, int64_t P, const float *depths, const uint32_t *incl, const uint2 *rects,
                            int gx, const KeyPlan &kp, char *b, const SortLayout &SL, int32_t *ranges,
                            int32_t *gsid_per_patch, cudaStream_t st) {
  KeyT *keys_a = reinterpret_cast<KeyT *>(b + SL.keys_a);
  KeyT *keys_b = reinterpret_cast<KeyT *>(b + SL.keys_b);
  int32_t *vals_a = reinterpret_cast<int32_t *>(b + SL.vals_a);
  {
    ProfScope ps(K_KEYS, st);
    k_keys<KeyT><<<(N + 255) / 256, 256, 0, st>>>(N, depths, incl, rects, gx, kp.shift, keys_a, vals_a);
  }
  GSB_CUDA_TRY(cudaGetLastError());
  size_t tmp = SL.sort_tmp_bytes;
  {
    ProfScope ps(K_SORT, st);
    GSB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b + SL.sort_tmp, tmp, keys_a, keys_b, vals_a, gsid_per_patch,
                                                 P, 0, kp.end_bit, st));
  }
  {
    ProfScope ps(K_RANGES, st);
    k_ranges<KeyT><<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, keys_b, kp.shift,
                                                                 reinterpret_cast<int2 *>(ranges));
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_sort_and_pack(int H, int W, int N, int64_t P, uint32_t depth_key_max, const float *us,
                         const float *cinv2ds, const float *alphas, const float *depths, const float *colors,
                         const void *bin_ws, const BinLayout &BL, void *ws, const SortLayout &SL,
                         int32_t *ranges, int32_t *gsid_per_patch, bool pack, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  GSB_CUDA_TRY(cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)gx * gy, st));
  if (P <= 0 || N <= 0) return 0;
  const char *bb = static_cast<const char *>(bin_ws);
  char *b = static_cast<char *>(ws);
  const uint32_t *incl = reinterpret_cast<const uint32_t *>(bb + BL.offsets);
  const uint2 *rects = reinterpret_cast<const uint2 *>(bb + BL.rects);
  const KeyPlan kp = key_plan(H, W, depth_key_max);
  int rc = kp.narrow ? keys_sort_ranges<uint32_t>(N, P, depths, incl, rects, gx, kp, b, SL, ranges, gsid_per_patch, st)
                     : keys_sort_ranges<uint64_t>(N, P, depths, incl, rects, gx, kp, b, SL, ranges, gsid_per_patch, st);
  if (rc || !pack) return rc;  // !pack: the caller already holds the per-Gaussian records
  return launch_pack_only(N, nullptr, us, cinv2ds, alphas, colors, reinterpret_cast<Rec *>(b + SL.recs), st);
}

}  // namespace gsb

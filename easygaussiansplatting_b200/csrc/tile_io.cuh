// Shared-memory staging of array-of-struct rows: coalesced float4 global traffic, odd row
// stride in shared memory so a thread can walk its own row without bank conflicts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsb {

constexpr int PG = 128;  // Gaussians (threads) per CTA

template <int K>
struct TileT {
  static constexpr int S = (K & 1) ? K : K + 1;  // odd row stride -> conflict-free rows
  static constexpr int FLOATS = PG * S;
};

__device__ __forceinline__ bool aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// global [rows, K] (contiguous) -> smem tile
template <int K>
__device__ __forceinline__ void tile_fetch(const float *__restrict__ g, long long base_row,
                                           int n_valid, float *sm, int tid) {
  constexpr int S = TileT<K>::S;
  const float *src = g + base_row * K;
  if (n_valid == PG && aligned16(src)) {
    const float4 *src4 = reinterpret_cast<const float4 *>(src);
#pragma unroll 4
    for (int v = tid; v < PG * K / 4; v += PG) {
      float4 x = __ldg(src4 + v);
      int e = 4 * v;
      sm[((e + 0) / K) * S + (e + 0) % K] = x.x;
      sm[((e + 1) / K) * S + (e + 1) % K] = x.y;
      sm[((e + 2) / K) * S + (e + 2) % K] = x.z;
      sm[((e + 3) / K) * S + (e + 3) % K] = x.w;
    }
  } else {
    for (int e = tid; e < n_valid * K; e += PG) sm[(e / K) * S + e % K] = __ldg(src + e);
  }
}

// smem tile -> global [rows, K]
template <int K>
__device__ __forceinline__ void tile_flush(float *__restrict__ g, long long base_row,
                                           int n_valid, const float *sm, int tid) {
  constexpr int S = TileT<K>::S;
  float *dst = g + base_row * K;
  if (n_valid == PG && aligned16(dst)) {
    float4 *dst4 = reinterpret_cast<float4 *>(dst);
#pragma unroll 4
    for (int v = tid; v < PG * K / 4; v += PG) {
      int e = 4 * v;
      float4 x;
      x.x = sm[((e + 0) / K) * S + (e + 0) % K];
      x.y = sm[((e + 1) / K) * S + (e + 1) % K];
      x.z = sm[((e + 2) / K) * S + (e + 2) % K];
      x.w = sm[((e + 3) / K) * S + (e + 3) % K];
      dst4[v] = x;
    }
  } else {
    for (int e = tid; e < n_valid * K; e += PG) dst[e] = sm[(e / K) * S + e % K];
  }
}

#define ROW(K) (sm + tid * TileT<K>::S)

}  // namespace gsb

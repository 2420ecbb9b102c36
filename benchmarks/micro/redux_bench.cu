// Microbenchmark (B200): warp-instruction throughput of REDUX.SUM (redux.sync.add.s32), SHFL.BFLY,
// LDS.128, STS.64, REDG and their overlap with FFMA -- design input for the backward rasterizer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux_bench redux_bench.cu && ./redux_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, int *out, float *fout) {
  __shared__ float4 sm[512];
  int v = threadIdx.x * 7 + blockIdx.x;
  float f = (float)v, g = 1.0001f, h = 0.5f;
  int acc = 0;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x]);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x & 7) * 3]);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0) {  // REDUX only
        acc += __reduce_add_sync(0xffffffffu, v + u + i);
      } else if (MODE == 1) {  // SHFL only
        acc += __shfl_xor_sync(0xffffffffu, v + u + i, 1 + (u & 15));
      } else if (MODE == 2) {  // FFMA only (4 per unit)
        f = fmaf(f, g, h); g = fmaf(g, f, h); h = fmaf(h, g, f); f = fmaf(f, h, g);
      } else if (MODE == 3) {  // REDUX + 4 FFMA
        acc += __reduce_add_sync(0xffffffffu, v + u + i);
        f = fmaf(f, g, h); g = fmaf(g, f, h); h = fmaf(h, g, f); f = fmaf(f, h, g);
      } else if (MODE == 4) {  // SHFL + 4 FFMA
        acc += __shfl_xor_sync(0xffffffffu, v + u + i, 1 + (u & 15));
        f = fmaf(f, g, h); g = fmaf(g, f, h); h = fmaf(h, g, f); f = fmaf(f, h, g);
      } else if (MODE == 5 || MODE == 6 || MODE == 11 || MODE == 12) {  // LDS.128: distinct / 8 addr / uniform / 4 addr
        float4 q;
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(&sm[0]);
        const uint32_t rot = ((i * 8 + u) & 7) * 16;
        const uint32_t ad = MODE == 5 ? base + threadIdx.x % 32 * 16 + rot
                          : MODE == 6 ? base + (threadIdx.x & 7) * 48 + rot
                          : MODE == 11 ? base + rot : base + (threadIdx.x & 3) * 16 + rot * 4;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(ad));
        f += q.x + q.w;
      } else if (MODE == 13) {  // LDS.64 uniform
        float2 q;
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(&sm[0]);
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(q.x), "=f"(q.y) : "r"(base + ((i * 8 + u) & 7) * 16));
        f += q.x + q.y;
      } else if (MODE == 7 || MODE == 14 || MODE == 15) {  // STS.64 distinct / STS.128 uniform / STS.128 lane 0 only
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(&sm[0]);
        const uint32_t rot = ((i * 8 + u) & 7) * 16;
        if (MODE == 7) asm volatile("st.volatile.shared.v2.f32 [%0], {%1,%2};" ::"r"(base + threadIdx.x % 32 * 8 + rot), "f"(f), "f"(g) : "memory");
        if (MODE == 14) asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(base + rot), "f"(g), "f"(g), "f"(h), "f"(h) : "memory");
        if (MODE == 15 && (threadIdx.x & 31) == 0) asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(base + rot), "f"(g), "f"(g), "f"(h), "f"(h) : "memory");
        f += 1.0f;
      } else if (MODE == 8) {  // REDG spread, 9 lanes active
        if ((threadIdx.x & 31) < 9) atomicAdd(fout + ((size_t)(v * 977 + i * 131 + u) & 0xfffff) * 9 + (threadIdx.x & 31), f);
      } else if (MODE == 9) {  // REDG spread, 32 lanes, 4-byte
        atomicAdd(fout + ((size_t)((v >> 2) * 977 + i * 131 + u) & 0xfffff) * 9 + (threadIdx.x & 3), f);
      } else if (MODE == 10) {  // red.v4.f32 (16-byte vector reduction), all lanes
        float *p = fout + (((size_t)(v * 977 + i * 131 + u) & 0xfffff) * 16);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f), "f"(g), "f"(h), "f"(f) : "memory");
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  fout[blockIdx.x * blockDim.x + threadIdx.x] = f + g + h;
}

template <int MODE>
void run(const char *name, int units_per_iter, int *out, float *fout) {
  const int iters = (MODE >= 8 && MODE <= 10) ? 200 : 2000, grid = 148 * 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, 128>>>(10, out, fout);
  cudaEventRecord(a);
  k<MODE><<<grid, 128>>>(iters, out, fout);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  // warp-instructions of the measured kind per SM per cycle (1965 MHz nominal)
  double winst = (double)grid * 4 * iters * 8 * units_per_iter;
  double cyc = ms * 1e-3 * 1.965e9;
  printf("%-28s %8.3f ms  %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SM)  err=%s\n", name, ms, winst / 148 / cyc,
         148 * cyc / winst, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int *out; float *fout;
  cudaMalloc(&out, 148 * 8 * 128 * 4);
  cudaMalloc(&fout, (size_t)(1 << 20) * 16 * 4 + 4096);
  cudaMemset(fout, 0, (size_t)(1 << 20) * 16 * 4 + 4096);
  run<0>("REDUX.SUM", 1, out, fout);
  run<1>("SHFL.BFLY", 1, out, fout);
  run<2>("FFMA x4", 4, out, fout);
  run<3>("REDUX + 4 FFMA (per REDUX)", 1, out, fout);
  run<4>("SHFL + 4 FFMA (per SHFL)", 1, out, fout);
  run<5>("LDS.128 32 distinct", 1, out, fout);
  run<6>("LDS.128 8 addr (48B stride)", 1, out, fout);
  run<11>("LDS.128 uniform address", 1, out, fout);
  run<12>("LDS.128 4 distinct addr", 1, out, fout);
  run<13>("LDS.64 uniform address", 1, out, fout);
  run<7>("STS.64 distinct", 1, out, fout);
  run<14>("STS.128 uniform all lanes", 1, out, fout);
  run<15>("STS.128 lane 0 only", 1, out, fout);
  run<8>("REDG 9 lanes spread", 1, out, fout);
  run<9>("REDG 32 lanes (8 rows x4)", 1, out, fout);
  run<10>("RED.v4.f32 32 lanes", 1, out, fout);
  return 0;
}

// Is MUFU.RCP exact at 1.0 (and at other powers of two)?  The backward rasterizer relies on
// tau * rcp(1 - 0) == tau for pixels that skip a record.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float *in, float *out, int n) {
  int i = threadIdx.x;
  if (i < n) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(in[i])); out[i] = y; }
}
int main() {
  float h[8] = {1.0f, 2.0f, 0.5f, 4.0f, 0.25f, 1.0f - 0.002f, 0.01f, 1024.0f}, o[8];
  float *d, *e; cudaMalloc(&d, 32); cudaMalloc(&e, 32);
  cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(d, e, 8);
  cudaMemcpy(o, e, 32, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 8; i++) printf("rcp.approx(%g) = %.9g  exact=%d\n", h[i], o[i], o[i] == 1.0f / h[i]);
  return 0;
}

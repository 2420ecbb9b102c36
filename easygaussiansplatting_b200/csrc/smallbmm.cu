// Batched tiny matrix product C[n] = A[n] B[n] (or A[n] B with one shared B), one thread per
// batch element.  The reference's GSFunction.backward chains the per-Gaussian Jacobians with
// ~12 torch matmuls of shapes like [N,1,3]@[N,3,3] or [N,1,6]@[N,6,4] (gsmodel.py:72-85); at
// N = 1M torch dispatches each of them to batched GEMV/GEMM library kernels in chunks of
// 65535 matrices (13 ms of a 15 ms step).  These products are pure streaming -- every operand
// is read once -- so a plain SIMT kernel at HBM speed is all they need.  ops.py routes torch's
// matmul to this kernel for the Jacobian tensors it returns (JacobianTensor), which keeps the
// reference's gsmodel.py unmodified.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

constexpr int BMM_THREADS = 128;

template <int M, int K, int NN>
__global__ void __launch_bounds__(BMM_THREADS) k_small_bmm_t(long long batch, const float *__restrict__ A,
                                                             const float *__restrict__ B, int b_shared,
                                                             float *__restrict__ C) {
  const long long n = (long long)blockIdx.x * BMM_THREADS + threadIdx.x;
  if (n >= batch) return;
  const float *a = A + n * (M * K);
  const float *b = b_shared ? B : B + n * (K * NN);
  float *c = C + n * (M * NN);
  float ar[M * K];
#pragma unroll
  for (int i = 0; i < M * K; i++) ar[i] = __ldg(a + i);
#pragma unroll
  for (int j = 0; j < NN; j++) {
    float bc[K];
#pragma unroll
    for (int l = 0; l < K; l++) bc[l] = __ldg(b + l * NN + j);
#pragma unroll
    for (int i = 0; i < M; i++) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < K; l++) s = fmaf(ar[i * K + l], bc[l], s);
      c[i * NN + j] = s;
    }
  }
}

__global__ void __launch_bounds__(BMM_THREADS) k_small_bmm_generic(long long batch, int M, int K, int NN,
                                                                   const float *__restrict__ A,
                                                                   const float *__restrict__ B, int b_shared,
                                                                   float *__restrict__ C) {
  const long long n = (long long)blockIdx.x * BMM_THREADS + threadIdx.x;
  if (n >= batch) return;
  const float *a = A + n * (long long)(M * K);
  const float *b = b_shared ? B : B + n * (long long)(K * NN);
  float *c = C + n * (long long)(M * NN);
  for (int i = 0; i < M; i++)
    for (int j = 0; j < NN; j++) {
      float s = 0.f;
      for (int l = 0; l < K; l++) s = fmaf(__ldg(a + i * K + l), __ldg(b + l * NN + j), s);
      c[i * NN + j] = s;
    }
}

#define GSB_BMM_CASE(M_, K_, N_)                                                                   \
  if (M == M_ && K == K_ && NN == N_) {                                                            \
    k_small_bmm_t<M_, K_, N_><<<grid, BMM_THREADS, 0, st>>>(batch, A, B, b_shared, C);             \
    GSB_CUDA_TRY(cudaGetLastError());                                                              \
    return 0;                                                                                      \
  }

int launch_small_bmm(long long batch, int M, int K, int NN, const float *A, const float *B, int b_shared,
                     float *C, cudaStream_t st) {
  if (batch <= 0 || M <= 0 || K <= 0 || NN <= 0) return 0;
  const unsigned grid = (unsigned)((batch + BMM_THREADS - 1) / BMM_THREADS);
  ProfScope ps(K_BMM, st);
  // the shapes of gsmodel.py:72-85
  GSB_BMM_CASE(1, 3, 3) GSB_BMM_CASE(1, 3, 6) GSB_BMM_CASE(1, 6, 4) GSB_BMM_CASE(1, 6, 3)
  GSB_BMM_CASE(1, 2, 3) GSB_BMM_CASE(3, 1, 16) GSB_BMM_CASE(3, 1, 9) GSB_BMM_CASE(3, 1, 4) GSB_BMM_CASE(3, 1, 1)
  k_small_bmm_generic<<<grid, BMM_THREADS, 0, st>>>(batch, M, K, NN, A, B, b_shared, C);
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

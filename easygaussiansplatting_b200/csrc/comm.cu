// Multi-GPU gradient exchange over peer memory: the SUM all-reduce of the parameter
// gradients as a push reduce-scatter fused into the producing kernel + one reduce/broadcast
// kernel, instead of an NCCL all-reduce after the backward.
//
//   k_preprocess_bwd<.., PUSH> (fused.cu)  every rank stores each gradient tile straight into the
//        staging slot of the GPU that owns those Gaussians (posted NVLink stores, overlapping
//        the per-Gaussian math tile by tile) and finally raises its arrival flag on every peer;
//   k_grad_reduce_bcast (here)             waits for all arrivals, sums the `world` slots of its
//        own rows (local HBM reads) and stores the result into the result buffer of EVERY rank
//        (float4 peer stores); the last CTA raises the rank's done flag everywhere and waits
//        for the peers' done flags, so when the kernel retires the local result is complete and
//        every peer has finished reading its staging (safe to push the next step).
//
// Traffic per rank: (world-1)/world of the 236 B/Gaussian bucket out in each phase, the same as
// a ring all-reduce, but phase 1 rides under the backward kernel and nothing is staged twice.
// Flags are monotonically increasing epochs (no reset race).  Every spin has a timeout
// (GSB_EXCHANGE_TIMEOUT_S, default 120 s: rank skew from a checkpoint, an evaluation pass or a lazy
// module load is normal) after which the region's sticky status word is set and the rows this
// rank owns are POISONED with NaN in every rank's result, so a lost peer can never pass for a
// finished sum; GradExchange.status() / the next host call reports it.
//
// Peer mapping is plain CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) on cudaMalloc'd
// regions owned by this library; the 64-byte handles travel between the processes through
// torch.distributed (parallel.py) -- plumbing, not data path.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tile_io.cuh"

namespace gsb {

ExchangeGeom exchange_geom(int N, int k3, int world) {
  ExchangeGeom G{};
  G.world = world;
  G.ks = 3 * k3;
  const long long tiles = ((long long)N + PG - 1) / PG;
  G.tiles_per_rank = (int)((tiles + world - 1) / world);
  if (G.tiles_per_rank < 1) G.tiles_per_rank = 1;
  G.rpr = (long long)G.tiles_per_rank * PG;
  G.rows_total = G.rpr * world;
  const int k[5] = {G.ks, 4, 3, 3, 1};
  long long so = 0, ro = 0;
  for (int s = 0; s < 5; s++) {
    G.seg_k[s] = k[s];
    G.slot_off[s] = so;
    G.result_off[s] = ro;
    so += G.rpr * k[s];
    ro += G.rows_total * k[s];
  }
  G.slot_floats = so;  // rpr * (ks + 11); rpr % 128 == 0 keeps every segment 16-B aligned
  G.staging_off = kCtrlBytes;
  G.result_base = G.staging_off + (size_t)world * so * sizeof(float);
  G.bytes = G.result_base + (size_t)ro * sizeof(float);
  return G;
}

namespace {

struct ReduceArgs {
  const float *staging;        // local
  float *result[kMaxWorld];    // every rank's result buffer (peer pointers, own included)
  uint32_t *done[kMaxWorld];   // every rank's done[] array
  const uint32_t *arrive;      // local arrive[]
  const uint32_t *done_local;  // local done[]
  uint32_t *counter, *status;
  unsigned long long timeout_ns;
  long long rpr, slot_floats, slot_off[5], result_off[5];
  int seg_k[5];
  int world, rank;
  uint32_t epoch;
};

__device__ __forceinline__ bool wait_flags(const uint32_t *flags, int world, uint32_t epoch, uint32_t *status,
                                           unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  for (int s = 0; s < world; s++) {
    while ((int)(ld_acquire_sys(flags + s) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {  // a peer never arrived: report, do not hang
        atomicExch(status, 1u);
        return false;
      }
      __nanosleep(200);
    }
  }
  return true;
}

template <int WORLD>
__global__ void __launch_bounds__(256) k_grad_reduce_bcast(ReduceArgs a) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = wait_flags(a.arrive, WORLD, a.epoch, a.status, a.timeout_ns) && *a.status == 0;
  __syncthreads();
  {
    const float4 *st4 = reinterpret_cast<const float4 *>(a.staging);
    const long long slot4 = a.slot_floats / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float nan = __int_as_float(0x7fc00000);
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < slot4; v += stride) {
      float4 acc = make_float4(nan, nan, nan, nan);  // a timed-out exchange poisons its rows
      if (ok) {
        acc = st4[v];
#pragma unroll
        for (int s = 1; s < WORLD; s++) {
          const float4 x = st4[s * slot4 + v];
          acc.x += x.x, acc.y += x.y, acc.z += x.z, acc.w += x.w;
        }
      }
      // float offset inside the slot -> segment -> local tile -> global tile lt * world + rank
      // (tiles are dealt round-robin; a tile is 128 * K floats, a multiple of 4, so a float4
      // never straddles two tiles)
      const long long f = 4 * v;
      int seg = 0;
#pragma unroll
      for (int s = 1; s < 5; s++) seg += f >= a.slot_off[s];
      const unsigned in_seg = (unsigned)(f - a.slot_off[seg]);  // < 2^32: a slot segment is < 16 GiB
      const unsigned tile_floats = PG * a.seg_k[seg];
      const unsigned lt = in_seg / tile_floats;
      const long long ro = a.result_off[seg] + ((long long)lt * WORLD + a.rank) * tile_floats + (in_seg - lt * tile_floats);
#pragma unroll
      for (int p = 0; p < WORLD; p++) *reinterpret_cast<float4 *>(a.result[p] + ro) = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned n = atomicAdd(a.counter, 1u);
    if (n == gridDim.x - 1) {
      *a.counter = 0;
      __threadfence_system();
      for (int p = 0; p < WORLD; p++) st_release_sys(a.done[p] + a.rank, a.epoch);
      wait_flags(a.done_local, WORLD, a.epoch, a.status, a.timeout_ns);
    }
  }
}

}  // namespace

int launch_grad_reduce_bcast(const ExchangeGeom &G, int rank, void *const *regions, uint32_t epoch,
                             cudaStream_t st) {
  ReduceArgs a{};
  char *own = static_cast<char *>(regions[rank]);
  a.staging = reinterpret_cast<const float *>(own + G.staging_off);
  for (int p = 0; p < G.world; p++) {
    char *r = static_cast<char *>(regions[p]);
    a.result[p] = reinterpret_cast<float *>(r + G.result_base);
    a.done[p] = reinterpret_cast<uint32_t *>(r + 64);
  }
  a.arrive = reinterpret_cast<const uint32_t *>(own);
  a.done_local = reinterpret_cast<const uint32_t *>(own + 64);
  a.counter = reinterpret_cast<uint32_t *>(own + 132);
  a.status = reinterpret_cast<uint32_t *>(own + 136);
  a.rpr = G.rpr;
  a.slot_floats = G.slot_floats;
  for (int s = 0; s < 5; s++) {
    a.slot_off[s] = G.slot_off[s];
    a.result_off[s] = G.result_off[s];
    a.seg_k[s] = G.seg_k[s];
  }
  a.world = G.world;
  a.rank = rank;
  a.epoch = epoch;
  static const unsigned long long timeout_ns = [] {
    const char *e = getenv("GSB_EXCHANGE_TIMEOUT_S");
    const double sec = e != nullptr ? atof(e) : 120.0;
    return (unsigned long long)((sec > 0.001 ? sec : 0.001) * 1e9);
  }();
  a.timeout_ns = timeout_ns;
  ProfScope ps(K_GRAD_EXCHANGE, st);
  int dev = 0, sms = 148;  // (per current device: ranks of one process may sit on different GPUs)
  GSB_CUDA_TRY(cudaGetDevice(&dev));
  GSB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 4;
  switch (G.world) {
    case 1: k_grad_reduce_bcast<1><<<grid, 256, 0, st>>>(a); break;
    case 2: k_grad_reduce_bcast<2><<<grid, 256, 0, st>>>(a); break;
    case 4: k_grad_reduce_bcast<4><<<grid, 256, 0, st>>>(a); break;
    case 8: k_grad_reduce_bcast<8><<<grid, 256, 0, st>>>(a); break;
    default: return set_arg_error("grad exchange: world must be 1, 2, 4 or 8");
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

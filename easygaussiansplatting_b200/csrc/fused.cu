// Fused per-Gaussian path: one forward kernel that does the work of
// project + computeCov3D + computeCov2D + sh2Color + inverseCov2D without materialising
// their 109 floats/Gaussian of Jacobians, and one backward kernel that recomputes the
// intermediates and applies the vector-Jacobian products in registers (pg_fused_math.h)
// instead of the reference's materialise -> torch.bmm chain (gsmodel.py:21-49,72-85).
// HBM traffic: forward 232 B in / 44 B out, backward 268 B in / 232 B out per Gaussian
// (SH degree 3).  The 192-byte SH rows go through the coalescing shared-memory tile of
// tile_io.cuh; the 3/4-float rows are read and written straight from registers, so all loads
// of a CTA are in flight at once (0.128 -> 0.095 ms backward, 0.074 -> 0.065 ms forward at 1M:
// 80 % / 65 % of the measured HBM peak).
#include "common.cuh"
#include "kernels.h"
#include "pg_fused_math.h"
#include "tile_io.cuh"

namespace gsb {

__device__ __forceinline__ pg::Cam load_cam(const float *__restrict__ Rcw, const float *__restrict__ tcw,
                                            const float *__restrict__ twc, float fx, float fy, float cx,
                                            float cy, float tan_fovx, float tan_fovy) {
  pg::Cam c;
#pragma unroll
  for (int i = 0; i < 9; i++) c.R[i] = __ldg(Rcw + i);
#pragma unroll
  for (int i = 0; i < 3; i++) { c.t[i] = __ldg(tcw + i); c.twc[i] = __ldg(twc + i); }
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy; c.tan_fovx = tan_fovx; c.tan_fovy = tan_fovy;
  return c;
}

#define ROWK(K) (sm + tid * TileT<K>::S)

__device__ __forceinline__ void load_small_rows(const float *__restrict__ pws, const float *__restrict__ scales,
                                                const float *__restrict__ rots, long long i, bool valid,
                                                float (&pw)[3], float (&s)[3], float (&q)[4]) {
  if (!valid) return;
  pw[0] = __ldg(pws + 3 * i); pw[1] = __ldg(pws + 3 * i + 1); pw[2] = __ldg(pws + 3 * i + 2);
  s[0] = __ldg(scales + 3 * i); s[1] = __ldg(scales + 3 * i + 1); s[2] = __ldg(scales + 3 * i + 2);
  if (aligned16(rots)) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(rots) + i);
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
  } else {
    q[0] = __ldg(rots + 4 * i); q[1] = __ldg(rots + 4 * i + 1); q[2] = __ldg(rots + 4 * i + 2); q[3] = __ldg(rots + 4 * i + 3);
  }
}

template <int K3>
__global__ void __launch_bounds__(PG) k_preprocess_fwd(
    int N, const float *__restrict__ pws, const float *__restrict__ rots, const float *__restrict__ scales,
    const float *__restrict__ shs, const float *__restrict__ Rcw, const float *__restrict__ tcw,
    const float *__restrict__ twc, float fx, float fy, float cx, float cy, float tan_fovx, float tan_fovy,
    float *__restrict__ us, float *__restrict__ cinv2ds, float *__restrict__ colors,
    float *__restrict__ depths, int32_t *__restrict__ areas, const float *__restrict__ alphas,
    Rec *__restrict__ recs) {
  constexpr int KS = 3 * K3;
  constexpr int SMF = TileT<KS>::FLOATS > TileT<4>::FLOATS ? TileT<KS>::FLOATS : TileT<4>::FLOATS;
  __shared__ float sm[SMF];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * PG;
  const int nv = min(PG, (int)(N - base));
  const bool valid = tid < nv;
  const pg::Cam cam = load_cam(Rcw, tcw, twc, fx, fy, cx, cy, tan_fovx, tan_fovy);
  float pw[3] = {0.f, 0.f, 1.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
  // The 10 floats of position / scale / rotation go straight from global memory to registers
  // (a warp's strided loads use every byte of the sectors they touch; L1 serves the repeats) and
  // are in flight together with the SH tile -- one exposed memory latency per CTA instead of
  // one per array.
  load_small_rows(pws, scales, rots, base + tid, valid, pw, s, q);
  tile_fetch<KS>(shs, base, nv, sm, tid);
  __syncthreads();
  float u[2] = {0.f, 0.f}, conic[3] = {0.f, 0.f, 0.f}, col[3] = {0.f, 0.f, 0.f}, depth = -1.f;
  int area[2] = {0, 0};
  if (valid) {
    pg::forward_one<K3>(pw, q, s, ROWK(KS), cam, u, conic, col, &depth, area);
    const long long i = base + tid;
    *reinterpret_cast<float2 *>(us + 2 * i) = make_float2(u[0], u[1]);
    cinv2ds[3 * i] = conic[0]; cinv2ds[3 * i + 1] = conic[1]; cinv2ds[3 * i + 2] = conic[2];
    colors[3 * i] = col[0]; colors[3 * i + 1] = col[1]; colors[3 * i + 2] = col[2];
    *reinterpret_cast<int2 *>(areas + 2 * i) = make_int2(area[0], area[1]);
    depths[i] = depth;
    // the rasterizers' per-Gaussian record, straight from registers (saves the k_pack pass)
    if (recs != nullptr)
      recs[i] = build_record(u[0], u[1], conic[0], conic[1], conic[2], __ldg(alphas + i), col[0], col[1], col[2], (int)i);
  }
}

// PUSH = the multi-GPU variant: instead of writing the gradient tile to local
// buffers, the CTA stores it straight into the staging slot `rank` of the GPU that owns this
// tile's Gaussians (peer memory over NVLink -- posted stores, so the transfer of tile t
// overlaps the math of tile t+1), together with dL/dalpha from the rasterizer backward.  The
// last CTA to finish raises this rank's arrival flag on every peer; k_grad_reduce_bcast
// (comm.cu) then sums the slots of the rows it owns and broadcasts the result.
// MOM = the upstream gradients come as the rasterizer backward's raw moment rows [N,9] (see
// raster_bwd.cu): the conversion k_finalize_grads would do (dL/du = -(A Sx + B Sy, B Sx + C Sy),
// dL/dconic = (-Sxx/2, -Sxy, -Syy/2), dL/dcolor, dL/dalpha) happens here in registers, dL/du and
// dL/dalpha are still written out (the caller's densification statistics / alpha gradient), and
// the separate finalize pass with its 36 B/Gaussian round trip disappears.
template <int K3, bool PUSH, bool MOM>
__global__ void __launch_bounds__(PG) k_preprocess_bwd(
    int N, const float *__restrict__ pws, const float *__restrict__ rots, const float *__restrict__ scales,
    const float *__restrict__ shs, const float *__restrict__ Rcw, const float *__restrict__ tcw,
    const float *__restrict__ twc, float fx, float fy, float cx, float cy, float tan_fovx, float tan_fovy,
    const float *__restrict__ g_us, const float *__restrict__ g_cinv2ds, const float *__restrict__ g_colors,
    float *__restrict__ g_pws, float *__restrict__ g_shs, float *__restrict__ g_scales,
    float *__restrict__ g_rots, GradPush gp, MomentsIn mi) {
  constexpr int KS = 3 * K3;
  constexpr int SMF = TileT<KS>::FLOATS > TileT<4>::FLOATS ? TileT<KS>::FLOATS : TileT<4>::FLOATS;
  __shared__ float sm[SMF];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * PG;
  const int nv = min(PG, (int)(N - base));
  const bool valid = tid < nv;
  const pg::Cam cam = load_cam(Rcw, tcw, twc, fx, fy, cx, cy, tan_fovx, tan_fovy);
  float pw[3] = {0.f, 0.f, 1.f}, q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
  float gu[2] = {0.f, 0.f}, gci[3] = {0.f, 0.f, 0.f}, gcol[3] = {0.f, 0.f, 0.f};
  load_small_rows(pws, scales, rots, base + tid, valid, pw, s, q);  // see k_preprocess_fwd
  float galpha = 0.f;
  if (MOM) {
    if (valid) {
      const long long i = base + tid;
      float m[9];
#pragma unroll
      for (int k = 0; k < 9; k++) m[k] = __ldg(mi.moments + 9 * i + k);
      const float A = __ldg(mi.cinv2ds + 3 * i), B = __ldg(mi.cinv2ds + 3 * i + 1), C = __ldg(mi.cinv2ds + 3 * i + 2);
      // untouched Gaussians stay exactly 0 even if their conic is inf/NaN (k_finalize_grads)
      const bool none = (m[0] == 0.f) && (m[1] == 0.f);
      gu[0] = none ? 0.f : -(A * m[0] + B * m[1]);
      gu[1] = none ? 0.f : -(B * m[0] + C * m[1]);
      gci[0] = -0.5f * m[2]; gci[1] = -m[3]; gci[2] = -0.5f * m[4];
      gcol[0] = m[6]; gcol[1] = m[7]; gcol[2] = m[8];
      galpha = m[5];
      *reinterpret_cast<float2 *>(mi.dus_out + 2 * i) = make_float2(gu[0], gu[1]);
      if (mi.dalphas_out != nullptr) mi.dalphas_out[i] = galpha;
    }
  } else if (valid) {
    const long long i = base + tid;
    const float2 g2 = __ldg(reinterpret_cast<const float2 *>(g_us) + i);
    gu[0] = g2.x; gu[1] = g2.y;
    gci[0] = __ldg(g_cinv2ds + 3 * i); gci[1] = __ldg(g_cinv2ds + 3 * i + 1); gci[2] = __ldg(g_cinv2ds + 3 * i + 2);
    gcol[0] = __ldg(g_colors + 3 * i); gcol[1] = __ldg(g_colors + 3 * i + 1); gcol[2] = __ldg(g_colors + 3 * i + 2);
  }
  tile_fetch<KS>(shs, base, nv, sm, tid);
  __syncthreads();
  float gpw[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
  // dL/dsh overwrites the thread's own SH row in place, then the tile is flushed
  if (valid) pg::backward_one<K3>(pw, q, s, ROWK(KS), cam, gu, gci, gcol, gpw, gq, gs, ROWK(KS));
  __syncthreads();
  long long drow = base;
  if (PUSH) {  // destination = slot `rank` on the owner of this tile, rows relative to its range
    // tiles are dealt round-robin over the ranks, so peer and local stores interleave for the
    // whole kernel instead of one NVLink burst at one end of it
    const int owner = (int)(blockIdx.x % gp.world);
    float *sb = gp.slot[owner];
    drow = (long long)(blockIdx.x / gp.world) * PG;
    g_shs = sb;
    g_rots = sb + gp.off_rots;
    g_pws = sb + gp.off_pws;
    g_scales = sb + gp.off_scales;
    if (valid) (sb + gp.off_alphas)[drow + tid] = MOM ? galpha : __ldg(gp.g_alphas + base + tid);
  }
  // g_shs == nullptr: the caller re-expands dL/dsh = sum over views of Y(dir_v) (x) dL/dcolor_v itself
  // (multi-view data parallel: 12 B/view travel instead of the 192 B row, k_sh_expand below)
  if (g_shs != nullptr) tile_flush<KS>(g_shs, drow, nv, sm, tid);
  if (valid) {  // the 10 small gradient floats: direct strided stores (no staging, no barriers)
    const long long o = drow + tid;
    g_pws[3 * o] = gpw[0]; g_pws[3 * o + 1] = gpw[1]; g_pws[3 * o + 2] = gpw[2];
    g_scales[3 * o] = gs[0]; g_scales[3 * o + 1] = gs[1]; g_scales[3 * o + 2] = gs[2];
    if (aligned16(g_rots)) {
      reinterpret_cast<float4 *>(g_rots)[o] = make_float4(gq[0], gq[1], gq[2], gq[3]);
    } else {
      g_rots[4 * o] = gq[0]; g_rots[4 * o + 1] = gq[1]; g_rots[4 * o + 2] = gq[2]; g_rots[4 * o + 3] = gq[3];
    }
  }
  if (PUSH) {
    __threadfence_system();  // this thread's peer stores are ordered before the flag below
    __syncthreads();
    if (tid == 0) {
      const unsigned done = atomicAdd(gp.counter, 1u);
      if (done == gridDim.x - 1) {  // last tile of this rank
        *gp.counter = 0;            // re-armed for the next launch (stream-ordered)
        __threadfence_system();
        for (int p = 0; p < gp.world; p++) st_release_sys(gp.flags[p] + gp.rank, gp.epoch);
      }
    }
  }
}

// dL/dsh of one Gaussian summed over V views from the views' dL/dcolor:
//   colour = 0.5 + sum_l Y_l(dir) sh_l is linear in sh and never clamped (kernel.cu:735-774), so
//   dL/dsh[l][c] = sum_v Y_l(dir_v) dL/dcolor_v[c]   with dir_v = (pw - twc_v) / |pw - twc_v|,
// the same Y as backward_one evaluates (for V = 1 the result equals the dL/dsh row the per-Gaussian
// backward writes up to the compiler's FMA contraction of the basis polynomials).  Multi-view data parallel
// training exchanges the 12-byte dL/dcolor per view instead of the 192-byte dL/dsh row and
// expands on every rank (parallel.MultiViewStep).  Views are summed in index order.
template <int K3>
__global__ void __launch_bounds__(PG) k_sh_expand(int N, int V, const float *__restrict__ pws,
                                                  const float *__restrict__ twcs, const float *__restrict__ gcols,
                                                  float *__restrict__ g_shs) {
  constexpr int KS = 3 * K3;
  __shared__ float sm[TileT<KS>::FLOATS];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * PG;
  const int nv = min(PG, (int)(N - base));
  const bool valid = tid < nv;
  float acc[KS];
#pragma unroll
  for (int i = 0; i < KS; i++) acc[i] = 0.f;
  if (valid) {
    const long long i = base + tid;
    const float pw[3] = {__ldg(pws + 3 * i), __ldg(pws + 3 * i + 1), __ldg(pws + 3 * i + 2)};
    for (int v = 0; v < V; v++) {
      const float *gc = gcols + ((size_t)v * N + i) * 3;
      const float g0 = __ldg(gc), g1 = __ldg(gc + 1), g2 = __ldg(gc + 2);
      float Y[K3], dY[K3][3];
      float r[3] = {0.f, 0.f, 0.f};
      if (K3 > 1) {
        const float d0 = pw[0] - __ldg(twcs + 3 * v), d1 = pw[1] - __ldg(twcs + 3 * v + 1),
                    d2 = pw[2] - __ldg(twcs + 3 * v + 2);
        const float ninv = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        r[0] = d0 * ninv; r[1] = d1 * ninv; r[2] = d2 * ninv;
      }
      pg::sh_basis<K3>(r[0], r[1], r[2], Y, dY);
#pragma unroll
      for (int l = 0; l < K3; l++) {
        acc[3 * l] += g0 * Y[l];
        acc[3 * l + 1] += g1 * Y[l];
        acc[3 * l + 2] += g2 * Y[l];
      }
    }
    float *row = ROWK(KS);
#pragma unroll
    for (int e = 0; e < KS; e++) row[e] = acc[e];
  }
  __syncthreads();
  tile_flush<KS>(g_shs, base, nv, sm, tid);
}

#define GSB_DISPATCH_K3(k3, CALL)                                           \
  switch (k3) {                                                             \
    case 1: { constexpr int K3 = 1; CALL; } break;                          \
    case 4: { constexpr int K3 = 4; CALL; } break;                          \
    case 9: { constexpr int K3 = 9; CALL; } break;                          \
    case 16: { constexpr int K3 = 16; CALL; } break;                        \
    default: return set_arg_error("sh_dim3 must be 1, 4, 9 or 16");         \
  }

int launch_preprocess_fwd(int N, int k3, const float *pws, const float *rots, const float *scales,
                          const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                          float fy, float cx, float cy, float width, float height, float *us, float *cinv2ds,
                          float *colors, float *depths, int32_t *areas, const float *alphas, Rec *recs,
                          cudaStream_t st) {
  if (N <= 0) return 0;
  const float tfx = width / (2 * fx), tfy = height / (2 * fy);
  const int nb = (N + PG - 1) / PG;
  ProfScope ps(K_PRE_FWD, st);
  GSB_DISPATCH_K3(k3, (k_preprocess_fwd<K3><<<nb, PG, 0, st>>>(N, pws, rots, scales, shs, Rcw, tcw, twc, fx, fy, cx,
                                                                cy, tfx, tfy, us, cinv2ds, colors, depths, areas,
                                                                alphas, recs)));
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_sh_expand(int N, int k3, int V, const float *pws, const float *twcs, const float *gcols, float *g_shs,
                     cudaStream_t st) {
  if (N <= 0) return 0;
  const int nb = (N + PG - 1) / PG;
  ProfScope ps(K_SH_EXPAND, st);
  GSB_DISPATCH_K3(k3, (k_sh_expand<K3><<<nb, PG, 0, st>>>(N, V, pws, twcs, gcols, g_shs)));
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_preprocess_bwd(int N, int k3, const float *pws, const float *rots, const float *scales,
                          const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                          float fy, float cx, float cy, float width, float height, const float *g_us,
                          const float *g_cinv2ds, const float *g_colors, float *g_pws, float *g_shs,
                          float *g_scales, float *g_rots, const MomentsIn *mi, cudaStream_t st) {
  if (N <= 0) return 0;
  const float tfx = width / (2 * fx), tfy = height / (2 * fy);
  const int nb = (N + PG - 1) / PG;
  ProfScope ps(K_PRE_BWD, st);
  if (mi != nullptr) {
    GSB_DISPATCH_K3(k3, (k_preprocess_bwd<K3, false, true><<<nb, PG, 0, st>>>(
                            N, pws, rots, scales, shs, Rcw, tcw, twc, fx, fy, cx, cy, tfx, tfy, nullptr, nullptr, nullptr,
                            g_pws, g_shs, g_scales, g_rots, GradPush{}, *mi)));
  } else {
    GSB_DISPATCH_K3(k3, (k_preprocess_bwd<K3, false, false><<<nb, PG, 0, st>>>(
                            N, pws, rots, scales, shs, Rcw, tcw, twc, fx, fy, cx, cy, tfx, tfy, g_us, g_cinv2ds, g_colors,
                            g_pws, g_shs, g_scales, g_rots, GradPush{}, MomentsIn{})));
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_preprocess_bwd_push(int N, int k3, const float *pws, const float *rots, const float *scales,
                               const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                               float fy, float cx, float cy, float width, float height, const float *g_us,
                               const float *g_cinv2ds, const float *g_colors, const GradPush &gp,
                               const MomentsIn *mi, cudaStream_t st) {
  if (N <= 0) return 0;
  const float tfx = width / (2 * fx), tfy = height / (2 * fy);
  const int nb = (N + PG - 1) / PG;
  ProfScope ps(K_PRE_BWD, st);
  if (mi != nullptr) {
    GSB_DISPATCH_K3(k3, (k_preprocess_bwd<K3, true, true><<<nb, PG, 0, st>>>(
                            N, pws, rots, scales, shs, Rcw, tcw, twc, fx, fy, cx, cy, tfx, tfy, nullptr, nullptr, nullptr,
                            nullptr, nullptr, nullptr, nullptr, gp, *mi)));
  } else {
    GSB_DISPATCH_K3(k3, (k_preprocess_bwd<K3, true, false><<<nb, PG, 0, st>>>(
                            N, pws, rots, scales, shs, Rcw, tcw, twc, fx, fy, cx, cy, tfx, tfy, g_us, g_cinv2ds, g_colors,
                            nullptr, nullptr, nullptr, nullptr, gp, MomentsIn{})));
  }
  GSB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gsb

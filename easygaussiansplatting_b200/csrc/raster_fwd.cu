// Forward tile rasterizer ("draw").  Replaces reference kernel.cu:152-271 + fetch2shared :13-44.
//
// One CTA per 16x16 tile, 8 warps; every warp owns an 8x4-pixel sub-rectangle (one pixel
// per lane) and walks the tile's depth-sorted record list on its own -- no block barrier
// per record (the reference does a __syncthreads_count per Gaussian).  Records arrive in
// shared memory as double-buffered batches moved by cp.async.bulk (1-D TMA) + mbarrier
// from the packed, sorted 48-B record stream (one contiguous span per tile).
// Per batch each warp first tests 32 records at a time, one per lane, against its 8x4
// rectangle using the record's conservative {alpha' >= 0.002} half-extents and keeps only a
// ballot mask of records that can contribute; only those are evaluated per pixel.  Skipped
// records are exactly those the reference would `continue` on for all 32 pixels, so image,
// contrib (1 + list index of the last contributing record) and final_tau are unchanged.
// A warp stops when all its pixels reached tau < 1e-4; the CTA stops when all warps did.
#include "common.cuh"
#include "kernels.h"

namespace gsb {

constexpr int DRAW_BATCH = 128;  // records per shared-memory stage (6 KB)

__global__ void __launch_bounds__(256) k_draw(int W, int H, int gx, const int2 *__restrict__ ranges,
                                              const Rec *__restrict__ recs, float *__restrict__ image,
                                              int32_t *__restrict__ contrib, float *__restrict__ final_tau) {
  __shared__ Rec sbuf[2][DRAW_BATCH];
  __shared__ __align__(8) uint64_t mbar[2];

  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // warp -> 8x4 sub-rectangle, lane -> pixel
  const int rx0 = tx * TILE + (warp & 1) * 8, ry0 = ty * TILE + (warp >> 1) * 4;
  const int px = rx0 + (lane & 7), py = ry0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)py * W + px;

  const int2 range = __ldg(ranges + tile);
  const int len = range.y - range.x;
  if (len <= 0) {  // kernel.cu:182-183: tile without patches keeps image 0, contrib 0, tau 0
    if (inside) {
      image[pix] = 0.f; image[HW + pix] = 0.f; image[2 * HW + pix] = 0.f;
      contrib[pix] = 0; final_tau[pix] = 0.f;
    }
    return;
  }
  const int nb = (len + DRAW_BATCH - 1) / DRAW_BATCH;
  const Rec *src = recs + range.x;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int b = 0; b < 2 && b < nb; b++) {
      const uint32_t bytes = (uint32_t)min(DRAW_BATCH, len - b * DRAW_BATCH) * (uint32_t)sizeof(Rec);
      mbar_expect_tx(&mbar[b], bytes);
      bulk_g2s(&sbuf[b][0], src + (size_t)b * DRAW_BATCH, bytes, &mbar[b]);
    }
  }

  const float fpx = (float)px, fpy = (float)py;
  const float bx0 = (float)rx0, bx1 = (float)(rx0 + 7), by0 = (float)ry0, by1 = (float)(ry0 + 3);
  // a pixel is finished exactly when tau < 1e-4; pixels outside the image start finished
  float tau = inside ? 1.0f : 0.0f, cr = 0.f, cg = 0.f, cb = 0.f;
  int cont = 0;
  bool warp_done = __all_sync(0xffffffffu, tau < TAU_STOP);

  int b = 0;
  for (; b < nb; b++) {
    const int s = b & 1;
    mbar_wait(&mbar[s], (b >> 1) & 1);
    const int nrec = min(DRAW_BATCH, len - b * DRAW_BATCH);
    if (!warp_done) {
      for (int c0 = 0; c0 < nrec; c0 += 32) {
        const int j = c0 + lane;
        bool hit = false;
        if (j < nrec) hit = rec_can_touch(sbuf[s][j].q0, sbuf[s][j].q1, bx0, bx1, by0, by1);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          const Rec *r = &sbuf[s][c0 + k];
          const float4 q0 = r->q0, q1 = r->q1;
          float g;
          const float ap = alpha_prime(q1, q0.x - fpx, q0.y - fpy, &g);
          if (tau >= TAU_STOP && ap >= ALPHA_SKIP) {
            const float4 q2 = r->q2;
            const float w = tau * ap;
            cr = fmaf(w, q2.x, cr);
            cg = fmaf(w, q2.y, cg);
            cb = fmaf(w, q2.z, cb);
            cont = b * DRAW_BATCH + c0 + k + 1;
            tau = tau * (1.0f - ap);
          }
        }
        warp_done = __all_sync(0xffffffffu, tau < TAU_STOP);
        if (warp_done) break;
      }
    }
    // all warps are past stage s -> it may be refilled; also the CTA-wide early out
    const int all_done = __syncthreads_and(warp_done ? 1 : 0);
    if (all_done) break;
    if (tid == 0 && b + 2 < nb) {
      const uint32_t bytes = (uint32_t)min(DRAW_BATCH, len - (b + 2) * DRAW_BATCH) * (uint32_t)sizeof(Rec);
      fence_proxy_async();
      mbar_expect_tx(&mbar[s], bytes);
      bulk_g2s(&sbuf[s][0], src + (size_t)(b + 2) * DRAW_BATCH, bytes, &mbar[s]);
    }
  }
  // early exit: the next stage's bulk copy may still be in flight into our shared memory
  if (b + 1 < nb) mbar_wait(&mbar[(b + 1) & 1], ((b + 1) >> 1) & 1);
  if (inside) {
    image[pix] = cr; image[HW + pix] = cg; image[2 * HW + pix] = cb;
    contrib[pix] = cont; final_tau[pix] = tau;
  }
}

int launch_draw(int H, int W, const int32_t *ranges, const Rec *recs, const int32_t *gsid, float *image,
                int32_t *contrib, float *final_tau, int *tile_counter, cudaStream_t st) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  if (gx <= 0 || gy <= 0) return 0;
  return launch_draw2(H, W, ranges, recs, gsid, image, contrib, final_tau, tile_counter, st);
}

}  // namespace gsb

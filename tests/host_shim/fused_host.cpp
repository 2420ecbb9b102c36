// TEST-ONLY host build of the fused per-Gaussian math (easygaussiansplatting_b200/csrc/
// pg_fused_math.h is host+device).  Lets the CPU test-suite check the analytic
// vector-Jacobian products against the oracle's materialised-Jacobian chain without a GPU.
// Never part of the product library.
#include <stdint.h>

#include "../../easygaussiansplatting_b200/csrc/pg_fused_math.h"

using namespace gsb::pg;

static Cam make_cam(const float *Rcw, const float *tcw, const float *twc, float fx, float fy, float cx,
                    float cy, float width, float height) {
  Cam c;
  for (int i = 0; i < 9; i++) c.R[i] = Rcw[i];
  for (int i = 0; i < 3; i++) { c.t[i] = tcw[i]; c.twc[i] = twc[i]; }
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
  c.tan_fovx = width / (2 * fx);
  c.tan_fovy = height / (2 * fy);
  return c;
}

template <int K3>
static void fwd(int N, const float *pws, const float *rots, const float *scales, const float *shs,
                const Cam &cam, float *us, float *cinv, float *cols, float *depths, int32_t *areas) {
  for (int i = 0; i < N; i++) {
    int ar[2];
    forward_one<K3>(pws + 3 * i, rots + 4 * i, scales + 3 * i, shs + (size_t)3 * K3 * i, cam, us + 2 * i,
                    cinv + 3 * i, cols + 3 * i, depths + i, ar);
    areas[2 * i] = ar[0];
    areas[2 * i + 1] = ar[1];
  }
}

template <int K3>
static void bwd(int N, const float *pws, const float *rots, const float *scales, const float *shs,
                const Cam &cam, const float *gu, const float *gci, const float *gcol, float *gpw, float *gsh,
                float *gs, float *gq) {
  for (int i = 0; i < N; i++)
    backward_one<K3>(pws + 3 * i, rots + 4 * i, scales + 3 * i, shs + (size_t)3 * K3 * i, cam, gu + 2 * i,
                     gci + 3 * i, gcol + 3 * i, gpw + 3 * i, gq + 4 * i, gs + 3 * i, gsh + (size_t)3 * K3 * i);
}

extern "C" {

int fused_forward_host(int N, int k3, const float *pws, const float *rots, const float *scales,
                       const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                       float fy, float cx, float cy, float width, float height, float *us, float *cinv,
                       float *cols, float *depths, int32_t *areas) {
  Cam cam = make_cam(Rcw, tcw, twc, fx, fy, cx, cy, width, height);
  switch (k3) {
    case 1: fwd<1>(N, pws, rots, scales, shs, cam, us, cinv, cols, depths, areas); break;
    case 4: fwd<4>(N, pws, rots, scales, shs, cam, us, cinv, cols, depths, areas); break;
    case 9: fwd<9>(N, pws, rots, scales, shs, cam, us, cinv, cols, depths, areas); break;
    case 16: fwd<16>(N, pws, rots, scales, shs, cam, us, cinv, cols, depths, areas); break;
    default: return -1;
  }
  return 0;
}

int fused_backward_host(int N, int k3, const float *pws, const float *rots, const float *scales,
                        const float *shs, const float *Rcw, const float *tcw, const float *twc, float fx,
                        float fy, float cx, float cy, float width, float height, const float *gu,
                        const float *gci, const float *gcol, float *gpw, float *gsh, float *gs, float *gq) {
  Cam cam = make_cam(Rcw, tcw, twc, fx, fy, cx, cy, width, height);
  switch (k3) {
    case 1: bwd<1>(N, pws, rots, scales, shs, cam, gu, gci, gcol, gpw, gsh, gs, gq); break;
    case 4: bwd<4>(N, pws, rots, scales, shs, cam, gu, gci, gcol, gpw, gsh, gs, gq); break;
    case 9: bwd<9>(N, pws, rots, scales, shs, cam, gu, gci, gcol, gpw, gsh, gs, gq); break;
    case 16: bwd<16>(N, pws, rots, scales, shs, cam, gu, gci, gcol, gpw, gsh, gs, gq); break;
    default: return -1;
  }
  return 0;
}
}

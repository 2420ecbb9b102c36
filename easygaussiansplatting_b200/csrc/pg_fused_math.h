// Per-Gaussian math of the fused preprocess path, shared by the CUDA kernels (fused.cu) and a
// host build used only by tests (tests/host_shim/fused_host.cpp) -- hence host+device, no
// CUDA-only intrinsics.
//
// forward : (pw, q, s, sh) -> u, conic, colour, depth, radius   == project + computeCov3D +
//           computeCov2D + sh2Color + inverseCov2D of the reference (kernel.cu:274-807) without
//           materialising their Jacobians;
// backward: vector-Jacobian products of the same five stages (docs/backward.md B.1-B.5.3,
//           gsmodel.py:72-85) taken analytically in registers:
//             dL/dcov2d = dL/dconic . dconic/dcov2d
//             G = sym(dL/dcov2d);  dL/dSigma = M^T G M (x2 off-diagonal);  dL/dM = 2 G M Sigma
//             dL/dJ = dL/dM Rcw^T -> dL/dpc through J(x_clamped, y_clamped, z)   (reference
//             quirk kept: derivatives taken as if x,y were unclamped, kernel.cu:527-534)
//             dL/dM3 = 2 sym(dL/dSigma) M3;  dL/ds_j = sum_i R_ij dL/dM3_ij;
//             dL/dq_p = sum_ij dR_ij/dq_p s_j dL/dM3_ij
//             dL/dsh_lc = dL/dcolour_c Y_l;  dL/dpw += (I - r r^T)/|d| sum_l dY_l (dL/dcolour . sh_l)
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD inline
#endif

namespace gsb {
namespace pg {

struct Cam {
  float R[9], t[3], twc[3];
  float fx, fy, cx, cy, tan_fovx, tan_fovy;
};

constexpr float kMinDepth = 0.2f;

GSB_HD void quat_to_R(const float q[4], float R[9]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w);       R[2] = 2.f * (x * z + y * w);
  R[3] = 2.f * (x * y + z * w);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
  R[6] = 2.f * (x * z - y * w);       R[7] = 2.f * (y * z + x * w);       R[8] = 1.f - 2.f * (x * x + y * y);
}

// real SH basis up to degree 3 and its partial derivatives wrt the (unit) direction, with
// xx, yy, zz treated as independent monomials exactly as the reference does.
template <int K3>
GSB_HD void sh_basis(float x, float y, float z, float Y[K3], float dY[K3][3]) {
#define GSB_SET(l, v, dx, dy, dz)                        \
  if (l < K3) {                                          \
    Y[l % K3] = (v);                                     \
    dY[l % K3][0] = (dx); dY[l % K3][1] = (dy); dY[l % K3][2] = (dz); \
  }
  const float C1a = -0.4886025119029199f, C1b = 0.4886025119029199f;
  const float C20 = 1.0925484305920792f, C22 = 0.31539156525252005f, C24 = 0.5462742152960396f;
  const float C30 = -0.5900435899266435f, C31 = 2.890611442640554f, C32 = -0.4570457994644658f,
              C33 = 0.3731763325901154f, C35 = 1.445305721320277f;
  GSB_SET(0, 0.28209479177387814f, 0.f, 0.f, 0.f)
  GSB_SET(1, C1a * y, 0.f, C1a, 0.f)
  GSB_SET(2, C1b * z, 0.f, 0.f, C1b)
  GSB_SET(3, C1a * x, C1a, 0.f, 0.f)
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  GSB_SET(4, C20 * xy, C20 * y, C20 * x, 0.f)
  GSB_SET(5, -C20 * yz, 0.f, -C20 * z, -C20 * y)
  GSB_SET(6, C22 * (2.f * zz - xx - yy), -2.f * C22 * x, -2.f * C22 * y, 4.f * C22 * z)
  GSB_SET(7, -C20 * xz, -C20 * z, 0.f, -C20 * x)
  GSB_SET(8, C24 * (xx - yy), 2.f * C24 * x, -2.f * C24 * y, 0.f)
  GSB_SET(9, C30 * y * (3.f * xx - yy), C30 * 6.f * xy, C30 * (3.f * xx - 3.f * yy), 0.f)
  GSB_SET(10, C31 * xy * z, C31 * yz, C31 * xz, C31 * xy)
  GSB_SET(11, C32 * y * (4.f * zz - xx - yy), C32 * (-2.f * xy), C32 * (4.f * zz - xx - 3.f * yy), C32 * 8.f * yz)
  GSB_SET(12, C33 * z * (2.f * zz - 3.f * xx - 3.f * yy), C33 * (-6.f * xz), C33 * (-6.f * yz),
          C33 * (6.f * zz - 3.f * xx - 3.f * yy))
  GSB_SET(13, C32 * x * (4.f * zz - xx - yy), C32 * (4.f * zz - 3.f * xx - yy), C32 * (-2.f * xy), C32 * 8.f * xz)
  GSB_SET(14, C35 * z * (xx - yy), C35 * 2.f * xz, C35 * (-2.f * yz), C35 * (xx - yy))
  GSB_SET(15, C30 * x * (xx - 3.f * yy), C30 * (3.f * xx - 3.f * yy), C30 * (-6.f * xy), 0.f)
#undef GSB_SET
}

// Everything the backward needs again is recomputed from the parameters by this same
// function, so forward and backward see identical intermediates.
struct Geo {
  float pc[3];        // camera-frame position
  float xc, yc;       // x, y after the 1.3*tan_fov clamp (kernel.cu:458-461)
  float zi;           // 1/z
  float R3[9], M3[9]; // R(q), R(q) diag(s)
  float S6[6];        // cov3d upper triangle
  float M[6];         // J Rcw (2x3)
  float MS[6];        // M Sigma
  float c2[3];        // cov2d (a, b, c) with the +0.3 blur
  float di;           // 1 / det(cov2d)
  bool keep;          // z >= 0.2 and 1/det not NaN
};

GSB_HD void geometry(const float pw[3], const float q[4], const float s[3], const Cam &cam, Geo &g) {
  const float *R = cam.R;
  g.pc[0] = fmaf(R[0], pw[0], fmaf(R[1], pw[1], fmaf(R[2], pw[2], cam.t[0])));
  g.pc[1] = fmaf(R[3], pw[0], fmaf(R[4], pw[1], fmaf(R[5], pw[2], cam.t[1])));
  g.pc[2] = fmaf(R[6], pw[0], fmaf(R[7], pw[1], fmaf(R[8], pw[2], cam.t[2])));
  g.keep = !(g.pc[2] < kMinDepth);
  const float z = g.keep ? g.pc[2] : 1.0f;
  g.zi = 1.0f / z;
  quat_to_R(q, g.R3);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) g.M3[3 * r + c] = g.R3[3 * r + c] * s[c];
  int n = 0;
  for (int r = 0; r < 3; r++)
    for (int c = r; c < 3; c++)
      g.S6[n++] = g.M3[3 * r] * g.M3[3 * c] + g.M3[3 * r + 1] * g.M3[3 * c + 1] + g.M3[3 * r + 2] * g.M3[3 * c + 2];
  const float limx = 1.3f * cam.tan_fovx, limy = 1.3f * cam.tan_fovy;
  const float px = g.keep ? g.pc[0] : 0.f, py = g.keep ? g.pc[1] : 0.f;
  g.xc = fminf(limx, fmaxf(-limx, px / z)) * z;
  g.yc = fminf(limy, fmaxf(-limy, py / z)) * z;
  const float zi2 = g.zi * g.zi;
  const float J00 = cam.fx * g.zi, J11 = cam.fy * g.zi;
  const float J02 = -(cam.fx * g.xc) * zi2, J12 = -(cam.fy * g.yc) * zi2;
  for (int c = 0; c < 3; c++) {
    g.M[c] = fmaf(J00, R[c], J02 * R[6 + c]);
    g.M[3 + c] = fmaf(J11, R[3 + c], J12 * R[6 + c]);
  }
  const float S[9] = {g.S6[0], g.S6[1], g.S6[2], g.S6[1], g.S6[3], g.S6[4], g.S6[2], g.S6[4], g.S6[5]};
  for (int a = 0; a < 2; a++)
    for (int c = 0; c < 3; c++)
      g.MS[3 * a + c] = g.M[3 * a] * S[c] + g.M[3 * a + 1] * S[3 + c] + g.M[3 * a + 2] * S[6 + c];
  g.c2[0] = g.MS[0] * g.M[0] + g.MS[1] * g.M[1] + g.MS[2] * g.M[2] + 0.3f;
  g.c2[1] = g.MS[0] * g.M[3] + g.MS[1] * g.M[4] + g.MS[2] * g.M[5];
  g.c2[2] = g.MS[3] * g.M[3] + g.MS[4] * g.M[4] + g.MS[5] * g.M[5] + 0.3f;
  g.di = 1.0f / (g.c2[0] * g.c2[2] - g.c2[1] * g.c2[1]);
  if (g.di != g.di) g.keep = false;  // NaN determinant: culled (kernel.cu:301-305)
}

// forward outputs for one Gaussian.  areas exactly as inverseCov2D (IEEE sqrt, *3, ceil).
template <int K3>
GSB_HD void forward_one(const float pw[3], const float q[4], const float s[3], const float *sh /*[K3][3]*/,
                        const Cam &cam, float u[2], float conic[3], float col[3], float *depth, int area[2]) {
  Geo g;
  geometry(pw, q, s, cam, g);
  // colour is not gated by depth in the reference (kernel.cu:619-726)
  float Y[K3], dY[K3][3];
  float x = 0.f, y = 0.f, z = 0.f;
  if (K3 > 1) {
    const float d0 = pw[0] - cam.twc[0], d1 = pw[1] - cam.twc[1], d2 = pw[2] - cam.twc[2];
    const float ninv = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    x = d0 * ninv; y = d1 * ninv; z = d2 * ninv;
  }
  sh_basis<K3>(x, y, z, Y, dY);
  col[0] = col[1] = col[2] = 0.5f;
  for (int l = 0; l < K3; l++)
    for (int c = 0; c < 3; c++) col[c] = fmaf(Y[l], sh[3 * l + c], col[c]);
  if (!g.keep) {
    u[0] = u[1] = 0.f;
    conic[0] = conic[1] = conic[2] = 0.f;
    area[0] = area[1] = 0;
    *depth = -1.0f;
    return;
  }
  u[0] = fmaf(g.pc[0] * cam.fx, g.zi, cam.cx);
  u[1] = fmaf(g.pc[1] * cam.fy, g.zi, cam.cy);
  conic[0] = g.di * g.c2[2];
  conic[1] = -g.di * g.c2[1];
  conic[2] = g.di * g.c2[0];
  area[0] = (int)ceilf(3.0f * sqrtf(fabsf(g.c2[0])));
  area[1] = (int)ceilf(3.0f * sqrtf(fabsf(g.c2[2])));
  *depth = g.pc[2];
}

// backward for one Gaussian.  gu[2], gconic[3], gcol[3] are the splatB outputs for it.
// gsh may alias sh (row-private in-place update).
template <int K3>
GSB_HD void backward_one(const float pw[3], const float q[4], const float s[3], const float *sh,
                         const Cam &cam, const float gu[2], const float gconic[3], const float gcol[3],
                         float gpw[3], float gq[4], float gs[3], float *gsh /*[K3][3]*/) {
  // ---- colour branch (independent of the depth cull, like sh2Color)
  float Y[K3], dY[K3][3];
  float d[3] = {0.f, 0.f, 1.f}, ninv = 1.f, r[3] = {0.f, 0.f, 0.f};
  if (K3 > 1) {
    d[0] = pw[0] - cam.twc[0]; d[1] = pw[1] - cam.twc[1]; d[2] = pw[2] - cam.twc[2];
    ninv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    r[0] = d[0] * ninv; r[1] = d[1] * ninv; r[2] = d[2] * ninv;
  }
  sh_basis<K3>(r[0], r[1], r[2], Y, dY);
  float gdir[3] = {0.f, 0.f, 0.f};
  for (int l = 0; l < K3; l++) {
    const float t = gcol[0] * sh[3 * l] + gcol[1] * sh[3 * l + 1] + gcol[2] * sh[3 * l + 2];
    if (l > 0) {
      gdir[0] = fmaf(t, dY[l][0], gdir[0]);
      gdir[1] = fmaf(t, dY[l][1], gdir[1]);
      gdir[2] = fmaf(t, dY[l][2], gdir[2]);
    }
    gsh[3 * l] = gcol[0] * Y[l];
    gsh[3 * l + 1] = gcol[1] * Y[l];
    gsh[3 * l + 2] = gcol[2] * Y[l];
  }
  gpw[0] = gpw[1] = gpw[2] = 0.f;
  if (K3 > 1) {
    const float rg = r[0] * gdir[0] + r[1] * gdir[1] + r[2] * gdir[2];
    for (int a = 0; a < 3; a++) gpw[a] = (gdir[a] - r[a] * rg) * ninv;
  }
  gq[0] = gq[1] = gq[2] = gq[3] = 0.f;
  gs[0] = gs[1] = gs[2] = 0.f;

  // ---- geometry branch
  Geo g;
  geometry(pw, q, s, cam, g);
  const bool any = (gu[0] != 0.f) || (gu[1] != 0.f) || (gconic[0] != 0.f) || (gconic[1] != 0.f) || (gconic[2] != 0.f);
  if (!g.keep || !any) return;
  const float a = g.c2[0], b = g.c2[1], c = g.c2[2], di = g.di, d2 = di * di;
  // dL/dcov2d = dL/dconic . dconic/dcov2d   (B.5.3)
  const float ga = gconic[0] * (-c * c * d2) + gconic[1] * (b * c * d2) + gconic[2] * (di - a * c * d2);
  const float gb = gconic[0] * (2.f * b * c * d2) + gconic[1] * (-di - 2.f * b * b * d2) + gconic[2] * (2.f * a * b * d2);
  const float gc = gconic[0] * (di - a * c * d2) + gconic[1] * (a * b * d2) + gconic[2] * (-a * a * d2);
  // G = [[ga, gb/2], [gb/2, gc]]
  const float G00 = ga, G01 = 0.5f * gb, G11 = gc;
  const float *M = g.M;
  // GM (2x3)
  float GM[6];
  for (int k = 0; k < 3; k++) {
    GM[k] = G00 * M[k] + G01 * M[3 + k];
    GM[3 + k] = G01 * M[k] + G11 * M[3 + k];
  }
  // dL/dSigma_ij = (2 - delta_ij) (M^T G M)_ij
  float g3[6];
  {
    int n = 0;
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) {
        const float v = M[i] * GM[j] + M[3 + i] * GM[3 + j];
        g3[n++] = (i == j) ? v : 2.f * v;
      }
  }
  // dL/dM = 2 G (M Sigma);  dL/dJ = dL/dM Rcw^T
  float dM[6];
  for (int k = 0; k < 3; k++) {
    dM[k] = 2.f * (G00 * g.MS[k] + G01 * g.MS[3 + k]);
    dM[3 + k] = 2.f * (G01 * g.MS[k] + G11 * g.MS[3 + k]);
  }
  const float *R = cam.R;
  float dJ[6];
  for (int a2 = 0; a2 < 2; a2++)
    for (int k = 0; k < 3; k++)
      dJ[3 * a2 + k] = dM[3 * a2] * R[3 * k] + dM[3 * a2 + 1] * R[3 * k + 1] + dM[3 * a2 + 2] * R[3 * k + 2];
  const float zi = g.zi, zi2 = zi * zi, zi3 = zi2 * zi;
  float gpc[3];
  gpc[0] = dJ[2] * (-cam.fx * zi2);
  gpc[1] = dJ[5] * (-cam.fy * zi2);
  gpc[2] = dJ[0] * (-cam.fx * zi2) + dJ[2] * (2.f * cam.fx * g.xc * zi3) + dJ[4] * (-cam.fy * zi2) +
           dJ[5] * (2.f * cam.fy * g.yc * zi3);
  // projection u = (fx x/z + cx, fy y/z + cy)   (B.1.2, unclamped x, y)
  gpc[0] += gu[0] * (cam.fx * zi);
  gpc[1] += gu[1] * (cam.fy * zi);
  gpc[2] += -(gu[0] * (g.pc[0] * cam.fx) + gu[1] * (g.pc[1] * cam.fy)) * zi2;
  for (int j = 0; j < 3; j++) gpw[j] += gpc[0] * R[j] + gpc[1] * R[3 + j] + gpc[2] * R[6 + j];
  // cov3d: dL/dM3 = 2 sym(g3) M3
  const float G3[9] = {g3[0], 0.5f * g3[1], 0.5f * g3[2], 0.5f * g3[1], g3[3], 0.5f * g3[4],
                       0.5f * g3[2], 0.5f * g3[4], g3[5]};
  float dM3[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      dM3[3 * i + j] = 2.f * (G3[3 * i] * g.M3[j] + G3[3 * i + 1] * g.M3[3 + j] + G3[3 * i + 2] * g.M3[6 + j]);
  for (int j = 0; j < 3; j++) gs[j] = g.R3[j] * dM3[j] + g.R3[3 + j] * dM3[3 + j] + g.R3[6 + j] * dM3[6 + j];
  float dR[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) dR[3 * i + j] = dM3[3 * i + j] * s[j];
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  // dR/dq (halved; the factor 2 is applied at the end)
  gq[0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
  gq[1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - w * dR[5] + z * dR[6] + w * dR[7] - 2.f * x * dR[8]);
  gq[2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + w * dR[2] + x * dR[3] + z * dR[5] - w * dR[6] + z * dR[7] - 2.f * y * dR[8]);
  gq[3] = 2.f * (-2.f * z * dR[0] - w * dR[1] + x * dR[2] + w * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
}

}  // namespace pg
}  // namespace gsb

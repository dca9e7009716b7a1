// K3: batched PnP, one warp per frame (stands in for cv::solvePnP
// SOLVEPNP_ITERATIVE at detect_pose.py:509-526 and for the reprojection gate of
// transform_helper.py:98-121).
//
// Minimises sum_i |pi(K, dist, R(r) X_i + t) - u_i|^2 over (r,t) with
// Levenberg-Marquardt in float64: analytic Rodrigues derivative, 27 normal-
// equation sums reduced with warp shuffles, 6x6 Cholesky solve on the device.
// Frames without an extrinsic guess start from a normalised DLT (smallest
// eigenvector of the 12x12 A^T A by inverse iteration, then polar
// orthonormalisation), the same initialisation family OpenCV uses for
// non-planar point sets.  OpenCV lands on the least-squares minimiser (SURVEY.md
// section 6: <1e-9 rad from scipy), so parity is "converge to the minimiser".
// The epilogue reproduces the reference's mean L2 reprojection error in float32.
#include "agt_common.cuh"

namespace {

constexpr int PNP_WARPS = 4;
constexpr int PNP_MAX_ITERS = 60;

struct Proj {
  double u, v;          // projected pixel
  double j[2][3];       // d(u,v)/d(Xc)
  bool ok;
};

__device__ __forceinline__ Proj project_point(const agt_camera& cam, double X, double Y, double Z) {
  Proj p;
  p.ok = Z > 1e-9;
  double iz = p.ok ? agt_rcp_newton(Z) : 1.0;
  double x = X * iz, y = Y * iz;
  double dxx = 1.0, dxy = 0.0, dyx = 0.0, dyy = 1.0, xd = x, yd = y;
  if (cam.has_dist) {
    double r2 = x * x + y * y;
    double rad = 1.0 + r2 * (cam.k1 + r2 * (cam.k2 + r2 * cam.k3));
    double drad = cam.k1 + r2 * (2.0 * cam.k2 + r2 * 3.0 * cam.k3);   // d rad / d r2
    xd = x * rad + 2.0 * cam.p1 * x * y + cam.p2 * (r2 + 2.0 * x * x);
    yd = y * rad + cam.p1 * (r2 + 2.0 * y * y) + 2.0 * cam.p2 * x * y;
    dxx = rad + x * drad * 2.0 * x + 2.0 * cam.p1 * y + cam.p2 * 6.0 * x;
    dxy = x * drad * 2.0 * y + 2.0 * cam.p1 * x + cam.p2 * 2.0 * y;
    dyx = y * drad * 2.0 * x + cam.p1 * 2.0 * x + 2.0 * cam.p2 * y;
    dyy = rad + y * drad * 2.0 * y + cam.p1 * 6.0 * y + 2.0 * cam.p2 * x;
  }
  p.u = cam.fx * xd + cam.cx;
  p.v = cam.fy * yd + cam.cy;
  // d(x,y)/d(X,Y,Z) = [[iz,0,-x iz],[0,iz,-y iz]]
  p.j[0][0] = cam.fx * dxx * iz; p.j[0][1] = cam.fx * dxy * iz; p.j[0][2] = -cam.fx * (dxx * x + dxy * y) * iz;
  p.j[1][0] = cam.fy * dyx * iz; p.j[1][1] = cam.fy * dyy * iz; p.j[1][2] = -cam.fy * (dyx * x + dyy * y) * iz;
  return p;
}

// R(r) and dR/dr_i for i=0..2 (row-major 3x3 each): dR/dr_i = [a_i]x R with a_i = (r_i r + r x (I-R) e_i) / theta^2.
// One warp runs one frame, so a launch lasts as long as its longest chain of dependent float64 operations: theta comes
// from a Newton-refined reciprocal square root (no sqrt, no division: 1/theta and 1/theta^2 fall out of it), the skew
// products are written out (no multiplications by the zeros of [a]x).
__device__ __forceinline__ void rodrigues_with_jacobian(const double r[3], double R[9], double dR[3][9]) {
  const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (th2 < 1e-20) {                                       // R = I + [r]x, dR/dr_i = [e_i]x
    R[0] = 1; R[1] = -r[2]; R[2] = r[1];
    R[3] = r[2]; R[4] = 1; R[5] = -r[0];
    R[6] = -r[1]; R[7] = r[0]; R[8] = 1;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 9; ++k) dR[i][k] = 0.0;
    dR[0][5] = -1; dR[0][7] = 1; dR[1][2] = 1; dR[1][6] = -1; dR[2][1] = -1; dR[2][3] = 1;
    return;
  }
  const double ith = agt_rsqrt_newton(th2), th = th2 * ith, ith2 = ith * ith;
  const double kx = r[0] * ith, ky = r[1] * ith, kz = r[2] * ith;
  double sn, c;
  sincos(th, &sn, &c);
  const double c1 = 1.0 - c;
  R[0] = c + c1 * kx * kx;       R[1] = c1 * kx * ky - sn * kz; R[2] = c1 * kx * kz + sn * ky;
  R[3] = c1 * ky * kx + sn * kz; R[4] = c + c1 * ky * ky;       R[5] = c1 * ky * kz - sn * kx;
  R[6] = c1 * kz * kx - sn * ky; R[7] = c1 * kz * ky + sn * kx; R[8] = c + c1 * kz * kz;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // v = (I - R) e_i
    const double v[3] = {(i == 0) - R[0 + i], (i == 1) - R[3 + i], (i == 2) - R[6 + i]};
    const double ri = r[i] * ith2;
    const double a0 = fma(ri, r[0], (r[1] * v[2] - r[2] * v[1]) * ith2);
    const double a1 = fma(ri, r[1], (r[2] * v[0] - r[0] * v[2]) * ith2);
    const double a2 = fma(ri, r[2], (r[0] * v[1] - r[1] * v[0]) * ith2);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      dR[i][0 + cc] = a1 * R[6 + cc] - a2 * R[3 + cc];
      dR[i][3 + cc] = a2 * R[0 + cc] - a0 * R[6 + cc];
      dR[i][6 + cc] = a0 * R[3 + cc] - a1 * R[0 + cc];
    }
  }
}

struct Normal {
  double H[21];
  double g[6];
  double c;
};

// Evaluate cost and normal equations at pose p for this lane's (<= 2) points and reduce over the warp.
__device__ inline void evaluate(const agt_camera& cam, const double p[6], const double X[2][3], const double U[2][2],
                                const bool have[2], Normal& out, bool& all_in_front, float nrm[2]) {
  double R[9], dR[3][9];
  rodrigues_with_jacobian(p, R, dR);
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  bool front = true;
  nrm[0] = nrm[1] = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!have[s]) continue;
    const double* x = X[s];
    double Xc = R[0] * x[0] + R[1] * x[1] + R[2] * x[2] + p[3];
    double Yc = R[3] * x[0] + R[4] * x[1] + R[5] * x[2] + p[4];
    double Zc = R[6] * x[0] + R[7] * x[1] + R[8] * x[2] + p[5];
    Proj pr = project_point(cam, Xc, Yc, Zc);
    front = front && pr.ok;
    double e[2] = {pr.u - U[s][0], pr.v - U[s][1]};
    {   // this point's term of the reference's mean reprojection error (transform_helper.py:106-119), in float32 like the reference
      const float du = __fsub_rn((float)U[s][0], (float)pr.u), dv = __fsub_rn((float)U[s][1], (float)pr.v);
      nrm[s] = __fsqrt_rn(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
    }
    double J[2][6];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double d0 = dR[i][0] * x[0] + dR[i][1] * x[1] + dR[i][2] * x[2];
      double d1 = dR[i][3] * x[0] + dR[i][4] * x[1] + dR[i][5] * x[2];
      double d2 = dR[i][6] * x[0] + dR[i][7] * x[1] + dR[i][8] * x[2];
      J[0][i] = pr.j[0][0] * d0 + pr.j[0][1] * d1 + pr.j[0][2] * d2;
      J[1][i] = pr.j[1][0] * d0 + pr.j[1][1] * d1 + pr.j[1][2] * d2;
      J[0][3 + i] = pr.j[0][i];
      J[1][3 + i] = pr.j[1][i];
    }
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) acc[k++] += J[0][a] * J[0][b] + J[1][a] * J[1][b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += J[0][a] * e[0] + J[1][a] * e[1];
    acc[27] += e[0] * e[0] + e[1] * e[1];
  }
  // One transposing butterfly (31 exchanges instead of 28 x 5) leaves the total of sum k on lane k, then every lane
  // fetches the 28 totals: it solves the 6x6 system itself.
  {
    const int lane = threadIdx.x & 31;
    double v[32];
#pragma unroll
    for (int k = 0; k < 28; ++k) v[k] = acc[k];
#pragma unroll
    for (int k = 28; k < 32; ++k) v[k] = 0.0;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
      const bool up = (lane & h) != 0;
#pragma unroll
      for (int k = 0; k < h; ++k) {
        const double keep = up ? v[k + h] : v[k], send = up ? v[k] : v[k + h];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, h);
      }
    }
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = __shfl_sync(0xffffffffu, v[0], k);
  }
#pragma unroll
  for (int k = 0; k < 21; ++k) out.H[k] = acc[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) out.g[k] = acc[21 + k];
  out.c = acc[27];
  all_in_front = __all_sync(0xffffffffu, front);
}

// Smallest eigenvector of the symmetric N x N matrix M (PSD, row stride N, N <= 12) by inverse iteration
// on M + eps I (Cholesky).  Executed by one lane.
__device__ inline bool smallest_eigvec(double* M, double* v, const int N) {
  double tr = 0.0;
  for (int i = 0; i < N; ++i) tr += M[i * N + i];
  double eps = 1e-13 * tr + 1e-300;
  for (int i = 0; i < N; ++i) M[i * N + i] += eps;
  for (int j = 0; j < N; ++j) {
    double d = M[j * N + j];
    for (int k = 0; k < j; ++k) d -= M[j * N + k] * M[j * N + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    M[j * N + j] = d;
    for (int i = j + 1; i < N; ++i) {
      double s = M[i * N + j];
      for (int k = 0; k < j; ++k) s -= M[i * N + k] * M[j * N + k];
      M[i * N + j] = s / d;
    }
  }
  for (int i = 0; i < N; ++i) v[i] = 1.0 / sqrt(12.0) * ((i * 7 + 3) % 5 + 1);   // generic start vector
  for (int it = 0; it < 12; ++it) {
    for (int i = 0; i < N; ++i) {
      double s = v[i];
      for (int k = 0; k < i; ++k) s -= M[i * N + k] * v[k];
      v[i] = s / M[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
      double s = v[i];
      for (int k = i + 1; k < N; ++k) s -= M[k * N + i] * v[k];
      v[i] = s / M[i * N + i];
    }
    double n = 0.0;
    for (int i = 0; i < N; ++i) n += v[i] * v[i];
    n = 1.0 / sqrt(n);
    for (int i = 0; i < N; ++i) v[i] *= n;
  }
  return true;
}

// Eigen-decomposition of a symmetric 3x3 matrix (cyclic Jacobi): eigenvalues w ascending, eigenvectors as the ROWS of V.
__device__ inline void eig_sym3(const double C[6], double w[3], double V[9]) {
  double a[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
  double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};          // columns are eigenvectors
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
    if (off < 1e-300 || off < 1e-17 * (fabs(a[0]) + fabs(a[4]) + fabs(a[8]))) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      if (fabs(a[p * 3 + q]) < 1e-300) continue;
      double th = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * a[p * 3 + q]);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      for (int k = 0; k < 3; ++k) {                       // A <- A J
        double akp = a[k * 3 + p], akq = a[k * 3 + q];
        a[k * 3 + p] = c * akp - sn * akq; a[k * 3 + q] = sn * akp + c * akq;
      }
      for (int k = 0; k < 3; ++k) {                       // A <- J^T A
        double apk = a[p * 3 + k], aqk = a[q * 3 + k];
        a[p * 3 + k] = c * apk - sn * aqk; a[q * 3 + k] = sn * apk + c * aqk;
      }
      for (int k = 0; k < 3; ++k) {
        double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
        v[k * 3 + p] = c * vkp - sn * vkq; v[k * 3 + q] = sn * vkp + c * vkq;
      }
    }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a[0], a[4], a[8]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (d[idx[j]] > d[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  for (int i = 0; i < 3; ++i) {
    w[i] = d[idx[i]];
    for (int k = 0; k < 3; ++k) V[i * 3 + k] = v[k * 3 + idx[i]];
  }
}

// cv::undistortPoints for one normalised point: fixed-point iteration of the Brown-Conrady model (as OpenCV, 5 rounds suffice
// for an initial value; the LM that follows works on the distorted pixels)
__device__ inline void undistort_normalised(const agt_camera& cam, double& x, double& y) {
  if (!cam.has_dist) return;
  const double x0 = x, y0 = y;
  for (int it = 0; it < 8; ++it) {
    const double r2 = x * x + y * y;
    const double icd = 1.0 / (1.0 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2);
    const double dx = 2.0 * cam.p1 * x * y + cam.p2 * (r2 + 2.0 * x * x), dy = cam.p1 * (r2 + 2.0 * y * y) + 2.0 * cam.p2 * x * y;
    x = (x0 - dx) * icd; y = (y0 - dy) * icd;
  }
}

__device__ inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// R <- nearest rotation (polar factor) by Newton iteration R <- (R + R^-T)/2.
__device__ inline void orthonormalise(double* R) {
  for (int it = 0; it < 12; ++it) {
    double d = det3(R);
    if (fabs(d) < 1e-300) return;
    double id = 1.0 / d;
    // inverse transpose = cofactor matrix / det
    double C[9] = {(R[4] * R[8] - R[5] * R[7]) * id, (R[5] * R[6] - R[3] * R[8]) * id, (R[3] * R[7] - R[4] * R[6]) * id,
                   (R[2] * R[7] - R[1] * R[8]) * id, (R[0] * R[8] - R[2] * R[6]) * id, (R[1] * R[6] - R[0] * R[7]) * id,
                   (R[1] * R[5] - R[2] * R[4]) * id, (R[2] * R[3] - R[0] * R[5]) * id, (R[0] * R[4] - R[1] * R[3]) * id};
    double diff = 0.0;
    for (int k = 0; k < 9; ++k) {
      double n = 0.5 * (R[k] + C[k]);
      diff += fabs(n - R[k]);
      R[k] = n;
    }
    if (diff < 1e-15) break;
  }
}

// Initial pose of a PLANAR point set from the homography plane -> normalised image (the planar branch of OpenCV's
// cvFindExtrinsicCameraParams2): plane frame from the scatter's eigenvectors (rows of eV: the two in-plane axes are rows 2 and
// 1, the normal row 0), Hartley-normalised DLT for H (9x9 normal matrix, smallest eigenvector), [r1 r2 t] = H up to scale,
// r3 = r1 x r2, polar orthonormalisation, and back to the object frame.  The LM that follows only needs the right basin.
__device__ inline bool planar_init(const agt_camera& cam, const double X[2][3], const double xn[2][2], const bool have[2], int n,
                                   const double sum[5], const double eV[9], double* M, int lane, double p[6]) {
  double e1[3] = {eV[6], eV[7], eV[8]}, e2[3] = {eV[3], eV[4], eV[5]};
  double nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  double q[2][2], so = 0.0, si = 0.0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const double a = X[s][0] - sum[0], b = X[s][1] - sum[1], c = X[s][2] - sum[2];
    q[s][0] = e1[0] * a + e1[1] * b + e1[2] * c; q[s][1] = e2[0] * a + e2[1] * b + e2[2] * c;
    if (have[s]) {
      so += q[s][0] * q[s][0] + q[s][1] * q[s][1];
      const double d = xn[s][0] - sum[3], e = xn[s][1] - sum[4];
      si += d * d + e * e;
    }
  }
  so = agt_warp_sum(so); si = agt_warp_sum(si);
  if (!(so > 0.0) || !(si > 0.0)) return false;
  const double sco = sqrt(2.0 * n / so), sci = sqrt(2.0 * n / si);
  double acc[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) acc[k] = 0.0;
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (have[s]) {
      const double x = sco * q[s][0], y = sco * q[s][1], u = sci * (xn[s][0] - sum[3]), v = sci * (xn[s][1] - sum[4]);
      const double r0[9] = {x, y, 1.0, 0.0, 0.0, 0.0, -u * x, -u * y, -u}, r1[9] = {0.0, 0.0, 0.0, x, y, 1.0, -v * x, -v * y, -v};
      int k = 0;
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = a; b < 9; ++b) acc[k++] += r0[a] * r0[b] + r1[a] * r1[b];
    }
#pragma unroll
  for (int k = 0; k < 45; ++k) acc[k] = agt_warp_sum(acc[k]);
  __syncwarp();
  double h[9];
  bool eig_ok = true;
  if (lane == 0) {
    int k = 0;
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b) { M[a * 9 + b] = M[b * 9 + a] = acc[k]; ++k; }
    eig_ok = smallest_eigvec(M, h, 9);
  }
  eig_ok = __shfl_sync(0xffffffffu, (int)eig_ok, 0) != 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = __shfl_sync(0xffffffffu, h[k], 0);
  __syncwarp();
  if (!eig_ok) return false;
  // denormalise: H = Ti^-1 H' To, To = diag(sco, sco, 1), Ti^-1 = [[1/sci, 0, mu], [0, 1/sci, mv], [0, 0, 1]]
  double H[9];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double sc = c < 2 ? sco : 1.0;
    H[0 + c] = (h[0 + c] / sci + sum[3] * h[6 + c]) * sc;
    H[3 + c] = (h[3 + c] / sci + sum[4] * h[6 + c]) * sc;
    H[6 + c] = h[6 + c] * sc;
  }
  if (H[8] < 0.0) {                                       // the plane's origin (the centroid) lies in front of the camera
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = -H[k];
  }
  const double n1 = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
  if (!(n1 > 0.0) || !(n2 > 0.0)) return false;
  const double lam = 2.0 / (n1 + n2);
  const double r1[3] = {H[0] / n1, H[3] / n1, H[6] / n1}, r2[3] = {H[1] / n2, H[4] / n2, H[7] / n2};
  const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  double Rh[9] = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]};       // columns r1 r2 r3
  orthonormalise(Rh);
  const double th[3] = {H[2] * lam, H[5] * lam, H[8] * lam};
  // X_cam = Rh [e1; e2; n] (X - mean) + th
  double R[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = Rh[r * 3 + 0] * e1[c] + Rh[r * 3 + 1] * e2[c] + Rh[r * 3 + 2] * nn[c];
  agt_log_rotation(R, p);
#pragma unroll
  for (int r = 0; r < 3; ++r) p[3 + r] = th[r] - (R[r * 3 + 0] * sum[0] + R[r * 3 + 1] * sum[1] + R[r * 3 + 2] * sum[2]);
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) fin = fin && isfinite(p[k]);
  return fin;
}

// The stream pipeline's steps around the PnP of a frame, done by the frame's own warp in the same launch (agt_streams_front; a
// frame-step is a chain of dependent launches, and five of them were one-warp-per-frame bookkeeping):
//   before: agt_lk_merge (re-admit the tags whose four corners LK tracked, detect_pose.py stage 2) and agt_ape_prepare (the
//           extrinsic guess of the stream's state record, detect_pose.py:508)
//   after:  agt_accept_gate (detect_pose.py:494, 533, 539) and the reset of the refinement status of the frame
struct PnpFront {
  const float* tracked; const uint8_t* status; const uint8_t* prev_valid;       // LK results; tracked == nullptr: no merge
  float* img; uint8_t* valid;                                                   // the frame's corners, merged in place
  const int32_t* n_tags_in; int32_t* n_tags_out; int32_t* tracked_tags;
  const double* state; int enhance;
  uint8_t* gate; uint8_t* refine_status;                                        // nullable
};

template <bool kFront>
__global__ void __launch_bounds__(PNP_WARPS * 32)
pnp_kernel(agt_camera cam, const float* __restrict__ obj, const float* __restrict__ img_in, const uint8_t* __restrict__ valid_in,
           const double* __restrict__ guess, const uint8_t* __restrict__ use_guess, double* __restrict__ pose_out,
           uint8_t* __restrict__ ok_out, float* __restrict__ err_out, int32_t* __restrict__ iters_out, int n_pts, int batch,
           PnpFront F) {
  __shared__ double s_M[PNP_WARPS][144];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.x * PNP_WARPS + wid;
  if (f >= batch) return;

  // (the merged corners are read back through the pointers they were written through, not through a read-only path)
  const float* img = kFront ? F.img : img_in;
  const uint8_t* valid = kFront ? F.valid : valid_in;
  int n_tags = 0;
  if (kFront) {
    n_tags = F.n_tags_in[f];
    if (F.tracked != nullptr) {                            // agt_lk_merge: lane = tag
      const int tags = n_pts >> 2, before = n_tags;
      bool det = false;
      if (lane < tags) {
        const int64_t base = ((int64_t)f * tags + lane) * 4;
        det = F.valid[base] && F.valid[base + 1] && F.valid[base + 2] && F.valid[base + 3];
        if (!det && before < 2) {
          bool ok4 = true;
          for (int j = 0; j < 4; ++j) ok4 = ok4 && F.prev_valid[base + j] != 0 && F.status[base + j] == 1;
          if (ok4) {
            for (int j = 0; j < 4; ++j) {
              F.img[(base + j) * 2] = F.tracked[(base + j) * 2];
              F.img[(base + j) * 2 + 1] = F.tracked[(base + j) * 2 + 1];
              F.valid[base + j] = 1;
            }
            det = true;
          }
        }
      }
      n_tags = __popc(__ballot_sync(0xffffffffu, det));
      if (lane == 0 && F.tracked_tags != nullptr) F.tracked_tags[f] = n_tags - before;
      __syncwarp();
    }
    if (lane == 0) F.n_tags_out[f] = n_tags;
  }

  double X[2][3], U[2][2];
  bool have[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int p = lane + 32 * s;
    have[s] = p < n_pts && (valid == nullptr || valid[(int64_t)f * n_pts + p] != 0);
    if (have[s]) {
      X[s][0] = obj[p * 3]; X[s][1] = obj[p * 3 + 1]; X[s][2] = obj[p * 3 + 2];
      U[s][0] = img[((int64_t)f * n_pts + p) * 2]; U[s][1] = img[((int64_t)f * n_pts + p) * 2 + 1];
      bool fin = isfinite(U[s][0]) && isfinite(U[s][1]);
      have[s] = fin;
    }
    if (!have[s]) { X[s][0] = X[s][1] = X[s][2] = 0.0; U[s][0] = U[s][1] = 0.0; }
  }
  int n = __popc(__ballot_sync(0xffffffffu, have[0])) + __popc(__ballot_sync(0xffffffffu, have[1]));

  double p[6];
  bool ok = n >= 4;
  // (front: agt_ape_prepare - the guess is the one of the stream's state record, detect_pose.py:508)
  const double* gsrc = kFront ? F.state + (int64_t)f * AGT_STREAM_STATE_DOUBLES + AGT_STATE_GUESS : guess + (int64_t)f * 6;
  bool guessed = kFront ? (F.enhance != 0 && F.state[(int64_t)f * AGT_STREAM_STATE_DOUBLES + AGT_STATE_HAS_GUESS] != 0.0)
                        : (use_guess != nullptr && guess != nullptr && use_guess[f] != 0);
  if (guessed) {
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = gsrc[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) ok = ok && isfinite(p[k]);
  } else if (ok) {
    // ---------------- initial value without a guess (cv::solvePnP ITERATIVE: cvFindExtrinsicCameraParams2) --------------
    // normalised, undistorted image points; planar object points (smallest / middle eigenvalue of their scatter < 1e-3, the
    // rule OpenCV uses) start from a homography (>= 4 points: a single tag, a planar tag board), the others from a DLT (>= 6)
    double sum[5] = {0, 0, 0, 0, 0};
    double xn[2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      xn[s][0] = (U[s][0] - cam.cx) / cam.fx; xn[s][1] = (U[s][1] - cam.cy) / cam.fy;
      undistort_normalised(cam, xn[s][0], xn[s][1]);
      if (have[s]) { sum[0] += X[s][0]; sum[1] += X[s][1]; sum[2] += X[s][2]; sum[3] += xn[s][0]; sum[4] += xn[s][1]; }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) sum[k] = agt_warp_sum(sum[k]) / (n > 0 ? n : 1);
    double C6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (have[s]) {
        const double a = X[s][0] - sum[0], b = X[s][1] - sum[1], c = X[s][2] - sum[2];
        C6[0] += a * a; C6[1] += a * b; C6[2] += a * c; C6[3] += b * b; C6[4] += b * c; C6[5] += c * c;
      }
#pragma unroll
    for (int k = 0; k < 6; ++k) C6[k] = agt_warp_sum(C6[k]);
    double ew[3], eV[9];
    eig_sym3(C6, ew, eV);                                 // every lane: same inputs, same result
    const bool planar = !(ew[0] >= 1e-3 * ew[1]);
    if (planar) {
      ok = ok && ew[1] > 0.0 && planar_init(cam, X, xn, have, n, sum, eV, s_M[wid], lane, p);
    } else {
    ok = n >= 6;
    double so = 0.0, si = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (have[s]) {
        double a = X[s][0] - sum[0], b = X[s][1] - sum[1], c = X[s][2] - sum[2];
        so += a * a + b * b + c * c;
        double d = xn[s][0] - sum[3], e = xn[s][1] - sum[4];
        si += d * d + e * e;
      }
    so = agt_warp_sum(so); si = agt_warp_sum(si);
    ok = ok && so > 0.0 && si > 0.0;
    double sco = ok ? sqrt(3.0 * n / so) : 1.0, sci = ok ? sqrt(2.0 * n / si) : 1.0;
    // 4 symmetric 4x4 moment matrices: S, Sx, Sy, Sq
    double mom[40];
#pragma unroll
    for (int k = 0; k < 40; ++k) mom[k] = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (have[s]) {
        double h[4] = {sco * (X[s][0] - sum[0]), sco * (X[s][1] - sum[1]), sco * (X[s][2] - sum[2]), 1.0};
        double x = sci * (xn[s][0] - sum[3]), y = sci * (xn[s][1] - sum[4]);
        double q = x * x + y * y;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = a; b < 4; ++b) {
            double hh = h[a] * h[b];
            mom[k] += hh; mom[10 + k] += x * hh; mom[20 + k] += y * hh; mom[30 + k] += q * hh;
            ++k;
          }
      }
#pragma unroll
    for (int k = 0; k < 40; ++k) mom[k] = agt_warp_sum(mom[k]);
    double* M = s_M[wid];
    __syncwarp();
    if (lane == 0) {
      for (int k = 0; k < 144; ++k) M[k] = 0.0;
      int k = 0;
      for (int a = 0; a < 4; ++a)
        for (int b = a; b < 4; ++b) {
          double s0 = mom[k], sx = mom[10 + k], sy = mom[20 + k], sq = mom[30 + k];
          ++k;
          // blocks: (0,0)=S (1,1)=S (0,2)=-Sx (1,2)=-Sy (2,2)=Sq, symmetric
          M[(a)*12 + b] = M[(b)*12 + a] = s0;
          M[(4 + a) * 12 + 4 + b] = M[(4 + b) * 12 + 4 + a] = s0;
          M[(a)*12 + 8 + b] = M[(b)*12 + 8 + a] = -sx;
          M[(8 + b) * 12 + a] = M[(8 + a) * 12 + b] = -sx;
          M[(4 + a) * 12 + 8 + b] = M[(4 + b) * 12 + 8 + a] = -sy;
          M[(8 + b) * 12 + 4 + a] = M[(8 + a) * 12 + 4 + b] = -sy;
          M[(8 + a) * 12 + 8 + b] = M[(8 + b) * 12 + 8 + a] = sq;
        }
    }
    __syncwarp();
    double v[12];
    bool eig_ok = true;
    if (lane == 0) eig_ok = smallest_eigvec(M, v, 12);
    eig_ok = __shfl_sync(0xffffffffu, (int)eig_ok, 0) != 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = __shfl_sync(0xffffffffu, v[k], 0);
    ok = ok && eig_ok;
    // denormalise: P = Ti^-1 P' To ; Ti^-1 = [[1/sci,0,mx],[0,1/sci,my],[0,0,1]] ; To = [[sco I, -sco c],[0,1]]
    double Pn[12];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      Pn[0 + c] = v[0 + c] / sci + sum[3] * v[8 + c];
      Pn[4 + c] = v[4 + c] / sci + sum[4] * v[8 + c];
      Pn[8 + c] = v[8 + c];
    }
    double Rm[9], tt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Rm[r * 3 + 0] = Pn[r * 4 + 0] * sco; Rm[r * 3 + 1] = Pn[r * 4 + 1] * sco; Rm[r * 3 + 2] = Pn[r * 4 + 2] * sco;
      tt[r] = Pn[r * 4 + 3] - sco * (Pn[r * 4 + 0] * sum[0] + Pn[r * 4 + 1] * sum[1] + Pn[r * 4 + 2] * sum[2]);
    }
    double d = det3(Rm);
    ok = ok && fabs(d) > 1e-300 && isfinite(d);
    double sc = ok ? (d < 0 ? -1.0 : 1.0) / cbrt(fabs(d)) : 1.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] *= sc;
#pragma unroll
    for (int k = 0; k < 3; ++k) tt[k] *= sc;
    orthonormalise(Rm);
    agt_log_rotation(Rm, p);
    p[3] = tt[0]; p[4] = tt[1]; p[5] = tt[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) ok = ok && isfinite(p[k]);
    }
  }

  // ---------------- Levenberg-Marquardt ---------------------------------------------
  int iters = 0;
  float nrm[2] = {0.f, 0.f};            // per-point reprojection error of the pose in p (set by the evaluation that produced it)
  if (ok) {
    // One evaluation site: the loop body is "evaluate the trial pose, decide, solve for the next trial" (the first trial is the
    // initial value and is always taken) - the evaluation is the bulk of the kernel's code and is instantiated once.
    Normal cur, tri;
    bool front;
    float tnrm[2];
    double lam = 1e-3, dmax = 0.0;
    double q[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) q[a] = p[a];
    bool first = true;
    while (true) {
      evaluate(cam, q, X, U, have, tri, front, tnrm);
      bool stop = false;
      if (first) {
        cur = tri;
        nrm[0] = tnrm[0]; nrm[1] = tnrm[1];
        first = false;
      } else {
        if (front && tri.c < cur.c) {
#pragma unroll
          for (int a = 0; a < 6; ++a) p[a] = q[a];
          cur = tri;
          nrm[0] = tnrm[0]; nrm[1] = tnrm[1];
          lam = fmax(lam * 0.1, 1e-15);
          // an accepted step below 1e-9 (rad / m) ends the loop: the steps shrink at least by the damping factor (<= 1e-3, a tenth of
          // it per accepted step), so what is left is below 1e-12 - four orders of magnitude under the distance to cv::solvePnP,
          // which itself stops at a relative change of 1.2e-7.  (1e-11 bought one more evaluation per frame and nothing else.)
          stop = dmax < 1e-9;
        } else {
          lam *= 10.0;
          // a rejected step below 1e-8: the gradient is rounding noise, p is the minimiser to float64 accuracy.  (The bound was
          // 1e-10: one frame in eight then spent six or seven more evaluations raising the damping until a 1e-9 step had shrunk
          // below it - and a 64-frame launch lasts as long as its slowest frame: 10-11 evaluations instead of 5.)
          stop = dmax < 1e-8 || lam > 1e12;
        }
        ++iters;
      }
      if (stop || iters >= PNP_MAX_ITERS) break;
      // next trial: damped normal equations, solved in registers (packed upper triangle); a failed factorisation raises the damping
      double b[6];
      while (true) {
        double A[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) A[k] = cur.H[k];
#pragma unroll
        for (int a = 0; a < 6; ++a) { A[agt_hk(a, a)] *= (1.0 + lam); b[a] = -cur.g[a]; }
        if (agt_chol6_packed(A, b)) break;
        lam *= 10.0;
        ++iters;
        if (lam > 1e12) { ok = false; break; }
        if (iters >= PNP_MAX_ITERS) break;
      }
      if (!ok || iters >= PNP_MAX_ITERS) break;
      dmax = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) { q[a] = p[a] + b[a]; dmax = fmax(dmax, fabs(b[a])); }
    }
    // keep the rotation vector in OpenCV's range (angle <= pi)
    if (p[0] * p[0] + p[1] * p[1] + p[2] * p[2] > 3.14159265358979323846 * 3.14159265358979323846) {
      double R[9];
      agt_rodrigues(p, R);
      agt_log_rotation(R, p);
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) ok = ok && isfinite(p[a]);
  }

  // ---------------- epilogue: mean reprojection error as transform_helper.py:106-119 ----
  // (the per-point terms come from the evaluation of the final pose: same projection, no second pass)
  if (!ok) nrm[0] = nrm[1] = 0.f;
  float total = 0.f;
  const unsigned hm0 = __ballot_sync(0xffffffffu, have[0]), hm1 = __ballot_sync(0xffffffffu, have[1]);
  for (int q = 0; q < n_pts; ++q) {                     // point order, one rounding per addition (Python's sum())
    const float v0 = __shfl_sync(0xffffffffu, q < 32 ? nrm[0] : nrm[1], q & 31);
    if (((q < 32 ? hm0 : hm1) >> (q & 31)) & 1u) total = __fadd_rn(total, v0);
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 6; ++a) pose_out[(int64_t)f * 6 + a] = ok ? p[a] : 0.0;
    ok_out[f] = ok ? 1 : 0;
    const float err = ok && n > 0 ? __fdiv_rn(total, (float)n) : 0.f;
    err_out[f] = err;
    if (iters_out) iters_out[f] = iters;
    if (kFront) {
      if (F.gate != nullptr) F.gate[f] = ok && err < 2.0f && n_tags >= 2 ? 1 : 0;      // agt_accept_gate
      if (F.refine_status != nullptr) F.refine_status[f] = 0;                         // masked frames keep status 0
    }
  }
}

__global__ void project_kernel(agt_camera cam, const float* __restrict__ obj, const double* __restrict__ pose,
                               double* __restrict__ out, int n_pts, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t f = i / n_pts;
  int p = (int)(i - f * n_pts);
  double R[9];
  const double* ps = pose + f * 6;
  double r[3] = {ps[0], ps[1], ps[2]};
  agt_rodrigues(r, R);
  double x = obj[p * 3], y = obj[p * 3 + 1], z = obj[p * 3 + 2];
  double Xc = R[0] * x + R[1] * y + R[2] * z + ps[3];
  double Yc = R[3] * x + R[4] * y + R[5] * z + ps[4];
  double Zc = R[6] * x + R[7] * y + R[8] * z + ps[5];
  // cv::projectPoints divides by z without a cheirality test
  double iz = Zc != 0.0 ? 1.0 / Zc : 1.0;
  Proj pr = project_point(cam, Xc * iz, Yc * iz, 1.0);
  out[i * 2] = pr.u;
  out[i * 2 + 1] = pr.v;
}

}  // namespace

extern "C" int agt_pnp(agt_ctx* ctx, const float* d_obj_pts, const float* d_img_pts, const uint8_t* d_valid,
                       const double* d_guess, const uint8_t* d_use_guess, double* d_pose, uint8_t* d_ok,
                       float* d_reproj_err, int32_t* d_iters, int batch, int n_pts) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_pnp: call agt_set_camera first");
  if (!d_obj_pts || !d_img_pts || !d_pose || !d_ok || !d_reproj_err || batch < 0)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_pnp: null argument or negative batch");
  if (n_pts < 1 || n_pts > AGT_MAX_POINTS) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_pnp: n_pts must be 1..%d", AGT_MAX_POINTS);
  if (batch == 0) return AGT_OK;
  int blocks = (batch + PNP_WARPS - 1) / PNP_WARPS;
  pnp_kernel<false><<<blocks, PNP_WARPS * 32, 0, ctx->stream>>>(ctx->cam, d_obj_pts, d_img_pts, d_valid, d_guess, d_use_guess,
                                                                d_pose, d_ok, d_reproj_err, d_iters, n_pts, batch, PnpFront{});
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_streams_front(agt_ctx* ctx, const float* d_obj_pts, const float* d_tracked_pts, const uint8_t* d_lk_status,
                                 const uint8_t* d_prev_valid, float* d_img_pts, uint8_t* d_valid, const int32_t* d_n_tags_in,
                                 int32_t* d_n_tags, int32_t* d_tracked_tags, const double* d_state, int enhance_ape, double* d_pose,
                                 uint8_t* d_ok, float* d_reproj_err, int32_t* d_iters, uint8_t* d_gate, uint8_t* d_refine_status, int batch,
                                 int n_pts) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;
  if (!ctx->camera_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_streams_front: call agt_set_camera first");
  if (!d_obj_pts || !d_img_pts || !d_valid || !d_n_tags_in || !d_n_tags || !d_state || !d_pose || !d_ok || !d_reproj_err || batch < 0 ||
      n_pts < 4 || (n_pts & 3) || n_pts > AGT_MAX_POINTS || (d_tracked_pts != nullptr && (!d_lk_status || !d_prev_valid)))
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_streams_front: bad arguments");
  PnpFront F;
  F.tracked = d_tracked_pts; F.status = d_lk_status; F.prev_valid = d_prev_valid; F.img = d_img_pts; F.valid = d_valid;
  F.n_tags_in = d_n_tags_in; F.n_tags_out = d_n_tags; F.tracked_tags = d_tracked_tags; F.state = d_state; F.enhance = enhance_ape;
  F.gate = d_gate; F.refine_status = d_refine_status;
  int blocks = (batch + PNP_WARPS - 1) / PNP_WARPS;
  pnp_kernel<true><<<blocks, PNP_WARPS * 32, 0, ctx->stream>>>(ctx->cam, d_obj_pts, nullptr, nullptr, nullptr, nullptr, d_pose, d_ok,
                                                               d_reproj_err, d_iters, n_pts, batch, F);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_project(agt_ctx* ctx, const float* d_obj_pts, const double* d_pose, double* d_out, int batch, int n_pts) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_project: call agt_set_camera first");
  if (!d_obj_pts || !d_pose || !d_out || batch < 0 || n_pts < 0) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_project: bad arguments");
  int64_t total = (int64_t)batch * n_pts;
  if (total == 0) return AGT_OK;
  project_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(ctx->cam, d_obj_pts, d_pose, d_out, n_pts, total);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

// N3, first step: sub-pixel refinement of tag corners around predicted positions, one warp per corner.
//
// The reference's detector is the un-vendored swatbotics apriltag C library (detect_pose.py:86-95, :368-371; its
// refine_edges option snaps quad corners to the image gradient); the library is absent from the image, so the frozen
// semantics are OpenCV's, whose ArUco module detects the same tag36h11 family: cv::cornerSubPix(gray, corners, (win, win),
// (-1, -1), (COUNT + EPS, max_iters, eps)), restated in oracle/corner_oracle.py (bit-identical to cv2 there).
//
// Per iteration the warp resamples the (2 win + 3)^2 float patch around the current estimate (cv::getRectSubPix: bilinear,
// replicated border, x fraction >= 1e-4) into shared memory, forms the central differences and the Gaussian-weighted normal
// equations of "the gradient is orthogonal to the vector to the corner" over the (2 win + 1)^2 window in float64 (OpenCV
// accumulates in double as well; the order of the additions differs, which moves the result by ~1e-6 px), reduces the five
// sums by shuffles and takes the 2x2 step.  The point is kept when it moves further than the window.
#include "agt_common.cuh"

namespace {

constexpr int MAX_WIN = 7;
constexpr int MAX_PATCH = 2 * MAX_WIN + 3;       // 17
constexpr int CS_WARPS = 4;

__global__ void __launch_bounds__(CS_WARPS * 32)
corner_subpix_kernel(const uint8_t* __restrict__ img, int w, int h, int64_t pitch, int64_t stride, const float* __restrict__ pts,
                     const uint8_t* __restrict__ valid, const uint8_t* __restrict__ win_of, float* __restrict__ out, int n_pts, int64_t total,
                     int win, int max_iters, double eps2) {
  __shared__ float s_patch[CS_WARPS][MAX_PATCH][MAX_PATCH + 1];
  __shared__ float s_masks[CS_WARPS][2 * MAX_WIN + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t gid = (int64_t)blockIdx.x * CS_WARPS + wid;
  if (gid >= total) return;
  if (win_of != nullptr) win = min(max((int)win_of[gid], 1), MAX_WIN);       // per-point window (the detector sizes it by the tag's cells)
  const int n = 2 * win + 1, np = n + 2;
  float* s_mask = s_masks[wid];
  if (lane < n) {
    const float x = (float)(lane - win) / (float)win;
    s_mask[lane] = expf(-x * x);             // std::exp(float): correctly rounded here (checked against cv2 in the tests)
  }
  __syncwarp();
  const float tx = pts[gid * 2], ty = pts[gid * 2 + 1];
  if ((valid != nullptr && valid[gid] == 0) || !(tx >= 0.f && tx < (float)w && ty >= 0.f && ty < (float)h)) {
    if (lane == 0) { out[gid * 2] = tx; out[gid * 2 + 1] = ty; }
    return;
  }
  const uint8_t* f = img + (gid / n_pts) * stride;
  float (*P)[MAX_PATCH + 1] = s_patch[wid];
  float cx = tx, cy = ty;
  for (int it = 0; it < max_iters; ++it) {
    // ---- getRectSubPix: (2 win + 3)^2 bilinear samples around (cx, cy)
    const float x0 = __fsub_rn(cx, (float)(np - 1) * 0.5f), y0 = __fsub_rn(cy, (float)(np - 1) * 0.5f);
    const int ix = (int)floorf(x0), iy = (int)floorf(y0);
    const float a = fmaxf(__fsub_rn(x0, (float)ix), 0.0001f), b = __fsub_rn(y0, (float)iy);
    const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    for (int i = lane; i < np * np; i += 32) {
      const int r = i / np, c = i - r * np;
      const int xa = min(max(ix + c, 0), w - 1), xb = min(max(ix + c + 1, 0), w - 1);
      const int ya = min(max(iy + r, 0), h - 1), yb = min(max(iy + r + 1, 0), h - 1);
      const float p00 = (float)__ldg(f + (int64_t)ya * pitch + xa), p01 = (float)__ldg(f + (int64_t)ya * pitch + xb);
      const float p10 = (float)__ldg(f + (int64_t)yb * pitch + xa), p11 = (float)__ldg(f + (int64_t)yb * pitch + xb);
      const float left = __fadd_rn(__fmul_rn(p00, nb), __fmul_rn(p10, b)), right = __fadd_rn(__fmul_rn(p01, nb), __fmul_rn(p11, b));
      P[r][c] = __fadd_rn(__fmul_rn(left, na), __fmul_rn(right, a));
    }
    __syncwarp();
    // ---- Gaussian-weighted normal equations over the window, float64
    double sa = 0.0, sb = 0.0, sc = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < n * n; i += 32) {
      const int r = i / n, c = i - r * n;
      const double m = (double)__fmul_rn(s_mask[r], s_mask[c]);
      const double gx = (double)P[r + 1][c + 2] - (double)P[r + 1][c], gy = (double)P[r + 2][c + 1] - (double)P[r][c + 1];
      const double gxx = gx * gx * m, gxy = gx * gy * m, gyy = gy * gy * m;
      const double px = (double)(c - win), py = (double)(r - win);
      sa += gxx; sb += gxy; sc += gyy;
      s1 += gxx * px + gxy * py;
      s2 += gxy * px + gyy * py;
    }
    __syncwarp();
    sa = agt_warp_sum(sa); sb = agt_warp_sum(sb); sc = agt_warp_sum(sc); s1 = agt_warp_sum(s1); s2 = agt_warp_sum(s2);
    const double det = sa * sc - sb * sb;
    if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
    const double sl = 1.0 / det;
    const float nx = (float)((double)cx + sc * sl * s1 - sb * sl * s2), ny = (float)((double)cy - sb * sl * s1 + sa * sl * s2);
    const double dx = (double)nx - (double)cx, dy = (double)ny - (double)cy;
    cx = nx; cy = ny;
    if (!(cx >= 0.f && cx < (float)w && cy >= 0.f && cy < (float)h)) break;
    if (!(dx * dx + dy * dy > eps2)) break;
  }
  if (!(fabsf(cx - tx) <= (float)win && fabsf(cy - ty) <= (float)win)) { cx = tx; cy = ty; }     // poor convergence: keep the input
  if (lane == 0) { out[gid * 2] = cx; out[gid * 2 + 1] = cy; }
}

}  // namespace

// d_win [batch][n_pts] (nullable): window half-size per point instead of `win`
int agt_corner_subpix_windows(agt_ctx* ctx, const uint8_t* d_gray, int w, int h, int64_t pitch, int64_t stride, const float* d_pts,
                              const uint8_t* d_valid, const uint8_t* d_win, float* d_out, int batch, int n_pts, int win, int max_iters,
                              double eps) {
  if (!ctx) return AGT_ERR_INVALID;
  if ((int64_t)batch * n_pts == 0) return AGT_OK;
  if (!d_gray || !d_pts || !d_out || batch < 0 || n_pts < 0 || w < 1 || h < 1 || pitch < w || win < 1 || win > MAX_WIN || max_iters < 1 ||
      !(eps >= 0.0))
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_corner_subpix: bad arguments (window half-size 1..%d)", MAX_WIN);
  if (max_iters > 100) max_iters = 100;                 // cv::cornerSubPix clamps the count the same way
  const int64_t total = (int64_t)batch * n_pts;
  const int64_t blocks = (total + CS_WARPS - 1) / CS_WARPS;
  if (blocks > 0x7fffffffLL) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_corner_subpix: batch too large");
  corner_subpix_kernel<<<(unsigned)blocks, CS_WARPS * 32, 0, ctx->stream>>>(d_gray, w, h, pitch, stride, d_pts, d_valid, d_win, d_out, n_pts,
                                                                            total, win, max_iters, eps * eps);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_corner_subpix(agt_ctx* ctx, const uint8_t* d_gray, int w, int h, int64_t pitch, int64_t stride, const float* d_pts,
                                 const uint8_t* d_valid, float* d_out, int batch, int n_pts, int win, int max_iters, double eps) {
  return agt_corner_subpix_windows(ctx, d_gray, w, h, pitch, stride, d_pts, d_valid, nullptr, d_out, batch, n_pts, win, max_iters, eps);
}

extern "C" int agt_corner_subpix_host(agt_ctx* ctx, const uint8_t* h_gray, int w, int h, float* h_pts, int n_pts, int win, int max_iters,
                                      double eps) {
  if (!ctx) return AGT_ERR_INVALID;
  if (n_pts == 0) return AGT_OK;
  if (!h_gray || !h_pts || w < 1 || h < 1 || n_pts < 0) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_corner_subpix_host: bad arguments");
  AGT_CUDA(ctx, cudaSetDevice(ctx->device));
  uint8_t *dimg, *dp;
  int rc;
  const size_t img_bytes = (size_t)w * h, pts_bytes = sizeof(float) * 2 * (size_t)n_pts;
  if ((rc = agt_scratch(ctx, 0, img_bytes, reinterpret_cast<void**>(&dimg)))) return rc;
  if ((rc = agt_scratch(ctx, 7, 2 * pts_bytes + 64, reinterpret_cast<void**>(&dp)))) return rc;
  cudaStream_t st = ctx->stream;
  AGT_CUDA(ctx, cudaMemcpyAsync(dimg, h_gray, img_bytes, cudaMemcpyHostToDevice, st));
  AGT_CUDA(ctx, cudaMemcpyAsync(dp, h_pts, pts_bytes, cudaMemcpyHostToDevice, st));
  float* dout = reinterpret_cast<float*>(dp + ((pts_bytes + 63) & ~(size_t)63));
  if ((rc = agt_corner_subpix(ctx, dimg, w, h, w, (int64_t)img_bytes, reinterpret_cast<float*>(dp), nullptr, dout, 1, n_pts, win, max_iters, eps)))
    return rc;
  AGT_CUDA(ctx, cudaMemcpyAsync(h_pts, dout, pts_bytes, cudaMemcpyDeviceToHost, st));
  AGT_CUDA(ctx, cudaStreamSynchronize(st));
  return AGT_OK;
}

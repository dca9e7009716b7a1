// Frame ingest with lens undistortion: cv::undistort + crop + BGR2GRAY in one pass,
// bit-exact for 8-bit frames, written straight into level 0 of a pyramid.
//
// The reference undistorts every frame on the CPU before anything else looks at it
// (detect_pose.py:147-183 undistort_frame: cv.getOptimalNewCameraMatrix, cv.undistort, crop to the ROI;
// called from process_frame, detect_pose.py:611-619) and converts it to gray later (detect_pose.py:602).
// What cv::undistort computes is restated in oracle/undistort_oracle.py and pinned to the installed cv2 there:
// per output pixel the source coordinate through the inverse new camera matrix and the Brown-Conrady model in
// float64, rounded to 1/32 px; bilinear blend with cv::remap's table of int16 weights, (sum + 2^14) >> 15,
// BORDER_CONSTANT 0; then the 15-bit BGR2GRAY of the three blended channels.
//
// One thread per output pixel; the map of a pixel is the same for every frame of the batch, so a thread computes
// it once (float64, ~40 operations) and applies it to kFramesPerThread frames: the pass is bound by reading the
// BGR frame once (6.2 MB per 1080p frame) and writing the gray one.
#include <cmath>
#include <vector>

#include "agt_common.cuh"

namespace {

constexpr int kFramesPerThread = 8;

__global__ void __launch_bounds__(128)
undistort_gray_kernel(agt_camera cam, agt_undistort U, const short* __restrict__ tab, const uint8_t* __restrict__ src, int w, int h,
                      int channels, int64_t spitch, int64_t sstride, uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride,
                      int batch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= U.roi_w) return;
  // ---- the map of this pixel (cv::initUndistortRectifyMap on the stripe of cv::undistort that holds the row) ----
  const int j = U.roi_x + x, Y = U.roi_y + y;
  const int y0 = (Y / U.stripe) * U.stripe, i = Y - y0;
  const double cy = U.ncy - (double)y0;                       // principal point of the stripe's new camera matrix
  // inverse of the upper-triangular new camera matrix [[fx s cx] [0 fy cy] [0 0 1]]
  const double inv_ab = 1.0 / (U.nfx * U.nfy);
  const double ir0 = 1.0 / U.nfx, ir1 = -U.nskew * inv_ab, ir2 = (U.nskew * cy - U.ncx * U.nfy) * inv_ab;
  const double ir4 = 1.0 / U.nfy, ir5 = -cy / U.nfy;
  const double px = (double)i * ir1 + ir2 + (double)j * ir0;
  const double py = (double)i * ir4 + ir5;
  const double x2 = px * px, y2 = py * py, r2 = x2 + y2, _2xy = 2.0 * px * py;
  const double kr = 1.0 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2;
  const double xd = px * kr + cam.p1 * _2xy + cam.p2 * (r2 + 2.0 * x2);
  const double yd = py * kr + cam.p1 * (r2 + 2.0 * y2) + cam.p2 * _2xy;
  const int iu = __double2int_rn((cam.fx * xd + cam.cx) * 32.0), iv = __double2int_rn((cam.fy * yd + cam.cy) * 32.0);
  const int sx = (int)(short)(iu >> 5), sy = (int)(short)(iv >> 5);       // CV_16SC2 integer part (the cast wraps)
  const short4 wt = *reinterpret_cast<const short4*>(tab + 4 * ((iv & 31) * 32 + (iu & 31)));
  // taps outside the frame read the constant border 0: zero their weights instead of branching per load
  const bool x0ok = (unsigned)sx < (unsigned)w, x1ok = (unsigned)(sx + 1) < (unsigned)w;
  const bool y0ok = (unsigned)sy < (unsigned)h, y1ok = (unsigned)(sy + 1) < (unsigned)h;
  const int w00 = x0ok && y0ok ? wt.x : 0, w01 = x1ok && y0ok ? wt.y : 0, w10 = x0ok && y1ok ? wt.z : 0, w11 = x1ok && y1ok ? wt.w : 0;
  const int cx0 = min(max(sx, 0), w - 1), cx1 = min(max(sx + 1, 0), w - 1), cy0 = min(max(sy, 0), h - 1), cy1 = min(max(sy + 1, 0), h - 1);
  const int64_t o00 = (int64_t)cy0 * spitch + (int64_t)cx0 * channels, o01 = (int64_t)cy0 * spitch + (int64_t)cx1 * channels;
  const int64_t o10 = (int64_t)cy1 * spitch + (int64_t)cx0 * channels, o11 = (int64_t)cy1 * spitch + (int64_t)cx1 * channels;
  const int b0 = blockIdx.z * kFramesPerThread, b1 = min(b0 + kFramesPerThread, batch);
  if (channels == 3 && cx1 == cx0 + 1 && ((spitch | sstride) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    // interior pixel of a BGR frame: the two taps of a row are six consecutive bytes B0 G0 R0 B1 G1 R1.  Three aligned
    // word loads + two funnel shifts fetch them (half the cache wavefronts of six byte loads), one permute per channel
    // puts (p0, p1) side by side and dp2a blends them with the pair of 16-bit weights.
    const uint32_t wt = (uint32_t)w00 | ((uint32_t)w01 << 16), wb = (uint32_t)w10 | ((uint32_t)w11 << 16);
    const int64_t a0 = o00 & ~(int64_t)3, a1 = o10 & ~(int64_t)3;
    const uint32_t sh0 = (uint32_t)(o00 & 3) * 8u, sh1 = (uint32_t)(o10 & 3) * 8u;
    // the third word may lie past the end of the last row of the last frame: clamp it inside the batch
    const int64_t last_word = ((int64_t)(batch - 1) * sstride + (int64_t)h * spitch - 4) & ~(int64_t)3;
#pragma unroll 1
    for (int b = b0; b < b1; ++b) {
      const int64_t fo = (int64_t)b * sstride;
      const uint32_t* t = reinterpret_cast<const uint32_t*>(src + fo + a0);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(src + fo + a1);
      const bool t2ok = fo + a0 + 8 <= last_word, u2ok = fo + a1 + 8 <= last_word;
      const uint32_t t0 = __ldg(t), t1 = __ldg(t + 1), t2 = t2ok ? __ldg(t + 2) : 0u;
      const uint32_t u0 = __ldg(u), u1 = __ldg(u + 1), u2 = u2ok ? __ldg(u + 2) : 0u;
      const uint32_t tl = __funnelshift_r(t0, t1, sh0), th = __funnelshift_r(t1, t2, sh0);     // B0 G0 R0 B1 | G1 R1 . .
      const uint32_t ul = __funnelshift_r(u0, u1, sh1), uh = __funnelshift_r(u1, u2, sh1);
      const uint32_t vb = __dp2a_lo(wb, __byte_perm(ul, uh, 0x0030), __dp2a_lo(wt, __byte_perm(tl, th, 0x0030), 1u << 14)) >> 15;
      const uint32_t vg = __dp2a_lo(wb, __byte_perm(ul, uh, 0x0041), __dp2a_lo(wt, __byte_perm(tl, th, 0x0041), 1u << 14)) >> 15;
      const uint32_t vr = __dp2a_lo(wb, __byte_perm(ul, uh, 0x0052), __dp2a_lo(wt, __byte_perm(tl, th, 0x0052), 1u << 14)) >> 15;
      dst[(int64_t)b * dstride + (int64_t)y * dpitch + x] = (uint8_t)((vb * 3735u + vg * 19235u + vr * 9798u + 16384u) >> 15);
    }
    return;
  }
  for (int b = b0; b < b1; ++b) {
    const uint8_t* f = src + (int64_t)b * sstride;
    int v[3];
    for (int c = 0; c < channels; ++c)
      v[c] = (w00 * __ldg(f + o00 + c) + w01 * __ldg(f + o01 + c) + w10 * __ldg(f + o10 + c) + w11 * __ldg(f + o11 + c) + (1 << 14)) >> 15;
    const int gray = channels == 3 ? (v[0] * 3735 + v[1] * 19235 + v[2] * 9798 + 16384) >> 15 : v[0];
    dst[(int64_t)b * dstride + (int64_t)y * dpitch + x] = (uint8_t)gray;
  }
}

// cv::initInterTab2D(INTER_LINEAR, fixed point): float32 products scaled by 2^15, saturated to int16, the rounding
// remainder folded into one weight by a loop that also looks at the (not yet written, zero) start of the next entry
void build_remap_table(std::vector<short>& out) {
  std::vector<int> flat(1024 * 4 + 8, 0);
  float t1[32][2];
  for (int i = 0; i < 32; ++i) { float x = (float)i * (1.f / 32.f); t1[i][0] = 1.f - x; t1[i][1] = x; }
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      const int base = (i * 32 + j) * 4;
      int isum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const float v = t1[i][k1] * t1[j][k2];
          int q = (int)nearbyint((double)v * 32768.0);
          q = q < -32768 ? -32768 : (q > 32767 ? 32767 : q);
          flat[base + k1 * 2 + k2] = q;
          isum += q;
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int big = base + 3, small = base + 3;
        for (int k1 = 1; k1 <= 2; ++k1)
          for (int k2 = 1; k2 <= 2; ++k2) {
            const int idx = base + k1 * 2 + k2;
            if (flat[idx] < flat[small]) small = idx;
            else if (flat[idx] > flat[big]) big = idx;
          }
        if (diff < 0) flat[big] -= diff; else flat[small] -= diff;
      }
    }
  out.resize(1024 * 4);
  for (int k = 0; k < 1024 * 4; ++k) out[k] = (short)flat[k];
}

}  // namespace

extern "C" int agt_set_undistort(agt_ctx* ctx, const double* new_K, int width, int height, int roi_x, int roi_y, int roi_w, int roi_h) {
  if (!ctx) return AGT_ERR_INVALID;
  if (!ctx->camera_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_set_undistort: call agt_set_camera first");
  if (!new_K || width < 1 || height < 1 || roi_x < 0 || roi_y < 0 || roi_w < 1 || roi_h < 1 || roi_x + roi_w > width || roi_y + roi_h > height)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_set_undistort: bad arguments");
  if (!(new_K[0] != 0.0) || !(new_K[4] != 0.0) || new_K[3] != 0.0 || new_K[6] != 0.0 || new_K[7] != 0.0 || new_K[8] != 1.0)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_set_undistort: the new camera matrix must be [[fx s cx] [0 fy cy] [0 0 1]]");
  AGT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->d_remap_tab) {
    std::vector<short> tab;
    build_remap_table(tab);
    AGT_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_remap_tab), tab.size() * sizeof(short)));
    AGT_CUDA(ctx, cudaMemcpy(ctx->d_remap_tab, tab.data(), tab.size() * sizeof(short), cudaMemcpyHostToDevice));
  }
  agt_undistort& U = ctx->und;
  U.nfx = new_K[0]; U.nskew = new_K[1]; U.ncx = new_K[2]; U.nfy = new_K[4]; U.ncy = new_K[5];
  int stripe = (1 << 12) / width;
  if (stripe < 1) stripe = 1;
  if (stripe > height) stripe = height;
  U.stripe = stripe; U.width = width; U.height = height;
  U.roi_x = roi_x; U.roi_y = roi_y; U.roi_w = roi_w; U.roi_h = roi_h;
  ctx->undistort_set = 1;
  return AGT_OK;
}

extern "C" int agt_undistort_to_gray(agt_ctx* ctx, const uint8_t* d_src, int w, int h, int channels, int64_t src_pitch,
                                     int64_t src_stride, uint8_t* d_gray, int64_t dst_pitch, int64_t dst_stride, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set || !ctx->undistort_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_undistort_to_gray: call agt_set_camera and agt_set_undistort first");
  if (!d_src || !d_gray || batch < 0 || (channels != 1 && channels != 3) || src_pitch < (int64_t)w * channels || dst_pitch < ctx->und.roi_w)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_to_gray: bad arguments");
  if (w != ctx->und.width || h != ctx->und.height)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_to_gray: frame is %dx%d but agt_set_undistort was given %dx%d", w, h, ctx->und.width, ctx->und.height);
  agt_camera cam = ctx->cam;
  if (!cam.has_dist) { cam.k1 = cam.k2 = cam.p1 = cam.p2 = cam.k3 = 0.0; }
  const int groups = (batch + kFramesPerThread - 1) / kFramesPerThread;
  for (int g0 = 0; g0 < groups; g0 += 65535) {
    const int ng = groups - g0 < 65535 ? groups - g0 : 65535;
    const int64_t f0 = (int64_t)g0 * kFramesPerThread;
    dim3 grid((ctx->und.roi_w + 127) / 128, ctx->und.roi_h, ng);
    undistort_gray_kernel<<<grid, 128, 0, ctx->stream>>>(cam, ctx->und, ctx->d_remap_tab, d_src + f0 * src_stride, w, h, channels, src_pitch,
                                                         src_stride, d_gray + f0 * dst_stride, dst_pitch, dst_stride,
                                                         (int)(batch - f0));
    AGT_LAUNCH_CHECK(ctx);
  }
  return AGT_OK;
}

extern "C" int agt_undistort_to_gray_host(agt_ctx* ctx, const uint8_t* h_src, int w, int h, int channels, uint8_t* h_gray) {
  if (!ctx) return AGT_ERR_INVALID;
  if (!ctx->undistort_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_undistort_to_gray_host: call agt_set_undistort first");
  if (!h_src || !h_gray || w < 1 || h < 1 || (channels != 1 && channels != 3)) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_to_gray_host: bad arguments");
  AGT_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t in_bytes = (size_t)w * h * channels, out_bytes = (size_t)ctx->und.roi_w * ctx->und.roi_h;
  uint8_t *din, *dout;
  int rc;
  if ((rc = agt_scratch(ctx, 0, in_bytes, reinterpret_cast<void**>(&din)))) return rc;
  if ((rc = agt_scratch(ctx, 1, out_bytes, reinterpret_cast<void**>(&dout)))) return rc;
  AGT_CUDA(ctx, cudaMemcpyAsync(din, h_src, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = agt_undistort_to_gray(ctx, din, w, h, channels, (int64_t)w * channels, (int64_t)in_bytes, dout, ctx->und.roi_w, (int64_t)out_bytes, 1)))
    return rc;
  AGT_CUDA(ctx, cudaMemcpyAsync(h_gray, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AGT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGT_OK;
}

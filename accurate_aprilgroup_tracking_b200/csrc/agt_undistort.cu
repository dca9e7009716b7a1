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
// The map of a pixel is the same for every frame of the batch, so a thread computes it once (float64, ~40 operations)
// and applies it to a group of frames; the pass is bound by reading the BGR frame once (6.2 MB per 1080p frame) and
// writing the gray one.
//
// undistort_gray_tiled_kernel (16-byte aligned frames: every real camera format): a CTA owns a 32x32 output tile.  The
// source pixels the tile reads form a small bounding box (the lens map is smooth), found once by a block-wide min/max over
// the pixel maps; per frame that box is staged into shared memory with 16-byte asynchronous copies, three frames deep, so
// that tens of KB per SM are in flight without holding registers, and the four taps of every pixel are gathered from
// shared memory.  Tiles whose box does not fit a stage (extreme distortion) and unaligned frames take the per-pixel
// global-memory path (undistort_gray_kernel), which computes the same integers.
#include <cmath>
#include <vector>

#include "agt_common.cuh"

namespace {

constexpr int kFramesPerThread = 8;

// The map of output pixel (x, y) of the crop (cv::initUndistortRectifyMap on the stripe of cv::undistort that holds the row):
// integer source position (sx, sy) and the four int16 bilinear weights of cv::remap's table.
__device__ __forceinline__ void undistort_map(const agt_camera& cam, const agt_undistort& U, const short* __restrict__ tab, int x, int y,
                                              int& sx, int& sy, short4& wt) {
  const int j = U.roi_x + x, Y = U.roi_y + y;
  const int y0 = (Y / U.stripe) * U.stripe, i = Y - y0;
  const double cy = U.ncy - (double)y0;                       // principal point of the stripe's new camera matrix
  // inverse of the upper-triangular new camera matrix [[fx s cx] [0 fy cy] [0 0 1]]
  // (the stripe-independent entries are correctly rounded quotients / one product: computed once on the host)
  const double inv_ab = U.inv_ab, ir0 = U.ir0, ir1 = U.ir1, ir4 = U.ir4;
  const double ir2 = (U.nskew * cy - U.ncx * U.nfy) * inv_ab, ir5 = -cy / U.nfy;
  const double px = (double)i * ir1 + ir2 + (double)j * ir0;
  const double py = (double)i * ir4 + ir5;
  const double x2 = px * px, y2 = py * py, r2 = x2 + y2, _2xy = 2.0 * px * py;
  const double kr = 1.0 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2;
  const double xd = px * kr + cam.p1 * _2xy + cam.p2 * (r2 + 2.0 * x2);
  const double yd = py * kr + cam.p1 * (r2 + 2.0 * y2) + cam.p2 * _2xy;
  const int iu = __double2int_rn((cam.fx * xd + cam.cx) * 32.0), iv = __double2int_rn((cam.fy * yd + cam.cy) * 32.0);
  sx = (int)(short)(iu >> 5); sy = (int)(short)(iv >> 5);       // CV_16SC2 integer part (the cast wraps)
  wt = *reinterpret_cast<const short4*>(tab + 4 * ((iv & 31) * 32 + (iu & 31)));
}

__global__ void __launch_bounds__(128)
undistort_gray_kernel(agt_camera cam, agt_undistort U, const short* __restrict__ tab, const uint8_t* __restrict__ src, int w, int h,
                      int channels, int64_t spitch, int64_t sstride, uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride,
                      int batch, int keep_channels) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= U.roi_w) return;
  int sx, sy;
  short4 wt;
  undistort_map(cam, U, tab, x, y, sx, sy, wt);
  // taps outside the frame read the constant border 0: zero their weights instead of branching per load
  const bool x0ok = (unsigned)sx < (unsigned)w, x1ok = (unsigned)(sx + 1) < (unsigned)w;
  const bool y0ok = (unsigned)sy < (unsigned)h, y1ok = (unsigned)(sy + 1) < (unsigned)h;
  const int w00 = x0ok && y0ok ? wt.x : 0, w01 = x1ok && y0ok ? wt.y : 0, w10 = x0ok && y1ok ? wt.z : 0, w11 = x1ok && y1ok ? wt.w : 0;
  const int cx0 = min(max(sx, 0), w - 1), cx1 = min(max(sx + 1, 0), w - 1), cy0 = min(max(sy, 0), h - 1), cy1 = min(max(sy + 1, 0), h - 1);
  const int64_t o00 = (int64_t)cy0 * spitch + (int64_t)cx0 * channels, o01 = (int64_t)cy0 * spitch + (int64_t)cx1 * channels;
  const int64_t o10 = (int64_t)cy1 * spitch + (int64_t)cx0 * channels, o11 = (int64_t)cy1 * spitch + (int64_t)cx1 * channels;
  const int b0 = blockIdx.z * kFramesPerThread, b1 = min(b0 + kFramesPerThread, batch);
  if (keep_channels) {
    // cv::undistort alone (the frame process_frame hands back for display, detect_pose.py:611-619): the blended channels as they are
    for (int b = b0; b < b1; ++b) {
      const uint8_t* f = src + (int64_t)b * sstride;
      for (int c = 0; c < channels; ++c)
        dst[(int64_t)b * dstride + (int64_t)y * dpitch + (int64_t)x * channels + c] =
            (uint8_t)((w00 * __ldg(f + o00 + c) + w01 * __ldg(f + o01 + c) + w10 * __ldg(f + o10 + c) + w11 * __ldg(f + o11 + c) + (1 << 14)) >> 15);
    }
    return;
  }
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

// ---------------------------------------------------------------------------------------------------------------
// Tiled variant: see the header comment.  256 threads, a 32x32 output tile, 4 pixels per thread (a warp = 32
// consecutive pixels of one row, rows ty, ty+8, ty+16, ty+24), UT_GROUP frames per CTA.
// ---------------------------------------------------------------------------------------------------------------
// (8 pixels per thread - 32x64 tiles, 14 KB stages - amortise the per-frame overhead over more pixels but measured slower:
// 2.65e5 frames/s at 64 registers with spills, 3.10e5 at 80 registers and 3 CTAs per SM, against 3.33e5 at 32 frames per CTA)
constexpr int UT_TILE = 32, UT_THREADS = 256, UT_PX = 4, UT_TILE_H = 8 * UT_PX, UT_STAGES = 3;
constexpr int UT_STAGE_BYTES = 8192;
constexpr int UT_CHUNKS = (UT_STAGE_BYTES / 16 + UT_THREADS - 1) / UT_THREADS;      // 16-byte requests per thread and frame

// Every pixel takes the same branch-free path: the top-left tap is clamped into the frame, the tap to its right is the
// next pixel in memory and the tap row below is `dy` rows further (0 on the last row); taps that fall outside the frame
// have weight 0 (BORDER_CONSTANT 0), and for sx == -1 (only the right-hand taps are inside) the weights move to the
// left-hand slots, which then sit on column 0.  Pixels outside the crop keep weight 0 and are not stored.
template <int CH>
__global__ void __launch_bounds__(UT_THREADS, 4)
undistort_gray_tiled_kernel(agt_camera cam, agt_undistort U, const short* __restrict__ tab, const uint8_t* __restrict__ src, int w, int h,
                            int64_t spitch, int64_t sstride, uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride, int batch,
                            int group) {
  __shared__ __align__(16) uint8_t s_buf[UT_STAGES][UT_STAGE_BYTES];
  __shared__ int s_box[4];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int x = blockIdx.x * UT_TILE + tx;
  if (tid == 0) { s_box[0] = 0x7fffffff; s_box[1] = 0x7fffffff; s_box[2] = -1; s_box[3] = -1; }
  __syncthreads();
  // ---- the maps of this thread's pixels: packed weight pairs, clamped top-left tap, row step ------------------------
  uint32_t wtop[UT_PX], wbot[UT_PX];
  int tap[UT_PX];                          // cx0 | cy0 << 14 | dy << 28 | live << 29   (frames up to 16384 x 16384)
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -1, by1 = -1;
#pragma unroll
  for (int k = 0; k < UT_PX; ++k) {
    const int y = blockIdx.y * UT_TILE_H + ty + 8 * k;
    tap[k] = 0; wtop[k] = 0; wbot[k] = 0;
    if (x < U.roi_w && y < U.roi_h) {
      int sx, sy;
      short4 wt;
      undistort_map(cam, U, tab, x, y, sx, sy, wt);
      const bool x0ok = (unsigned)sx < (unsigned)w, x1ok = (unsigned)(sx + 1) < (unsigned)w;
      const bool y0ok = (unsigned)sy < (unsigned)h, y1ok = (unsigned)(sy + 1) < (unsigned)h;
      int w00 = x0ok && y0ok ? wt.x : 0, w01 = x1ok && y0ok ? wt.y : 0, w10 = x0ok && y1ok ? wt.z : 0, w11 = x1ok && y1ok ? wt.w : 0;
      if (sx == -1) { w00 = w01; w01 = 0; w10 = w11; w11 = 0; }          // column 0 is the right-hand tap
      const int cx0 = min(max(sx, 0), w - 1), cy0 = min(max(sy, 0), h - 1), cy1 = min(max(sy + 1, 0), h - 1);
      wtop[k] = (uint32_t)w00 | ((uint32_t)w01 << 16); wbot[k] = (uint32_t)w10 | ((uint32_t)w11 << 16);
      tap[k] = cx0 | (cy0 << 14) | ((cy1 - cy0) << 28) | (1 << 29);
      bx0 = min(bx0, cx0); bx1 = max(bx1, min(cx0 + 1, w - 1)); by0 = min(by0, cy0); by1 = max(by1, cy1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
    bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
  }
  if (tx == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
  __syncthreads();
  bx0 = s_box[0]; by0 = s_box[1]; bx1 = s_box[2]; by1 = s_box[3];
  if (bx1 < 0) return;                                       // no live pixel in this tile
  const int b0 = blockIdx.z * group, nb = min(group, batch - b0);
  const int xb0 = (bx0 * CH) & ~15, xb1 = ((bx1 + 1) * CH + 15) & ~15;      // staged byte range of a source row
  const int P = xb1 - xb0 + 16, R = by1 - by0 + 1;          // 16 bytes of slack: the fetches below may read past a row's last tap
  uint8_t* out = dst + (int64_t)b0 * dstride + (int64_t)(blockIdx.y * UT_TILE_H + ty) * dpitch + x;
  const int64_t dp8 = 8 * dpitch;
  if (R * P > UT_STAGE_BYTES) {
    // the bounding box does not fit a stage (extreme distortion): the same taps from global memory
    for (int b = 0; b < nb; ++b, out += dstride) {
      const uint8_t* f = src + (int64_t)(b0 + b) * sstride;
#pragma unroll
      for (int k = 0; k < UT_PX; ++k) {
        if (!(tap[k] >> 29 & 1)) continue;
        const int cx0 = tap[k] & 0x3fff, cy0 = (tap[k] >> 14) & 0x3fff;
        const uint8_t* t = f + (int64_t)cy0 * spitch + (int64_t)cx0 * CH;
        const uint8_t* u = t + ((tap[k] >> 28) & 1) * spitch;
        const int dxb = cx0 + 1 < w ? CH : 0;                // (the weight of a tap past the last column is 0)
        const int w00 = wtop[k] & 0xffff, w01 = wtop[k] >> 16, w10 = wbot[k] & 0xffff, w11 = wbot[k] >> 16;
        int v[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
          v[c] = (w00 * __ldg(t + c) + w01 * __ldg(t + dxb + c) + w10 * __ldg(u + c) + w11 * __ldg(u + dxb + c) + (1 << 14)) >> 15;
        const int gray = CH == 3 ? (v[0] * 3735 + v[1] * 19235 + v[CH - 1] * 9798 + 16384) >> 15 : v[0];
        out[k * dp8] = (uint8_t)gray;
      }
    }
    return;
  }
  // ---- per-pixel offsets inside a stage; staging plan of this thread (<= 2 chunks of 16 bytes per frame) ----------
  uint32_t off[UT_PX];                     // bits 0-15: offset of the top-left tap, bits 16-31: offset of the tap row below
  uint32_t live = 0;
#pragma unroll
  for (int k = 0; k < UT_PX; ++k) {
    off[k] = 0;
    if (tap[k] >> 29 & 1) {
      const uint32_t o00 = (uint32_t)((((tap[k] >> 14) & 0x3fff) - by0) * P + (tap[k] & 0x3fff) * CH - xb0);
      off[k] = o00 | ((o00 + ((tap[k] >> 28) & 1) * P) << 16);
      live |= 1u << k;
    }
  }
  const int cpr = (xb1 - xb0) >> 4, n_chunks = R * cpr;      // <= UT_STAGE_BYTES / 16
  const uint8_t* g_ptr[UT_CHUNKS];
  uint32_t s_off[UT_CHUNKS];
#pragma unroll
  for (int q = 0; q < UT_CHUNKS; ++q) {
    const int j = tid + UT_THREADS * q;
    const int r = j / cpr, c = j - r * cpr;
    g_ptr[q] = src + (int64_t)b0 * sstride + (int64_t)(by0 + r) * spitch + xb0 + 16 * c;
    s_off[q] = j < n_chunks ? (uint32_t)(r * P + 16 * c) : 0xffffffffu;
  }
  const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(&s_buf[0][0]);
  int issued = 0;
  uint32_t issue_stage = buf0;
  auto issue = [&]() {
    if (issued < nb) {
#pragma unroll
      for (int q = 0; q < UT_CHUNKS; ++q) {
        if (s_off[q] != 0xffffffffu)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(issue_stage + s_off[q]), "l"(g_ptr[q]) : "memory");
        g_ptr[q] += sstride;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    ++issued;
    issue_stage = issue_stage == buf0 + (UT_STAGES - 1) * UT_STAGE_BYTES ? buf0 : issue_stage + UT_STAGE_BYTES;
  };
  issue();
  issue();
  uint32_t stage = buf0;
  for (int b = 0; b < nb; ++b, out += dstride) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");    // frame b has landed (frame b+1 may still travel)
    __syncthreads();                                         // ... for every thread; and the stage of frame b-1 is free again
    issue();                                                 // frame b+2 into it
#pragma unroll
    for (int k = 0; k < UT_PX; ++k) {
      const uint32_t o00 = stage + (off[k] & 0xffffu), o10 = stage + (off[k] >> 16);
      uint32_t gray;
      if (CH == 3) {
        // the two taps of a row are six consecutive bytes B0 G0 R0 B1 G1 R1: three aligned word loads + two funnel shifts
        // fetch them, one permute per channel puts (p0, p1) side by side and dp2a blends them with the pair of 16-bit weights
        const uint32_t sh = (o00 & 3u) * 8u;                 // P % 4 == 0: both rows have the same misalignment
        uint32_t t0, t1, t2, u0, u1, u2;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t0) : "r"(o00 & ~3u));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(t1) : "r"(o00 & ~3u));
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(t2) : "r"(o00 & ~3u));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u0) : "r"(o10 & ~3u));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(u1) : "r"(o10 & ~3u));
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(u2) : "r"(o10 & ~3u));
        const uint32_t tl = __funnelshift_r(t0, t1, sh), th = __funnelshift_r(t1, t2, sh);     // B0 G0 R0 B1 | G1 R1 . .
        const uint32_t ul = __funnelshift_r(u0, u1, sh), uh = __funnelshift_r(u1, u2, sh);
        const uint32_t wt = wtop[k], wb = wbot[k];
        // (B0 B1 G0 G1) in one word serves the blue (low half) and green (high half) blends
        const uint32_t tbg = __byte_perm(tl, th, 0x4130), ubg = __byte_perm(ul, uh, 0x4130);
        const uint32_t vb = __dp2a_lo(wb, ubg, __dp2a_lo(wt, tbg, 1u << 14)) >> 15;
        const uint32_t vg = __dp2a_hi(wb, ubg, __dp2a_hi(wt, tbg, 1u << 14)) >> 15;
        const uint32_t vr = __dp2a_lo(wb, __byte_perm(ul, uh, 0x0052), __dp2a_lo(wt, __byte_perm(tl, th, 0x0052), 1u << 14)) >> 15;
        gray = (vb * 3735u + vg * 19235u + vr * 9798u + 16384u) >> 15;
      } else {
        uint32_t p00, p01, p10, p11;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p00) : "r"(o00));
        asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(p01) : "r"(o00));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p10) : "r"(o10));
        asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(p11) : "r"(o10));
        gray = (__dp2a_lo(wbot[k], p10 | (p11 << 8), __dp2a_lo(wtop[k], p00 | (p01 << 8), 1u << 14))) >> 15;
      }
      if (live >> k & 1) out[k * dp8] = (uint8_t)gray;
    }
    stage = stage == buf0 + (UT_STAGES - 1) * UT_STAGE_BYTES ? buf0 : stage + UT_STAGE_BYTES;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
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
  U.inv_ab = 1.0 / (U.nfx * U.nfy); U.ir0 = 1.0 / U.nfx; U.ir1 = -U.nskew * U.inv_ab; U.ir4 = 1.0 / U.nfy;
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
  const bool aligned = ((reinterpret_cast<uintptr_t>(d_src) | (uintptr_t)src_pitch | (uintptr_t)src_stride) & 15) == 0 &&
                       w <= 16384 && h <= 16384;
  if (aligned) {
    // frames per CTA: the float64 pixel maps are computed once per CTA, so large batches amortise them over more frames
    const int group = batch >= 256 ? 64 : (batch >= 128 ? 32 : 16);
    const int groups = (batch + group - 1) / group;
    for (int g0 = 0; g0 < groups; g0 += 65535) {
      const int ng = groups - g0 < 65535 ? groups - g0 : 65535;
      const int64_t f0 = (int64_t)g0 * group;
      dim3 grid((ctx->und.roi_w + UT_TILE - 1) / UT_TILE, (ctx->und.roi_h + UT_TILE_H - 1) / UT_TILE_H, ng);
      if (channels == 3)
        undistort_gray_tiled_kernel<3><<<grid, UT_THREADS, 0, ctx->stream>>>(cam, ctx->und, ctx->d_remap_tab, d_src + f0 * src_stride, w, h,
                                                                             src_pitch, src_stride, d_gray + f0 * dst_stride, dst_pitch,
                                                                             dst_stride, (int)(batch - f0), group);
      else
        undistort_gray_tiled_kernel<1><<<grid, UT_THREADS, 0, ctx->stream>>>(cam, ctx->und, ctx->d_remap_tab, d_src + f0 * src_stride, w, h,
                                                                             src_pitch, src_stride, d_gray + f0 * dst_stride, dst_pitch,
                                                                             dst_stride, (int)(batch - f0), group);
      AGT_LAUNCH_CHECK(ctx);
    }
    return AGT_OK;
  }
  const int groups = (batch + kFramesPerThread - 1) / kFramesPerThread;
  for (int g0 = 0; g0 < groups; g0 += 65535) {
    const int ng = groups - g0 < 65535 ? groups - g0 : 65535;
    const int64_t f0 = (int64_t)g0 * kFramesPerThread;
    dim3 grid((ctx->und.roi_w + 127) / 128, ctx->und.roi_h, ng);
    undistort_gray_kernel<<<grid, 128, 0, ctx->stream>>>(cam, ctx->und, ctx->d_remap_tab, d_src + f0 * src_stride, w, h, channels, src_pitch,
                                                         src_stride, d_gray + f0 * dst_stride, dst_pitch, dst_stride,
                                                         (int)(batch - f0), 0);
    AGT_LAUNCH_CHECK(ctx);
  }
  return AGT_OK;
}

extern "C" int agt_undistort_frames(agt_ctx* ctx, const uint8_t* d_src, int w, int h, int channels, int64_t src_pitch, int64_t src_stride,
                             uint8_t* d_dst, int64_t dst_pitch, int64_t dst_stride, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;
  if (!ctx->camera_set || !ctx->undistort_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_undistort_frames: call agt_set_camera and agt_set_undistort first");
  if (!d_src || !d_dst || batch < 0 || (channels != 1 && channels != 3) || src_pitch < (int64_t)w * channels ||
      dst_pitch < (int64_t)ctx->und.roi_w * channels)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_frames: bad arguments");
  if (w != ctx->und.width || h != ctx->und.height)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_frames: frame is %dx%d but agt_set_undistort was given %dx%d", w, h, ctx->und.width, ctx->und.height);
  agt_camera cam = ctx->cam;
  if (!cam.has_dist) { cam.k1 = cam.k2 = cam.p1 = cam.p2 = cam.k3 = 0.0; }
  const int groups = (batch + kFramesPerThread - 1) / kFramesPerThread;
  for (int g0 = 0; g0 < groups; g0 += 65535) {
    const int ng = groups - g0 < 65535 ? groups - g0 : 65535;
    const int64_t f0 = (int64_t)g0 * kFramesPerThread;
    dim3 grid((ctx->und.roi_w + 127) / 128, ctx->und.roi_h, ng);
    undistort_gray_kernel<<<grid, 128, 0, ctx->stream>>>(cam, ctx->und, ctx->d_remap_tab, d_src + f0 * src_stride, w, h, channels, src_pitch,
                                                         src_stride, d_dst + f0 * dst_stride, dst_pitch, dst_stride, (int)(batch - f0), 1);
    AGT_LAUNCH_CHECK(ctx);
  }
  return AGT_OK;
}

extern "C" int agt_undistort_to_gray_host(agt_ctx* ctx, const uint8_t* h_src, int w, int h, int channels, uint8_t* h_gray) {
  return agt_undistort_frame_host(ctx, h_src, w, h, channels, nullptr, h_gray);
}

extern "C" int agt_undistort_frame_host(agt_ctx* ctx, const uint8_t* h_src, int w, int h, int channels, uint8_t* h_frame, uint8_t* h_gray) {
  if (!ctx) return AGT_ERR_INVALID;
  if (!ctx->undistort_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_undistort_frame_host: call agt_set_undistort first");
  if (!h_src || (!h_gray && !h_frame) || w < 1 || h < 1 || (channels != 1 && channels != 3))
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_undistort_frame_host: bad arguments");
  AGT_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t in_bytes = (size_t)w * h * channels, gray_bytes = (size_t)ctx->und.roi_w * ctx->und.roi_h, frame_bytes = gray_bytes * channels;
  uint8_t *din, *dout, *dframe;
  int rc;
  if ((rc = agt_scratch(ctx, 0, in_bytes, reinterpret_cast<void**>(&din)))) return rc;
  if ((rc = agt_scratch(ctx, 1, gray_bytes, reinterpret_cast<void**>(&dout)))) return rc;
  if (h_frame && (rc = agt_scratch(ctx, 2, frame_bytes, reinterpret_cast<void**>(&dframe)))) return rc;
  AGT_CUDA(ctx, cudaMemcpyAsync(din, h_src, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (h_gray) {
    if ((rc = agt_undistort_to_gray(ctx, din, w, h, channels, (int64_t)w * channels, (int64_t)in_bytes, dout, ctx->und.roi_w, (int64_t)gray_bytes, 1)))
      return rc;
    AGT_CUDA(ctx, cudaMemcpyAsync(h_gray, dout, gray_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (h_frame) {
    if ((rc = agt_undistort_frames(ctx, din, w, h, channels, (int64_t)w * channels, (int64_t)in_bytes, dframe, (int64_t)ctx->und.roi_w * channels,
                            (int64_t)frame_bytes, 1)))
      return rc;
    AGT_CUDA(ctx, cudaMemcpyAsync(h_frame, dframe, frame_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  AGT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AGT_OK;
}

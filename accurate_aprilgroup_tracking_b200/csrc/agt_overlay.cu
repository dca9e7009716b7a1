// N4: the overlay that follows the path, batched on the device.
//
// After every accepted pose the reference projects all 48 group corners and stamps a filled red disc on each
// (detect_pose.py:441-465 _project_draw_points -> cv.projectPoints; draw.py:120-153 draw_squares_and_3d_pts: np.round
// to integers, the bounds test `0 <= y < 720 and 0 <= x < 1280`, cv.circle(img, (x, y), 5, (0, 0, 255), -1)) - 0.7 ms of CPU
// per frame.  agt_project already yields the projections of a whole batch; this kernel stamps the discs into the batch of BGR
// frames, bit-identical to the cv.circle loop: the disc is OpenCV's midpoint-circle raster (row half-widths 5 4 4 4 3 0 for
// radius 5, computed on the host by the same integer recurrence), clipped to the frame; all discs have one colour, so
// overlapping stamps commute.  The second window's anti-aliased wire-frame (draw.py:40-118) stays on the host.
#include "agt_common.cuh"

namespace {

constexpr int MAX_RADIUS = 15;
struct DiscRows { int radius; int half[MAX_RADIUS + 1]; };

// thread block per (frame, point); threads over the (2r+1)^2 pixels of the stamp
__global__ void draw_points_kernel(uint8_t* __restrict__ frames, int w, int h, int64_t pitch, int64_t stride, const double* __restrict__ pts,
                                   const uint8_t* __restrict__ frame_mask, int n_pts, int bound_w, int bound_h, DiscRows disc,
                                   uint8_t b, uint8_t g, uint8_t r) {
  const int f = blockIdx.y, p = blockIdx.x;
  if (frame_mask != nullptr && frame_mask[f] == 0) return;
  const double px = pts[((int64_t)f * n_pts + p) * 2], py = pts[((int64_t)f * n_pts + p) * 2 + 1];
  if (!(fabs(px) < 1e9) || !(fabs(py) < 1e9)) return;
  const int cx = (int)rint(px), cy = (int)rint(py);                  // np.round: half to even
  if (cy < 0 || cy >= bound_h || cx < 0 || cx >= bound_w) return;    // the reference's test (draw.py:150), on the centre only
  const int n = 2 * disc.radius + 1;
  uint8_t* img = frames + (int64_t)f * stride;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int dy = i / n - disc.radius, dx = i - (i / n) * n - disc.radius;
    const int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
    const int x = cx + dx, y = cy + dy;
    if (adx <= disc.half[ady] && x >= 0 && x < w && y >= 0 && y < h) {
      uint8_t* o = img + (int64_t)y * pitch + 3 * x;
      o[0] = b; o[1] = g; o[2] = r;
    }
  }
}

}  // namespace

extern "C" int agt_draw_points(agt_ctx* ctx, uint8_t* d_bgr, int w, int h, int64_t pitch, int64_t stride, const double* d_pts,
                               const uint8_t* d_frame_mask, int batch, int n_pts, int radius, int bound_w, int bound_h, int blue, int green,
                               int red) {
  if (!ctx) return AGT_ERR_INVALID;
  if ((int64_t)batch * n_pts == 0) return AGT_OK;
  if (!d_bgr || !d_pts || batch < 0 || n_pts < 0 || w < 1 || h < 1 || pitch < 3 * (int64_t)w || radius < 0 || radius > MAX_RADIUS || batch > 65535)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_draw_points: bad arguments (radius 0..%d, batch <= 65535)", MAX_RADIUS);
  // cv::circle(.., thickness < 0): the midpoint-circle recurrence of OpenCV's drawing.cpp, restated: row |dy| spans |dx| <= half[|dy|]
  DiscRows disc;
  disc.radius = radius;
  for (int i = 0; i <= MAX_RADIUS; ++i) disc.half[i] = -1;
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    if (disc.half[dy] < dx) disc.half[dy] = dx;
    if (disc.half[dx] < dy) disc.half[dx] = dy;
    ++dy; err += plus; plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask; dx += mask; minus -= mask & 2;
  }
  draw_points_kernel<<<dim3((unsigned)n_pts, (unsigned)batch), 128, 0, ctx->stream>>>(d_bgr, w, h, pitch, stride, d_pts, d_frame_mask, n_pts, bound_w,
                                                                                      bound_h, disc, (uint8_t)blue, (uint8_t)green, (uint8_t)red);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

// Synthetic input generator (tests / bench only - NOT part of the tracking path).
// CUDA twin of accurate_aprilgroup_tracking_b200/synth.py:render: ray-casts the
// 12-tag dodecahedron of SURVEY.md 8d with a 4x4 Gaussian-weighted sub-sample
// pattern and an integer-hash noise field.  The reference ships no data
// (.gitignore:131-142), so benchmarks at 4096 x 1080p need frames made on device.
#include "agt_common.cuh"

namespace {

struct RenderParams {
  double fx, fy, cx, cy;
  double inradius, cell;
  int n_tags, w, h, noise;
  int64_t pitch, stride;
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void render_kernel(RenderParams P, const double* __restrict__ poses, const uint32_t* __restrict__ seeds,
                              const double* __restrict__ tag_rt, const uint8_t* __restrict__ cells,
                              uint8_t* __restrict__ frames) {
  __shared__ float s_R[9], s_o[3], s_n[AGT_MAX_TAGS][3], s_num[AGT_MAX_TAGS];
  __shared__ float s_rk[AGT_MAX_TAGS][9], s_tk[AGT_MAX_TAGS][3];
  __shared__ uint8_t s_cells[AGT_MAX_TAGS][100];
  __shared__ int s_box[4];
  const int f = blockIdx.z;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid == 0) {
    const double* ps = poses + (int64_t)f * 6;
    double r[3] = {ps[0], ps[1], ps[2]}, R[9];
    agt_rodrigues(r, R);
    double o[3];
    for (int i = 0; i < 3; ++i) o[i] = -(R[0 + i] * ps[3] + R[3 + i] * ps[4] + R[6 + i] * ps[5]);
    for (int i = 0; i < 9; ++i) s_R[i] = (float)R[i];
    for (int i = 0; i < 3; ++i) s_o[i] = (float)o[i];
    for (int k = 0; k < P.n_tags; ++k) {
      const double* rt = tag_rt + k * 12;
      double n[3] = {rt[2], rt[5], rt[8]};
      for (int i = 0; i < 9; ++i) s_rk[k][i] = (float)rt[i];
      for (int i = 0; i < 3; ++i) { s_tk[k][i] = (float)rt[9 + i]; s_n[k][i] = (float)n[i]; }
      s_num[k] = (float)(P.inradius - (n[0] * o[0] + n[1] * o[1] + n[2] * o[2]));
    }
    // bounding box of the projected body (same rule as synth.bounding_box)
    double circ = P.inradius * 1.2584086 + 1e-4;
    double tz = fmax(ps[5] - circ, 1e-3);
    double u = P.fx * ps[3] / ps[5] + P.cx, v = P.fy * ps[4] / ps[5] + P.cy;
    double rad = fmax(P.fx, P.fy) * circ / tz * 1.15 + 3.0;
    s_box[0] = max((int)floor(fmax(u - rad, -1e6)), 0);
    s_box[1] = max((int)floor(fmax(v - rad, -1e6)), 0);
    s_box[2] = min((int)ceil(fmin(u + rad, 1e6)) + 1, P.w);
    s_box[3] = min((int)ceil(fmin(v + rad, 1e6)) + 1, P.h);
  }
  for (int i = tid; i < P.n_tags * 100; i += blockDim.x * blockDim.y) s_cells[i / 100][i % 100] = cells[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= P.w || y >= P.h) return;
  float val = 128.f;
  if (x >= s_box[0] && x < s_box[2] && y >= s_box[1] && y < s_box[3]) {
    const float offs[4] = {-0.75f, -0.25f, 0.25f, 0.75f};
    float acc = 0.f, wsum = 0.f;
    const float half = (float)(5.0 * P.cell), icell = (float)(1.0 / P.cell);
    for (int sy = 0; sy < 4; ++sy)
      for (int sx = 0; sx < 4; ++sx) {
        float ox = offs[sx], oy = offs[sy];
        float w = __expf(-(ox * ox + oy * oy) * 2.0f);       // sigma 0.5 px
        float dcx = (float)((x + ox - P.cx) / P.fx), dcy = (float)((y + oy - P.cy) / P.fy);
        // d_obj = R^T d_cam
        float d0 = s_R[0] * dcx + s_R[3] * dcy + s_R[6];
        float d1 = s_R[1] * dcx + s_R[4] * dcy + s_R[7];
        float d2 = s_R[2] * dcx + s_R[5] * dcy + s_R[8];
        float te = -INFINITY, tx = INFINITY;
        int face = 0;
        for (int k = 0; k < P.n_tags; ++k) {
          float den = s_n[k][0] * d0 + s_n[k][1] * d1 + s_n[k][2] * d2;
          float t = s_num[k] / den;
          if (den < 0.f) { if (t > te) { te = t; face = k; } }
          else if (den > 0.f) { if (t < tx) tx = t; }
        }
        float colour = 128.f;
        if (te < tx && te > 0.f) {
          float p0 = s_o[0] + te * d0 - s_tk[face][0], p1 = s_o[1] + te * d1 - s_tk[face][1], p2 = s_o[2] + te * d2 - s_tk[face][2];
          // q = R_k^T p
          float qx = s_rk[face][0] * p0 + s_rk[face][3] * p1 + s_rk[face][6] * p2;
          float qy = s_rk[face][1] * p0 + s_rk[face][4] * p1 + s_rk[face][7] * p2;
          int col = (int)floorf((qx + half) * icell), row = (int)floorf((half - qy) * icell);
          colour = (col >= 0 && col < 10 && row >= 0 && row < 10) ? (float)s_cells[face][row * 10 + col] : 235.f;
        }
        acc += w * colour;
        wsum += w;
      }
    val = acc / wsum;
  }
  if (P.noise) {
    uint32_t salt = hash32(seeds[f]);
    uint32_t hsh = hash32((uint32_t)(y * P.w + x) ^ salt);
    int s = (int)(hsh & 0xff) + (int)((hsh >> 8) & 0xff) + (int)((hsh >> 16) & 0xff) + (int)(hsh >> 24);
    val += ((float)s - 510.f) * (2.0f / 147.79715829f);
  }
  float q = floorf(val + 0.5f);
  q = fminf(fmaxf(q, 0.f), 255.f);
  frames[(int64_t)f * P.stride + (int64_t)y * P.pitch + x] = (uint8_t)q;
}

}  // namespace

extern "C" int agt_render(agt_ctx* ctx, const double* d_pose, const uint32_t* d_seed, uint8_t* d_frames, int w, int h,
                          int64_t pitch, int64_t stride, const double* d_tag_rt, const uint8_t* d_cells, int n_tags,
                          double inradius, double cell, int noise, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_render: call agt_set_camera first");
  if (!d_pose || !d_seed || !d_frames || !d_tag_rt || !d_cells || n_tags < 1 || n_tags > AGT_MAX_TAGS || batch < 0 || w < 1 || h < 1)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_render: bad arguments");
  if (batch == 0) return AGT_OK;
  RenderParams P;
  P.fx = ctx->cam.fx; P.fy = ctx->cam.fy; P.cx = ctx->cam.cx; P.cy = ctx->cam.cy;
  P.inradius = inradius; P.cell = cell; P.n_tags = n_tags; P.w = w; P.h = h; P.noise = noise;
  P.pitch = pitch; P.stride = stride;
  dim3 block(32, 8);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((w + 31) / 32, (h + 7) / 8, nb);
    render_kernel<<<grid, block, 0, ctx->stream>>>(P, d_pose + (int64_t)b0 * 6, d_seed + b0, d_tag_rt, d_cells,
                                                   d_frames + (int64_t)b0 * stride);
    AGT_LAUNCH_CHECK(ctx);
  }
  return AGT_OK;
}

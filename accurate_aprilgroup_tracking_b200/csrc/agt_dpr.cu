// K4: dense pose refinement - batched photometric Levenberg-Marquardt that never
// leaves the device.  One CTA per (frame, hypothesis).
//
// No reference code exists for this stage (README.md:20 names it as future work);
// the executable specification is oracle/dpr_oracle.py:
// visibility and pyramid level frozen at the initial pose; per sample a pinhole
// projection, bilinear intensity and bilinear Scharr gradient of level l, residual
// r = I - O, 1x6 Jacobian [Y x g, g] for the left perturbation R <- exp(w) R;
// H = J^T J (21), b = J^T r (6), c = 1/2 sum r^2; LM with lambda*diag(H), accept
// iff the cost decreases, stop rules as in the oracle.
//
// B200 mapping: the level-l region of interest (<= 272x272 u8, the projected
// bounding sphere plus a drift margin) is staged ONCE per CTA into shared memory
// with 128-bit loads and reused by every LM evaluation, so HBM sees each ROI byte
// once per refinement; no gradient plane is materialised - the 4x4 u8 footprint of
// a sample is fetched with two aligned 32-bit shared loads + a byte permute per
// row and the four Scharr values come from 16 dp4a.  The projection of a sample
// runs in float64 (see the note in the sample loop), intensity / gradient /
// Jacobian in float32; the 28 sums are reduced by warp shuffles and across warps in float64
// (the cost in float64 from the start); one thread runs the 6x6 Cholesky / LM
// bookkeeping in float64.  Samples whose footprint leaves the staged tile fall
// back to global loads, so results do not depend on the tile size.
#include <cooperative_groups.h>

#include "agt_dpr_plan.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int DPR_THREADS = 256;
constexpr int DPR_WARPS = DPR_THREADS / 32;
constexpr int TILE_ROWS = AGT_DPR_TILE_ROWS;
constexpr int TILE_PITCH = AGT_DPR_TILE_PITCH;
constexpr int TILE_BYTES = TILE_ROWS * TILE_PITCH + 16;   // +16: the unaligned fetch reads one word past its window
constexpr int NSUM = 28;                   // 21 H + 6 b + (cost kept separately in double) + count
constexpr double COS_VISIBLE = 0.25881904510252074;   // cos 75 deg
constexpr double LAMBDA0 = 1e-3, LAMBDA_MIN = 1e-9, LAMBDA_MAX = 1e6;
constexpr int MAX_EVALS = 50;
constexpr double TOL_ROT = 1e-6, TOL_TRANS = 1e-6, REJ_TOL_ROT = 5e-5, REJ_TOL_TRANS = 5e-6;

struct DprShared {
  // trial pose (float32 view used by the sample loop, float64 for the projection)
  float R[9], t[3];
  __align__(16) double Rd[12];   // trial rotation (row-major) followed by the translation: six 128-bit broadcast loads
  double td_unused_[1];
  double fxs, fys, ubase, vbase;   // fx*2^-l, fy*2^-l, cx*2^-l, cy*2^-l
  float fx, fy, cx, cy;
  float inv_scale;       // 2^-level
  float gscale;          // 2^-level / 32
  int level, lw, lh;
  int64_t lpitch;
  const uint8_t* limg;   // level image of this frame
  int tx0, ty0, tw, th;  // staged tile: origin (level px), size
  int rx0, ry0, rx1, ry1;   // predicted ROI (see agt_dpr_plan)
  int left_roi;          // a sample footprint left the predicted ROI at some evaluation
  int n_active;
  int act_begin[AGT_MAX_TAGS];   // first sample of each active tag
  int act_prefix[AGT_MAX_TAGS + 1];
  int stop;
  double wsum[DPR_WARPS][NSUM + 2];
  double tot[NSUM + 2];
  // LM state, touched by thread 0 only (kept out of registers)
  double Rc[9], tc[3], Hc[21], bc[6], cc, lam;
  double Rt[9], tt[3], dstep[6];
  int nc, evals, status;
};

__device__ __forceinline__ uint32_t ld4_unaligned_smem(const uint8_t* base, int off) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (off & ~3));
  return __byte_perm(w[0], w[1], 0x3210 + 0x1111 * (off & 3));
}

__device__ __forceinline__ uint32_t ld4_global(const uint8_t* p) {
  return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
}

// d = acc + sum_k px.byte[k] (unsigned) * coef.byte[k] (signed)
__device__ __forceinline__ int dp4c(uint32_t px, int coef, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(coef), "r"(acc));
  return d;
}

// ---- serial LM phase helpers (one thread per CTA runs them while 255 wait: keep the dependency chains short) ----
// 1/sqrt(d) in float64 from the MUFU.RSQ seed + two Newton steps (a sqrt + a division cost ~70 dependent instructions)
__device__ __forceinline__ double rsqrt_f64(double d) {
  double y = (double)rsqrtf((float)d);
  y = y * (1.5 - 0.5 * d * y * y);
  y = y * (1.5 - 0.5 * d * y * y);
  return y;
}

// Cholesky solve of the SPD 6x6 system A x = b (A full, row-major), multiplications by 1/L_jj instead of divisions.
__device__ inline bool chol6_solve_fast(double A[36], double b[6]) {
  double inv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * 6 + k] * A[j * 6 + k];
    if (!(d > 1e-30 && d < 1e30)) return agt_chol6_solve(A, b) && false;   // out of the float seed's range: not PD for our purposes
    const double y = rsqrt_f64(d);
    inv[j] = y;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = v * y;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= A[i * 6 + k] * b[k];
    b[i] = v * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= A[k * 6 + i] * b[k];
    b[i] = v * inv[i];
  }
  return true;
}

// exp([w]x) for an LM step: series for |w| < 0.5 (truncation < 1e-16), libm sincos otherwise
__device__ inline void rodrigues_step(const double w[3], double R[9]) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (t2 >= 0.25) { agt_rodrigues(w, R); return; }
  const double a = 1.0 - t2 / 6.0 * (1.0 - t2 / 20.0 * (1.0 - t2 / 42.0 * (1.0 - t2 / 72.0 * (1.0 - t2 / 110.0 * (1.0 - t2 / 156.0)))));
  const double b = 0.5 * (1.0 - t2 / 12.0 * (1.0 - t2 / 30.0 * (1.0 - t2 / 56.0 * (1.0 - t2 / 90.0 * (1.0 - t2 / 132.0 * (1.0 - t2 / 182.0))))));
  const double c = 1.0 - b * t2;
  R[0] = c + b * w[0] * w[0];        R[1] = b * w[0] * w[1] - a * w[2]; R[2] = b * w[0] * w[2] + a * w[1];
  R[3] = b * w[1] * w[0] + a * w[2]; R[4] = c + b * w[1] * w[1];        R[5] = b * w[1] * w[2] - a * w[0];
  R[6] = b * w[2] * w[0] - a * w[1]; R[7] = b * w[2] * w[1] + a * w[0]; R[8] = c + b * w[2] * w[2];
}

// packed signed-byte coefficient words (little endian: byte 0 multiplies the left-most pixel)
constexpr int C_DX0_3 = (int)0x000300FD;    // ( -3,  0,  3,  0)   horizontal difference at x0, row weight 3
constexpr int C_DX0_10 = (int)0x000A00F6;   // (-10,  0, 10,  0)   ... row weight 10
constexpr int C_DX1_3 = (int)0x0300FD00;    // (  0, -3,  0,  3)   horizontal difference at x0+1
constexpr int C_DX1_10 = (int)0x0A00F600;   // (  0,-10,  0, 10)
constexpr int C_SM0 = (int)0x00030A03;      // (  3, 10,  3,  0)   horizontal smoothing at x0
constexpr int C_SM1 = (int)0x030A0300;      // (  0,  3, 10,  3)   horizontal smoothing at x0+1
constexpr int C_SM0_NEG = (int)0x00FDF6FD;  // ( -3,-10, -3,  0)
constexpr int C_SM1_NEG = (int)0xFDF6FD00;  // (  0, -3,-10, -3)

// kCluster > 1: a thread-block cluster of kCluster CTAs shares one refinement.  Used when the batch is smaller than
// the machine (camera streams: one pose per stream per step): every CTA stages the ROI, takes every kCluster-th
// slice of the samples, the partial sums meet in CTA 0 through distributed shared memory, CTA 0 runs the LM step
// and the others read the new trial pose back over DSMEM.  Two cluster barriers per evaluation.
template <int kCluster>
__global__ void __launch_bounds__(DPR_THREADS, 2)
dpr_kernel(agt_pyramid pyr, agt_camera cam, const float4* __restrict__ samples, agt_model model,
           const double* __restrict__ init, int n_hyp, const uint8_t* __restrict__ mask, double* __restrict__ pose_out, float* __restrict__ cost_out,
           int32_t* __restrict__ nvalid_out, int32_t* __restrict__ evals_out, uint8_t* __restrict__ status_out,
           uint8_t* __restrict__ left_roi_out) {
  extern __shared__ __align__(16) uint8_t s_tile[];
  __shared__ DprShared S;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t job = blockIdx.x / kCluster;
  const int crank = kCluster > 1 ? (int)(blockIdx.x % kCluster) : 0;     // == cluster.block_rank() for (kCluster,1,1) clusters
  const int64_t frame = job / n_hyp;
  if (mask != nullptr && mask[frame] == 0) return;     // skipped frame: none of its outputs is written (whole cluster)

  double* const Rc = S.Rc; double* const tc = S.tc; double* const Hc = S.Hc; double* const bc = S.bc;
  double* const Rt = S.Rt; double* const tt = S.tt; double* const dstep = S.dstep;

  if (tid == 0) {
    S.cc = 0.0; S.lam = LAMBDA0; S.nc = 0; S.evals = 0; S.status = AGT_DPR_MAX_EVALS;
    const double* p0 = init + job * 6;
    double r0[3] = {p0[0], p0[1], p0[2]};
    agt_rodrigues(r0, Rc);
    tc[0] = p0[3]; tc[1] = p0[4]; tc[2] = p0[5];
    const agt_dpr_plan plan = agt_make_dpr_plan(cam, model.pitch, model.radius, tc, pyr.width, pyr.height, pyr.levels);
    const int lvl = plan.level;
    S.level = lvl;
    S.lw = pyr.width[lvl]; S.lh = pyr.height[lvl]; S.lpitch = pyr.pitch[lvl];
    S.limg = pyr.data[lvl] + frame * pyr.frame_stride[lvl];
    S.rx0 = plan.rx0; S.ry0 = plan.ry0; S.rx1 = plan.rx1; S.ry1 = plan.ry1;
    S.tx0 = plan.tx0; S.ty0 = plan.ty0; S.tw = plan.tw; S.th = plan.th;
    S.left_roi = 0;
    double sc = 1.0 / (double)(1 << lvl);
    S.inv_scale = (float)sc;
    S.gscale = (float)(sc / 32.0);
    S.fx = (float)cam.fx; S.fy = (float)cam.fy; S.cx = (float)cam.cx; S.cy = (float)cam.cy;
    // active tags
    int na = 0, pre = 0;
    for (int k = 0; k < model.n_tags; ++k) {
      double c[3], n[3];
      for (int i = 0; i < 3; ++i) {
        c[i] = Rc[i * 3] * model.centres[k][0] + Rc[i * 3 + 1] * model.centres[k][1] + Rc[i * 3 + 2] * model.centres[k][2] + tc[i];
        n[i] = Rc[i * 3] * model.normals[k][0] + Rc[i * 3 + 1] * model.normals[k][1] + Rc[i * 3 + 2] * model.normals[k][2];
      }
      double cn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      double d = -(n[0] * c[0] + n[1] * c[1] + n[2] * c[2]) / cn;
      if (d > COS_VISIBLE) {
        S.act_begin[na] = model.tag_begin[k];
        S.act_prefix[na] = pre;
        pre += model.tag_begin[k + 1] - model.tag_begin[k];
        ++na;
      }
    }
    S.act_prefix[na] = pre;
    S.n_active = na;
    for (int i = 0; i < 9; ++i) { S.R[i] = (float)Rc[i]; S.Rd[i] = Rc[i]; }
    for (int i = 0; i < 3; ++i) { S.t[i] = (float)tc[i]; S.Rd[9 + i] = tc[i]; }
    S.fxs = cam.fx * sc; S.fys = cam.fy * sc;
    S.ubase = cam.cx * sc; S.vbase = cam.cy * sc;
    S.stop = 0;
  }
  __syncthreads();

  // ---- stage the ROI tile (128-bit loads; the level rows are 16 B aligned) ----------
  {
    const int tw = S.tw, th = S.th;
    const uint8_t* src = S.limg + (int64_t)S.ty0 * S.lpitch + S.tx0;
    const bool vec = ((S.lpitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(S.limg) & 15) == 0);
    const int chunks = (tw + 15) >> 4;
    for (int i = tid; i < th * chunks; i += DPR_THREADS) {
      int r = i / chunks, c = i - r * chunks;
      const uint8_t* g = src + (int64_t)r * S.lpitch + 16 * c;
      uint4 v;
      if (vec && S.tx0 + 16 * c + 16 <= (int)S.lpitch) {
        v = __ldg(reinterpret_cast<const uint4*>(g));
      } else {
        uint32_t w4[4] = {0, 0, 0, 0};
        for (int b = 0; b < 16; ++b)
          if (S.tx0 + 16 * c + b < S.lw) w4[b >> 2] |= (uint32_t)__ldg(g + b) << (8 * (b & 3));
        v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
      *reinterpret_cast<uint4*>(s_tile + r * TILE_PITCH + 16 * c) = v;
    }
  }
  __syncthreads();

  const int n_act_samples = S.act_prefix[S.n_active];
  const float fx = S.fx, fy = S.fy, cx = S.cx, cy = S.cy, isc = S.inv_scale, gsc = S.gscale;
  const int lw = S.lw, lh = S.lh, tx0 = S.tx0, ty0 = S.ty0, tw = S.tw, th = S.th;
  const int64_t lpitch = S.lpitch;
  const uint8_t* limg = S.limg;

  while (true) {
    // ================= evaluate cost + normal equations at the trial pose =================
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    double cost = 0.0;
    int cnt = 0;
    // rotation / translation of the trial pose: float64 copies for the projection, float32 for the Jacobian
    const float R0 = S.R[0], R1 = S.R[1], R2 = S.R[2], R3 = S.R[3], R4 = S.R[4], R5 = S.R[5], R6 = S.R[6], R7 = S.R[7],
                R8 = S.R[8], t0 = S.t[0], t1 = S.t[1];
    int seg = 0;
    // software pipeline: the model record of the next sample is requested before this one is consumed
    float4 sm_next = make_float4(0.f, 0.f, 0.f, 0.f);
    const int j0 = crank * DPR_THREADS + tid;
    if (j0 < n_act_samples) {
      while (j0 >= S.act_prefix[seg + 1]) ++seg;
      sm_next = __ldg(&samples[S.act_begin[seg] + (j0 - S.act_prefix[seg])]);
    }
    for (int j = j0; j < n_act_samples; j += kCluster * DPR_THREADS) {
      const float4 sm = sm_next;
      const int jn = j + kCluster * DPR_THREADS;
      if (jn < n_act_samples) {
        while (jn >= S.act_prefix[seg + 1]) ++seg;
        sm_next = __ldg(&samples[S.act_begin[seg] + (jn - S.act_prefix[seg])]);
      }
      // Projection in float64: the accept/reject decisions of the LM loop compare costs that differ by ~1e-6
      // relative near convergence; float32 pixel coordinates (ulp 6e-5 px at 1080p) add ~1e-6 of noise to the
      // cost and flip 20 % of those decisions, float64 leaves 0.2 % (profiles/r01_dpr_precision_sweep.log).
      const double sx = sm.x, sy = sm.y, sz = sm.z;
      // (the float64 pose is read from shared memory with broadcast loads: 24 registers less per thread)
      const double2* P = reinterpret_cast<const double2*>(S.Rd);
      const double2 p01 = P[0], p23 = P[1], p45 = P[2], p67 = P[3], p8t = P[4], ptt = P[5];
      const double dX = p01.x * sx + p01.y * sy + p23.x * sz + p8t.y;
      const double dY = p23.y * sx + p45.x * sy + p45.y * sz + ptt.x;
      const double dZ = p67.x * sx + p67.y * sy + p8t.x * sz + ptt.y;
      if (!(dZ > 1e-6)) continue;
      float iz;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"((float)dZ));       // MUFU.RCP: Jacobian + Newton seed
      double r0 = (double)iz;                         // 2^-23 relative; one Newton step -> 2^-46 (1e-11 px at 1080p)
      r0 = r0 * (2.0 - dZ * r0);
      const double ul = (S.fxs * dX) * r0 + S.ubase, vl = (S.fys * dY) * r0 + S.vbase;
      // valid <=> 1 <= floor(ul) <= lw-3  <=>  1 <= ul < lw-2   (NaN fails both)
      if (!(ul >= 1.0 && ul < (double)(lw - 2) && vl >= 1.0 && vl < (double)(lh - 2))) continue;
      // floor + fraction without conversions: adding 2^52+2^51 rounding down leaves floor(ul) in the low mantissa word
      const double kMagic = 6755399441055744.0;
      const double tu = __dadd_rd(ul, kMagic), tv = __dadd_rd(vl, kMagic);
      const int x0 = __double2loint(tu), y0 = __double2loint(tv);
      const float a = (float)(ul - (tu - kMagic)), b = (float)(vl - (tv - kMagic));
      // float32 copies for the Jacobian (rounding of these only perturbs the LM step, not the cost)
      const float Yx = R0 * sm.x + R1 * sm.y + R2 * sm.z;
      const float Yy = R3 * sm.x + R4 * sm.y + R5 * sm.z;
      const float Yz = R6 * sm.x + R7 * sm.y + R8 * sm.z;
      const float X = Yx + t0, Y = Yy + t1;
      // 4x4 footprint rows y0-1..y0+2, columns x0-1..x0+2
      uint32_t row[4];
      const int lx = x0 - 1 - tx0, ly = y0 - 1 - ty0;
      if (lx >= 0 && ly >= 0 && lx + 4 <= tw && ly + 4 <= th) {
        const int off = ly * TILE_PITCH + lx;
#pragma unroll
        for (int r = 0; r < 4; ++r) row[r] = ld4_unaligned_smem(s_tile, off + r * TILE_PITCH);
      } else {
        if (x0 - 1 < S.rx0 || y0 - 1 < S.ry0 || x0 + 3 > S.rx1 || y0 + 3 > S.ry1) S.left_roi = 1;   // benign race: all write 1
        const uint8_t* g = limg + (int64_t)(y0 - 1) * lpitch + (x0 - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) row[r] = ld4_global(g + r * lpitch);
      }
      // Scharr at (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1): the [3 10 3] row weights are folded into the byte
      // coefficients and the three rows are chained through the dp4a accumulator - no integer combine afterwards
      const float gx00 = (float)dp4c(row[2], C_DX0_3, dp4c(row[1], C_DX0_10, dp4c(row[0], C_DX0_3, 0)));
      const float gx01 = (float)dp4c(row[2], C_DX1_3, dp4c(row[1], C_DX1_10, dp4c(row[0], C_DX1_3, 0)));
      const float gx10 = (float)dp4c(row[3], C_DX0_3, dp4c(row[2], C_DX0_10, dp4c(row[1], C_DX0_3, 0)));
      const float gx11 = (float)dp4c(row[3], C_DX1_3, dp4c(row[2], C_DX1_10, dp4c(row[1], C_DX1_3, 0)));
      const float gy00 = (float)dp4c(row[2], C_SM0, dp4c(row[0], C_SM0_NEG, 0));
      const float gy01 = (float)dp4c(row[2], C_SM1, dp4c(row[0], C_SM1_NEG, 0));
      const float gy10 = (float)dp4c(row[3], C_SM0, dp4c(row[1], C_SM0_NEG, 0));
      const float gy11 = (float)dp4c(row[3], C_SM1, dp4c(row[1], C_SM1_NEG, 0));
      const float i00 = (float)((row[1] >> 8) & 0xffu), i01 = (float)((row[1] >> 16) & 0xffu);
      const float i10 = (float)((row[2] >> 8) & 0xffu), i11 = (float)((row[2] >> 16) & 0xffu);
      const float w11 = a * b, w01 = a - w11, w10 = b - w11, w00 = 1.f - a - b + w11;
      const float I = w00 * i00 + w01 * i01 + w10 * i10 + w11 * i11;
      const float Gx = (w00 * gx00 + w01 * gx01 + w10 * gx10 + w11 * gx11) * gsc;
      const float Gy = (w00 * gy00 + w01 * gy01 + w10 * gy10 + w11 * gy11) * gsc;
      const float r = I - sm.w;
      const float g0 = Gx * fx * iz, g1 = Gy * fy * iz;
      const float g2 = -(g0 * X + g1 * Y) * iz;
      float J[6];
      J[0] = Yy * g2 - Yz * g1;
      J[1] = Yz * g0 - Yx * g2;
      J[2] = Yx * g1 - Yy * g0;
      J[3] = g0; J[4] = g1; J[5] = g2;
      int k = 0;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = p; q < 6; ++q) acc[k++] += J[p] * J[q];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[21 + p] += J[p] * r;
      cost += (double)r * (double)r;
      ++cnt;
    }
    // ---- reduce: shuffles within the warp, float64 across warps ------------------------
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = agt_warp_sum(acc[k]);
    cost = agt_warp_sum(cost);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 27; ++k) S.wsum[wid][k] = (double)acc[k];
      S.wsum[wid][27] = cost;
      S.wsum[wid][28] = (double)cnt;
    }
    __syncthreads();
    if (wid == 0) {
      if (lane < 29) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < DPR_WARPS; ++w) s += S.wsum[w][lane];
        S.tot[lane] = s;
      }
      __syncwarp();
    }
    if (kCluster > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();                                   // every CTA's partial sums are in its S.tot
      if (crank == 0 && wid == 0) {
        if (lane < 29) {
          double s = S.tot[lane];
          for (int r = 1; r < kCluster; ++r) s += cluster.map_shared_rank(&S, r)->tot[lane];
          S.tot[lane] = s;
        }
        __syncwarp();
      }
    }

    // ================= LM bookkeeping (one thread, float64) ================================
    if (tid == 0 && crank == 0) {
      double Hn[21], bn[6];
      for (int k = 0; k < 21; ++k) Hn[k] = S.tot[k];
      for (int k = 0; k < 6; ++k) bn[k] = S.tot[21 + k];
      double cn = 0.5 * S.tot[27];
      int nn = (int)S.tot[28];
      double cc = S.cc, lam = S.lam;
      int nc = S.nc, evals = S.evals + 1, status = S.status;
      bool need_step = false;
      if (evals == 1) {
        for (int k = 0; k < 21; ++k) Hc[k] = Hn[k];
        for (int k = 0; k < 6; ++k) bc[k] = bn[k];
        cc = cn; nc = nn;
        if (nn == 0) status = AGT_DPR_NONE; else need_step = true;
      } else {
        double nw = sqrt(dstep[0] * dstep[0] + dstep[1] * dstep[1] + dstep[2] * dstep[2]);
        double nt = sqrt(dstep[3] * dstep[3] + dstep[4] * dstep[4] + dstep[5] * dstep[5]);
        if (nn > 0 && cn < cc) {
          for (int k = 0; k < 9; ++k) Rc[k] = Rt[k];
          for (int k = 0; k < 3; ++k) tc[k] = tt[k];
          for (int k = 0; k < 21; ++k) Hc[k] = Hn[k];
          for (int k = 0; k < 6; ++k) bc[k] = bn[k];
          cc = cn; nc = nn;
          lam = fmax(lam / 10.0, LAMBDA_MIN);
          if (nw < TOL_ROT && nt < TOL_TRANS) status = AGT_DPR_CONVERGED; else need_step = true;
        } else {
          lam *= 10.0;
          if (nw < REJ_TOL_ROT && nt < REJ_TOL_TRANS) status = AGT_DPR_CONVERGED;
          else if (lam > LAMBDA_MAX) status = AGT_DPR_LAMBDA;
          else need_step = true;
        }
      }
      if (need_step && evals >= MAX_EVALS) need_step = false;     // status stays MAX_EVALS
      while (need_step) {
        double A[36], d[6];
        int k = 0;
        for (int p = 0; p < 6; ++p)
          for (int q = p; q < 6; ++q) { A[p * 6 + q] = Hc[k]; A[q * 6 + p] = Hc[k]; ++k; }
        for (int p = 0; p < 6; ++p) { A[p * 6 + p] += lam * A[p * 6 + p]; d[p] = -bc[p]; }
        if (chol6_solve_fast(A, d)) {
          for (int p = 0; p < 6; ++p) dstep[p] = d[p];
          double E[9];
          rodrigues_step(d, E);
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = E[r * 3] * Rc[c] + E[r * 3 + 1] * Rc[3 + c] + E[r * 3 + 2] * Rc[6 + c];
          for (int p = 0; p < 3; ++p) tt[p] = tc[p] + d[3 + p];
          for (int p = 0; p < 9; ++p) { S.R[p] = (float)Rt[p]; S.Rd[p] = Rt[p]; }
          for (int p = 0; p < 3; ++p) { S.t[p] = (float)tt[p]; S.Rd[9 + p] = tt[p]; }
          break;
        }
        lam *= 10.0;
        if (lam > LAMBDA_MAX) { status = AGT_DPR_LAMBDA; need_step = false; }
      }
      S.stop = need_step ? 0 : 1;
      S.cc = cc; S.lam = lam; S.nc = nc; S.evals = evals; S.status = status;
    }
    if (kCluster > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();                                   // CTA 0 has published the next trial pose (or stop)
      if (crank != 0) {
        const DprShared* S0 = cluster.map_shared_rank(&S, 0);
        if (tid < 9) { S.R[tid] = S0->R[tid]; S.Rd[tid] = S0->Rd[tid]; }
        if (tid < 3) { S.t[tid] = S0->t[tid]; S.Rd[9 + tid] = S0->Rd[9 + tid]; }
        if (tid == 0) S.stop = S0->stop;
      }
    }
    __syncthreads();
    if (S.stop) break;
  }
  if (kCluster > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (crank == 0 && tid == 0)
      for (int r = 1; r < kCluster; ++r) S.left_roi |= cluster.map_shared_rank(&S, r)->left_roi;
    cluster.sync();                                     // keep every CTA's shared memory alive until CTA 0 has read it
    if (crank != 0) return;
  }

  if (tid == 0) {
    double rv[3];
    agt_log_rotation(Rc, rv);
    double* po = pose_out + job * 6;
    po[0] = rv[0]; po[1] = rv[1]; po[2] = rv[2]; po[3] = tc[0]; po[4] = tc[1]; po[5] = tc[2];
    if (cost_out) cost_out[job] = (float)S.cc;
    if (nvalid_out) nvalid_out[job] = S.nc;
    if (evals_out) evals_out[job] = S.evals;
    if (status_out) status_out[job] = (uint8_t)S.status;
    if (left_roi_out) left_roi_out[job] = (uint8_t)S.left_roi;
  }
}

__global__ void dpr_rects_kernel(agt_pyramid pyr, agt_camera cam, double pitch, double radius, const double* __restrict__ init,
                                 int n_hyp, int32_t* __restrict__ rects, int batch) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= batch) return;
  int32_t r[4];
  agt_dpr_rect_l0(cam, pitch, radius, init + (int64_t)f * n_hyp * 6, n_hyp, pyr.width, pyr.height, pyr.levels, r);
  *reinterpret_cast<int4*>(rects + (int64_t)f * 4) = make_int4(r[0], r[1], r[2], r[3]);
}

__global__ void any_flag_kernel(const uint8_t* __restrict__ flags, int stride, uint8_t* __restrict__ out, int batch) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= batch) return;
  uint8_t any = 0;
  for (int k = 0; k < stride; ++k) any |= flags[(int64_t)f * stride + k];
  out[f] = any ? 1 : 0;
}

__global__ void select_best_kernel(const double* __restrict__ pose, const float* __restrict__ cost,
                                   const int32_t* __restrict__ nvalid, int n_hyp, int32_t* __restrict__ best,
                                   double* __restrict__ best_pose, int batch) {
  // one warp per frame: argmin of 2c/n, ties -> lowest index
  int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (f >= batch) return;
  double bs = INFINITY;
  int bi = 0x7fffffff;
  for (int h = lane; h < n_hyp; h += 32) {
    int n = nvalid[(int64_t)f * n_hyp + h];
    double s = n > 0 ? 2.0 * (double)cost[(int64_t)f * n_hyp + h] / (double)n : INFINITY;
    if (s < bs || (s == bs && h < bi)) { bs = s; bi = h; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    double os = __shfl_xor_sync(0xffffffffu, bs, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
  if (lane == 0) best[f] = bi;
  if (best_pose && lane < 6) best_pose[(int64_t)f * 6 + lane] = pose[((int64_t)f * n_hyp + bi) * 6 + lane];
}

}  // namespace

extern "C" int agt_refine(agt_ctx* ctx, const agt_pyramid* pyr, const double* d_init, int n_hyp, const uint8_t* d_mask,
                          double* d_pose, float* d_cost, int32_t* d_n_valid, int32_t* d_evals, uint8_t* d_status,
                          uint8_t* d_left_roi, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set || !ctx->model_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_refine: camera and surface model must be set");
  if (!pyr || !d_init || !d_pose || n_hyp < 1 || batch < 0) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_refine: bad arguments");
  if (pyr->levels < 1 || pyr->levels > AGT_MAX_LEVELS) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_refine: bad pyramid descriptor");
  if (ctx->cam.has_dist) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_refine: lens distortion is not supported; undistort the frame first (detect_pose.py:611-619)");
  int64_t jobs = (int64_t)batch * n_hyp;
  if (jobs == 0) return AGT_OK;
  if (jobs > 0x7fffffffLL) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_refine: batch*n_hyp too large");
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    AGT_CUDA(ctx, cudaFuncSetAttribute(dpr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    AGT_CUDA(ctx, cudaFuncSetAttribute(dpr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    AGT_CUDA(ctx, cudaFuncSetAttribute(dpr_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    AGT_CUDA(ctx, cudaFuncSetAttribute(dpr_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    attr_set[ctx->device & 63] = true;
  }
  // 2 CTAs/SM at 128 registers: a 3-CTA build (80 registers) spills in the sample loop and measured 20 % slower.
  // Small batches (fewer refinements than 2 CTA slots per SM) are spread over clusters of 2 / 4 / 8 CTAs.
  int cluster = 1;
  const int64_t slots = 2LL * ctx->sm_count;
  while (cluster < 8 && jobs * (cluster * 2) <= slots) cluster *= 2;
  if (cluster == 1) {
    dpr_kernel<1><<<(unsigned)jobs, DPR_THREADS, TILE_BYTES, ctx->stream>>>(*pyr, ctx->cam, ctx->model.samples, ctx->model, d_init,
                                                                         n_hyp, d_mask, d_pose, d_cost, d_n_valid, d_evals, d_status,
                                                                         d_left_roi);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(jobs * cluster));
    cfg.blockDim = dim3(DPR_THREADS);
    cfg.dynamicSmemBytes = TILE_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    agt_pyramid pv = *pyr;
    const float4* smp = ctx->model.samples;
    cudaError_t e;
    if (cluster == 2)
      e = cudaLaunchKernelEx(&cfg, dpr_kernel<2>, pv, ctx->cam, smp, ctx->model, d_init, n_hyp, d_mask, d_pose, d_cost, d_n_valid, d_evals, d_status, d_left_roi);
    else if (cluster == 4)
      e = cudaLaunchKernelEx(&cfg, dpr_kernel<4>, pv, ctx->cam, smp, ctx->model, d_init, n_hyp, d_mask, d_pose, d_cost, d_n_valid, d_evals, d_status, d_left_roi);
    else
      e = cudaLaunchKernelEx(&cfg, dpr_kernel<8>, pv, ctx->cam, smp, ctx->model, d_init, n_hyp, d_mask, d_pose, d_cost, d_n_valid, d_evals, d_status, d_left_roi);
    if (e != cudaSuccess) AGT_FAIL(ctx, AGT_ERR_CUDA, "agt_refine: cluster launch failed: %s", cudaGetErrorString(e));
  }
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_select_best(agt_ctx* ctx, const double* d_pose, const float* d_cost, const int32_t* d_n_valid,
                               int n_hyp, int32_t* d_best, double* d_best_pose, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!d_pose || !d_cost || !d_n_valid || !d_best || n_hyp < 1 || batch < 0)
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_select_best: bad arguments");
  if (batch == 0) return AGT_OK;
  int threads = 128;
  int blocks = (int)(((int64_t)batch * 32 + threads - 1) / threads);
  select_best_kernel<<<blocks, threads, 0, ctx->stream>>>(d_pose, d_cost, d_n_valid, n_hyp, d_best, d_best_pose, batch);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_dpr_rects(agt_ctx* ctx, const agt_pyramid* pyr, const double* d_init, int n_hyp, int32_t* d_rects, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!ctx->camera_set || !ctx->model_set) AGT_FAIL(ctx, AGT_ERR_NOT_READY, "agt_dpr_rects: camera and surface model must be set");
  if (!pyr || !d_init || !d_rects || n_hyp < 1 || batch < 0 || (reinterpret_cast<uintptr_t>(d_rects) & 15))
    AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_dpr_rects: bad arguments (d_rects must be 16-byte aligned)");
  if (batch == 0) return AGT_OK;
  dpr_rects_kernel<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(*pyr, ctx->cam, ctx->model.pitch, ctx->model.radius, d_init, n_hyp,
                                                                d_rects, batch);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

extern "C" int agt_any_flag(agt_ctx* ctx, const uint8_t* d_flags, int stride, uint8_t* d_out, int batch) {
  if (!ctx) return AGT_ERR_INVALID;
  if (batch == 0) return AGT_OK;          // nothing to do (empty tensors may carry null pointers)
  if (!d_flags || !d_out || stride < 1 || batch < 0) AGT_FAIL(ctx, AGT_ERR_INVALID, "agt_any_flag: bad arguments");
  if (batch == 0) return AGT_OK;
  any_flag_kernel<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(d_flags, stride, d_out, batch);
  AGT_LAUNCH_CHECK(ctx);
  return AGT_OK;
}

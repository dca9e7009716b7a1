// Reciprocal-throughput microbenchmark of the instructions the dense-refinement sample loop is made of.
// One CTA on one SM, W warps per SM sub-partition; every warp runs ITER x 32 instructions of one kind on 8
// independent dependency chains and the CTA reports cycles per warp-instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates scripts/pipe_rates.cu && ./pipe_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FFMA, FFMA2, FFMA2_BC, FMUL2, DFMA, DADD_RM, F64_F32, F32_F64, I2FP_U32, I2F_U8, I2FP_S32, DP4A, RCP, PRMT, IADD3, LDS128,
          MIX_FFMA_IADD, MIX_FFMA2_IADD, MIX_FFMA_DFMA, MIX_FFMA2_DP4A, N_OPS };
const char* kNames[N_OPS] = {"FFMA", "FFMA2", "FFMA2 (scalar bcast)", "FMUL2", "DFMA", "DADD.RM", "F2F.F64.F32", "F2F.F32.F64",
                             "I2FP.F32.U32", "I2F.U8 (byte sel)", "I2FP.F32.S32", "IDP.4A", "MUFU.RCP", "PRMT", "IADD3", "LDS.128 bcast",
                             "FFMA+IADD3 pairs", "FFMA2+IADD3 pairs", "FFMA+DFMA pairs", "FFMA2+IDP.4A pairs"};

template <int OP>
__global__ void bench(int iters, long long* out, float seed) {
  __shared__ __align__(16) double sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  float f[8]; double d[8]; unsigned long long p[8]; int n[8];
  for (int k = 0; k < 8; ++k) { f[k] = seed + k; d[k] = seed + k; p[k] = (unsigned long long)__float_as_uint(seed + k) * 0x100000001ull; n[k] = k + (int)seed; }
  const float a = seed * 0.5f, b = seed * 0.25f;
  const double da = seed * 0.5, db = seed * 0.25;
  const unsigned long long pa = (unsigned long long)__float_as_uint(a) * 0x100000001ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(a), "f"(b));
        if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(pa));
        if (OP == FFMA2_BC) asm volatile("{ .reg .b64 t; mov.b64 t, {%1, %1}; fma.rn.f32x2 %0, %0, t, t; }" : "+l"(p[k]) : "f"(a));
        if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pa));
        if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(da), "d"(db));
        if (OP == DADD_RM) asm volatile("add.rm.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(da));
        if (OP == F64_F32) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[k]) : "f"(f[k]));
        if (OP == F32_F64) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[k]) : "d"(d[k]));
        if (OP == I2FP_U32) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[k]) : "r"(n[k]));
        if (OP == I2FP_S32) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[k]) : "r"(n[k]));
        if (OP == I2F_U8) asm volatile("{ .reg .b32 t; bfe.u32 t, %1, 8, 8; cvt.rn.f32.u32 %0, t; }" : "=f"(f[k]) : "r"(n[k]));
        if (OP == DP4A) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(n[k]) : "r"(n[(k + 1) & 7]), "r"(0x000300FD));
        if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
        if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3021;" : "+r"(n[k]) : "r"(n[(k + 1) & 7]));
        if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(n[(k + 3) & 7]));
        if (OP == LDS128) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sh) + 16 * k)); d[k] += v.x; }
        if (OP == MIX_FFMA_IADD) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(a), "f"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(n[(k + 3) & 7])); }
        if (OP == MIX_FFMA2_IADD) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(pa)); asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(n[(k + 3) & 7])); }
        if (OP == MIX_FFMA_DFMA) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(a), "f"(b)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(da), "d"(db)); }
        if (OP == MIX_FFMA2_DP4A) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(pa)); asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(n[k]) : "r"(n[(k + 1) & 7]), "r"(0x000300FD)); }
      }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
  for (int k = 0; k < 8; ++k) acc += f[k] + (float)d[k] + (float)p[k] + (float)n[k];
  if (acc == 123.456f) out[1] = 1;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(long long* d_out) {
  const int iters = 2000;
  for (int w = 1; w <= 4; w *= 2) {
    bench<OP><<<1, 128 * w>>>(iters, d_out, 1.0f);
    bench<OP><<<1, 128 * w>>>(iters, d_out, 1.0f);
    long long cyc;
    cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
    const int per_iter = 32 * ((OP >= MIX_FFMA_IADD) ? 2 : 1);
    printf("%-22s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", kNames[OP], w, (double)cyc / ((double)iters * per_iter * w));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  run<FFMA>(d_out); run<FFMA2>(d_out); run<FFMA2_BC>(d_out); run<FMUL2>(d_out); run<DFMA>(d_out); run<DADD_RM>(d_out);
  run<F64_F32>(d_out); run<F32_F64>(d_out); run<I2FP_U32>(d_out); run<I2F_U8>(d_out); run<I2FP_S32>(d_out); run<DP4A>(d_out);
  run<RCP>(d_out); run<PRMT>(d_out); run<IADD3>(d_out); run<LDS128>(d_out);
  run<MIX_FFMA_IADD>(d_out); run<MIX_FFMA2_IADD>(d_out); run<MIX_FFMA_DFMA>(d_out); run<MIX_FFMA2_DP4A>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}

// How fast can SMs pull per-frame rectangles out of pinned (device-mapped) host memory?  Variants of the ROI gather of
// agt_refine_host on a synthetic rectangle list shaped like the bench workload (4096 x 1080p frames would need 8.5 GB
// of pinned memory; 1024 frames are enough for a steady state).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe scripts/gather_probe.cu && ./gather_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct Rect { int x0, y0, x1, y1, src; };

template <int PER_THREAD>
__global__ void __launch_bounds__(256) gather_chunks(const uint8_t* __restrict__ host, int w, int h, const Rect* __restrict__ rects,
                                                     uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride, int n) {
  for (int ri = blockIdx.x; ri < n; ri += gridDim.x) {
    const Rect r = rects[ri];
    const int cw = (r.x1 - r.x0) >> 4, rows = r.y1 - r.y0;
    const uint8_t* src = host + (int64_t)r.src * w * h + (int64_t)r.y0 * w + r.x0;
    uint8_t* out = dst + (int64_t)ri * dstride + (int64_t)r.y0 * dpitch + r.x0;
    const int total = cw * rows;
    for (int i = threadIdx.x; i < total; i += PER_THREAD * blockDim.x) {
      uint4 v[PER_THREAD];
#pragma unroll
      for (int k = 0; k < PER_THREAD; ++k) {
        int j = i + k * blockDim.x;
        if (j < total) { int y = j / cw, c = j - y * cw; v[k] = *reinterpret_cast<const uint4*>(src + (int64_t)y * w + 16 * c); }
      }
#pragma unroll
      for (int k = 0; k < PER_THREAD; ++k) {
        int j = i + k * blockDim.x;
        if (j < total) { int y = j / cw, c = j - y * cw; *reinterpret_cast<uint4*>(out + (int64_t)y * dpitch + 16 * c) = v[k]; }
      }
    }
  }
}

// one warp per row: every row is one contiguous run of 16-byte pieces
__global__ void __launch_bounds__(256) gather_rows(const uint8_t* __restrict__ host, int w, int h, const Rect* __restrict__ rects,
                                                   uint8_t* __restrict__ dst, int64_t dpitch, int64_t dstride, int n) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int ri = blockIdx.x; ri < n; ri += gridDim.x) {
    const Rect r = rects[ri];
    const int cw = (r.x1 - r.x0) >> 4, rows = r.y1 - r.y0;
    const uint8_t* src = host + (int64_t)r.src * w * h + (int64_t)r.y0 * w + r.x0;
    uint8_t* out = dst + (int64_t)ri * dstride + (int64_t)r.y0 * dpitch + r.x0;
    for (int y = wid; y < rows; y += 32) {          // 4 rows of this warp in flight
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (y + 8 * k < rows && lane < cw) v[k] = *reinterpret_cast<const uint4*>(src + (int64_t)(y + 8 * k) * w + 16 * lane);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (y + 8 * k < rows && lane < cw) *reinterpret_cast<uint4*>(out + (int64_t)(y + 8 * k) * dpitch + 16 * lane) = v[k];
    }
  }
}

int main() {
  const int w = 1920, h = 1080, n = 1024;
  uint8_t* hbuf;
  if (cudaHostAlloc(&hbuf, (size_t)n * w * h, cudaHostAllocMapped) != cudaSuccess) { printf("pinned alloc failed\n"); return 1; }
  for (size_t i = 0; i < (size_t)n * w * h; i += 4096) hbuf[i] = (uint8_t)i;
  uint8_t* hdev; cudaHostGetDevicePointer(&hdev, hbuf, 0);
  uint8_t* dbuf; cudaMalloc(&dbuf, (size_t)n * w * h);
  srand(7);
  for (int align = 16; align <= 128; align *= 8) {
    std::vector<Rect> rects(n);
    size_t bytes = 0;
    for (int i = 0; i < n; ++i) {
      int side = (i & 1) ? 440 + rand() % 60 : 220 + rand() % 40;        // level-1 and level-0 refinements
      int x0 = rand() % (w - side - 128), y0 = rand() % (h - side);
      int xa = x0 & ~(align - 1), xb = (x0 + side + align - 1) & ~(align - 1);
      rects[i] = {xa, y0, xb, y0 + side, i};
      bytes += (size_t)(xb - xa) * side;
    }
    Rect* drects; cudaMalloc(&drects, sizeof(Rect) * n);
    cudaMemcpy(drects, rects.data(), sizeof(Rect) * n, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 4; ++variant)
      for (int ctas = 32; ctas <= 512; ctas *= 2) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          if (variant == 0) gather_chunks<4><<<ctas, 256>>>(hdev, w, h, drects, dbuf, w, (int64_t)w * h, n);
          if (variant == 1) gather_chunks<8><<<ctas, 256>>>(hdev, w, h, drects, dbuf, w, (int64_t)w * h, n);
          if (variant == 2) gather_chunks<2><<<ctas, 256>>>(hdev, w, h, drects, dbuf, w, (int64_t)w * h, n);
          if (variant == 3) gather_rows<<<ctas, 256>>>(hdev, w, h, drects, dbuf, w, (int64_t)w * h, n);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        const char* names[4] = {"chunks x4", "chunks x8", "chunks x2", "warp per row"};
        printf("align %3d  %-13s %3d CTAs: %6.2f ms  %5.1f GB/s (%.0f MB)\n", align, names[variant], ctas, best, bytes / best / 1e6, bytes / 1e6);
      }
    cudaFree(drects);
  }
  // the copy engine on the same total, one contiguous piece
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t bytes = 100u << 20;
    cudaEventRecord(e0); cudaMemcpyAsync(dbuf, hbuf, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventRecord(e0); cudaMemcpyAsync(dbuf, hbuf, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("copy engine, contiguous %zu MB: %.2f ms  %.1f GB/s\n", bytes >> 20, ms, bytes / ms / 1e6);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

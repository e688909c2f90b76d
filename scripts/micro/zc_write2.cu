// Microbenchmark (round 2): 4096 observation rows of 227 floats into mapped (pinned) host memory: unaligned 908-B rows vs rows on
// 128-B boundaries (stride 256 floats), 16-B vs 32-B (st.global.v8.f32, sm_100) stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zc_write2 zc_write2.cu && ./zc_write2
#include <cuda_runtime.h>
#include <cstdio>
constexpr int N = 4096, D = 227;
__global__ void w16(float* out, const float* src, int stride) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float* o = out + size_t(row) * stride;
  const float* s = src + size_t(row) * D;
  const int head = (4 - int((reinterpret_cast<size_t>(o) >> 2) & 3)) & 3;
  if (lane < head) o[lane] = s[lane];
  const int nvec = (D - head) / 4;
  for (int v = lane; v < nvec; v += 32) {
    const int i = head + 4 * v;
    *reinterpret_cast<float4*>(o + i) = make_float4(s[i], s[i + 1], s[i + 2], s[i + 3]);
  }
  const int done = head + 4 * nvec;
  if (lane < D - done) o[done + lane] = s[done + lane];
}
__global__ void w32(float* out, const float* src, int stride) {  // rows must be 32-B aligned
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float* o = out + size_t(row) * stride;
  const float* s = src + size_t(row) * D;
  const int nvec = D / 8;
  if (lane < nvec) {
    const int i = 8 * lane;
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + i), "f"(s[i]), "f"(s[i + 1]), "f"(s[i + 2]), "f"(s[i + 3]),
                 "f"(s[i + 4]), "f"(s[i + 5]), "f"(s[i + 6]), "f"(s[i + 7]) : "memory");
  }
  const int done = 8 * nvec;
  if (lane < D - done) o[done + lane] = s[done + lane];
}
__global__ void w32full(float* out, const float* src, int stride) {  // whole padded row (256 floats) in one instruction per lane
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float* o = out + size_t(row) * stride;
  const float* s = src + size_t(row) * D;
  float v[8];
  for (int k = 0; k < 8; k++) v[k] = (8 * lane + k < D) ? s[8 * lane + k] : 0.f;
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 8 * lane), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
int main() {
  float *h, *dh, *src;
  cudaHostAlloc(&h, sizeof(float) * N * 256, cudaHostAllocMapped);
  cudaHostGetDevicePointer(&dh, h, 0);
  cudaMalloc(&src, sizeof(float) * N * D);
  cudaMemset(src, 1, sizeof(float) * N * D);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char* name, auto fn) {
    for (int i = 0; i < 5; i++) fn();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 50; i++) fn();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %8.1f us  %6.1f GB/s (payload)\n", name, ms / 50 * 1e3, N * D * 4.0 / (ms / 50 * 1e-3) / 1e9);
  };
  const int warps = 28, grid = (N + warps - 1) / warps;
  time("16 B stores, rows unaligned (stride 227)", [&] { w16<<<grid, warps * 32>>>(dh, src, 227); });
  time("16 B stores, rows 16-B aligned (stride 228)", [&] { w16<<<grid, warps * 32>>>(dh, src, 228); });
  time("16 B stores, rows 128-B aligned (stride 256)", [&] { w16<<<grid, warps * 32>>>(dh, src, 256); });
  time("32 B stores, rows 32-B aligned (stride 232)", [&] { w32<<<grid, warps * 32>>>(dh, src, 232); });
  time("32 B stores, rows 128-B aligned (stride 256)", [&] { w32<<<grid, warps * 32>>>(dh, src, 256); });
  time("32 B stores, whole padded row (stride 256)", [&] { w32full<<<grid, warps * 32>>>(dh, src, 256); });
  time("cudaMemcpyAsync D2H (contiguous)", [&] { cudaMemcpyAsync(h, src, sizeof(float) * N * D, cudaMemcpyDeviceToHost); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

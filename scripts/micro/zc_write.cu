// Microbenchmark: how fast can SMs stream 4096 rows of 908 B into mapped (pinned) host memory, by store width?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zc_write zc_write.cu && ./zc_write
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
constexpr int N = 4096, D = 227;
__global__ void w4(float* out, const float* src) {  // one warp per row, 4-byte stores (current kernel)
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  for (int i = lane; i < D; i += 32) __stwt(out + size_t(row) * D + i, src[size_t(row) * D + i]);
}
__global__ void w4_plain(float* out, const float* src) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  for (int i = lane; i < D; i += 32) out[size_t(row) * D + i] = src[size_t(row) * D + i];
}
__global__ void w16(float* out, const float* src) {  // one warp per row: scalar head to 16-B alignment, float4 body, scalar tail
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float* o = out + size_t(row) * D;
  const float* s = src + size_t(row) * D;
  const int head = (4 - int((reinterpret_cast<size_t>(o) >> 2) & 3)) & 3;  // floats until 16-B aligned
  if (lane < head) o[lane] = s[lane];
  const int nvec = (D - head) / 4;
  for (int v = lane; v < nvec; v += 32) {
    const int i = head + 4 * v;
    float4 t = make_float4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    *reinterpret_cast<float4*>(o + i) = t;
  }
  const int done = head + 4 * nvec;
  if (lane < D - done) o[done + lane] = s[done + lane];
}
__global__ void wflat16(float4* out, const float4* src, int n4) {  // fully coalesced flat copy, 16 B per thread
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) out[i] = src[i];
}
int main() {
  float *h, *dh, *src, *ddev;
  cudaHostAlloc(&h, sizeof(float) * N * D, cudaHostAllocMapped);
  cudaHostGetDevicePointer(&dh, h, 0);
  cudaMalloc(&src, sizeof(float) * N * D); cudaMalloc(&ddev, sizeof(float) * N * D);
  cudaMemset(src, 1, sizeof(float) * N * D);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char* name, auto fn) {
    for (int i = 0; i < 5; i++) fn();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 50; i++) fn();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s %8.1f us  %6.1f GB/s\n", name, ms / 50 * 1e3, N * D * 4.0 / (ms / 50 * 1e-3) / 1e9);
  };
  const int warps = 28, grid = (N + warps - 1) / warps;
  time("rows, 4 B stwt -> host", [&] { w4<<<grid, warps * 32>>>(dh, src); });
  time("rows, 4 B plain -> host", [&] { w4_plain<<<grid, warps * 32>>>(dh, src); });
  time("rows, 16 B body -> host", [&] { w16<<<grid, warps * 32>>>(dh, src); });
  time("flat 16 B -> host", [&] { wflat16<<<148, 256>>>((float4*)dh, (const float4*)src, N * D / 4); });
  time("rows, 4 B -> device", [&] { w4_plain<<<grid, warps * 32>>>(ddev, src); });
  time("cudaMemcpyAsync D2H", [&] { cudaMemcpyAsync(h, src, sizeof(float) * N * D, cudaMemcpyDeviceToHost); });
  return 0;
}

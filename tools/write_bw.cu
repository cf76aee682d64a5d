// write_bw.cu -- write-only HBM bandwidth on this GPU: cudaMemset, a streaming-store kernel (contiguous), and the row pattern of the
// coefficient stream (a warp writes 2 x 512 B of a 1056 B row, st.global.cs).  Build: nvcc -O3 -arch=sm_100a -o write_bw write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_fill(double2* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) __stcs(p + i, make_double2(1.0, 2.0));
}
// rows of 132 doubles: lanes write (re, im) pairs at [2 lane] and [64 + 2 lane]; 4 pad doubles stay untouched
__global__ void k_rows(double* p, size_t nrows) {
  const int lane = threadIdx.x & 31;
  size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = w; r < nrows; r += nw) {
    double* row = p + r * 132 + 2 * lane;
    __stcs(reinterpret_cast<double2*>(row), make_double2(1.0, 2.0));
    __stcs(reinterpret_cast<double2*>(row + 64), make_double2(3.0, 4.0));
  }
}
int main() {
  const size_t bytes = (size_t)4 << 30;
  void* p;
  cudaMalloc(&p, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    cudaMemsetAsync(p, 0, bytes);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemset      %.3f ms  %.0f GB/s\n", ms, bytes / ms * 1e-6);
  }
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32})
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      k_fill<<<blocks, 256>>>((double2*)p, bytes / 16);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      printf("k_fill  %5d CTAs %.3f ms  %.0f GB/s\n", blocks, ms, bytes / ms * 1e-6);
    }
  const size_t nrows = bytes / (132 * 8);
  for (int blocks : {148 * 4, 148 * 16, 148 * 32})
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      k_rows<<<blocks, 128>>>((double*)p, nrows);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      printf("k_rows  %5d CTAs %.3f ms  %.0f GB/s (written bytes: 1024 of every 1056)\n", blocks, ms, nrows * 1024.0 / ms * 1e-6);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

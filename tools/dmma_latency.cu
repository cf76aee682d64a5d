// DMMA m8n8k4 dependent-chain latency and per-SM throughput as a function of warps per SM and independent accumulator chains
// per warp (register operands).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH>
__global__ void k(double* out, int n, long long* clk) {
  double acc[NCH][2];
#pragma unroll
  for (int c = 0; c < NCH; ++c) acc[c][0] = acc[c][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) dmma884(acc[c][0], acc[c][1], a, b);
  }
  long long t1 = clock64();
  double t = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) t += acc[c][0] + acc[c][1];
  if (t == 123.456) out[0] = t;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int NCH>
void run(int nw, int nsm, double* out, long long* clk) {
  const int n = 20000;
  k<NCH><<<nsm, nw * 32>>>(out, n, clk);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
  printf("warps/SM %2d chains %d: %.1f clk per DMMA per warp, %.2f clk per DMMA per SM\n", nw, NCH, (double)c / n / NCH, (double)c / n / NCH / nw);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* out; long long* clk; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&clk, 8));
  for (int nw : {1, 4, 8, 12, 16}) {
    run<1>(nw, p.multiProcessorCount, out, clk);
    run<2>(nw, p.multiProcessorCount, out, clk);
    run<4>(nw, p.multiProcessorCount, out, clk);
    run<8>(nw, p.multiProcessorCount, out, clk);
  }
  return 0;
}

// FP64 peak micro-benchmark for B200 (sm_100a): DFMA vector pipe and DMMA (mma.sync f64) tensor path.
// Defines the roofline denominator for the Mie contraction kernels (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// m8n8k4: 256 FMA per warp instruction
template <int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
  double c0[NT], c1[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// m16n8k4: 512 FMA per warp instruction
template <int NT>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: 2048 FMA per warp instruction
template <int NT>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%4,%5,%4,%5,%4,%5}, {%6,%6,%6,%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// shared-memory broadcast LDS.128 throughput (uniform address per warp)
__global__ void __launch_bounds__(256) k_lds_bcast(double* out, int iters) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, 1.0);
  __syncthreads();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int w = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double vx, vy;
      unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(w * 64 + j * 4 + (it & 3)) & 1023]);
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
      s0 += vx; s1 += vy;
    }
  }
  if (s0 + s1 + s2 + s3 == 123.456) out[0] = s0;
}

template <typename F>
double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 8));
  int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, nsm);
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = nsm * bps;
    {
      double ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 16 * iters * 256.0 * grid;
      printf(" \"dfma_tflops_bps%d\": %.3f,\n", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { k_dmma884<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
      printf(" \"dmma_m8n8k4_tflops_bps%d\": %.3f,\n", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { k_dmma1684<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 512 * 8 * iters * 8.0 * grid;
      printf(" \"dmma_m16n8k4_tflops_bps%d\": %.3f,\n", bps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { k_dmma16816<8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
      double fl = 2.0 * 2048 * 8 * (iters / 4) * 8.0 * grid;
      printf(" \"dmma_m16n8k16_tflops_bps%d\": %.3f,\n", bps, fl / ms * 1e-9);
    }
  }
  {
    int grid = nsm * 2;
    double ms = time_ms([&] { k_lds_bcast<<<grid, 256>>>(out, iters); }, 5);
    double n = 16.0 * iters * 8.0 * grid;   // warp-level LDS.128 instructions
    printf(" \"lds128_bcast_warpinstr_per_clk_per_sm_at_1p9GHz\": %.3f,\n", n / (ms * 1e-3) / nsm / 1.9e9);
  }
  // sustained DFMA for ~2 s to see the power-capped rate
  {
    int grid = nsm * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 60; ++n) k_dfma<16><<<grid, 256>>>(out, iters * 8, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 16 * iters * 8 * 256.0 * grid * n;
    printf(" \"dfma_tflops_sustained\": %.3f, \"sustained_ms\": %.1f\n}\n", fl / ms * 1e-9, ms);
  }
  return 0;
}

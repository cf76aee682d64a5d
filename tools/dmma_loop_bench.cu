// Micro-benchmark of the k_contract inner loop structure: 12 warps/CTA, 1 CTA/SM, per k4 step 2 A-frag pairs + 8 B-frag pairs
// from shared memory, 32 DMMA m8n8k4.  Variants: operands from registers (R), from smem (S), with group epilogue (E).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int TROW = 388, LAH = 194, SB = 132, STAGE = 4 * TROW + 4 * SB, NST = 8;
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_loop(double* out, int nsteps, int nk_per_group) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < NST * STAGE; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lk = lane & 3, lr = lane >> 2, a0 = (warp % 12) * 16;
  double accp[2][8][2], accm[2][8][2], mu[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) accp[i][j][0] = accp[i][j][1] = accm[i][j][0] = accm[i][j][1] = 0.0;
    mu[i][0] = mu[i][1] = mu[i][2] = mu[i][3] = 0.0;
  }
  double rp = 1.0 + lane * 1e-9, rq = 0.5;
  for (int step = 0; step < nsteps; ++step) {
    const int s = step % NST;
    const double* tb = sm + (size_t)s * STAGE + lk * TROW + a0 + lr;
    const double* cf = sm + (size_t)s * STAGE + 4 * TROW + lk * SB + lr;
    double ap[2], aq[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ap[i] = (MODE & 1) ? tb[8 * i] : rp;
      aq[i] = (MODE & 1) ? tb[LAH + 8 * i] : rq;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double bp = (MODE & 1) ? cf[8 * j] : rq;
      const double bm = (MODE & 1) ? cf[64 + 8 * j] : rp;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dmma884(accp[i][j][0], accp[i][j][1], ap[i], bp);
        dmma884(accm[i][j][0], accm[i][j][1], aq[i], bm);
      }
    }
    if ((MODE & 2) && (step % nk_per_group) == nk_per_group - 1) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double pr = accp[i][j][0], pi = accp[i][j][1], mr = accm[i][j][0], mi = accm[i][j][1];
          mu[i][0] = fma(pr, pr, fma(pi, pi, mu[i][0]));
          mu[i][1] = fma(mr, mr, fma(mi, mi, mu[i][1]));
          mu[i][2] = fma(pr, mr, fma(pi, mi, mu[i][2]));
          mu[i][3] = fma(pi, mr, fma(-pr, mi, mu[i][3]));
          accp[i][j][0] = accp[i][j][1] = accm[i][j][0] = accm[i][j][1] = 0.0;
        }
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) t += accp[i][j][0] + accp[i][j][1] + accm[i][j][0] + accm[i][j][1];
    t += mu[i][0] + mu[i][1] + mu[i][2] + mu[i][3];
  }
  if (t == 123.456) out[0] = t;
}
template <int MODE, int NW>
void run(const char* name, double* out, int nsm, int nk) {
  const int smem = NST * STAGE * 8;
  CK(cudaFuncSetAttribute(k_loop<MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int nsteps = 20000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_loop<MODE, NW><<<nsm, NW * 32, smem>>>(out, nsteps, nk);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_loop<MODE, NW><<<nsm, NW * 32, smem>>>(out, nsteps, nk);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  double fl = 2.0 * 256 * 32 * NW * (double)nsteps * nsm;
  printf("%-40s NW=%2d nk=%d  %.2f TFLOP/s (DMMA only)  %.1f clk/step @1.965GHz\n", name, NW, nk, fl / ms * 1e-9, ms * 1e-3 * 1.965e9 / nsteps);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* out; CK(cudaMalloc(&out, 8));
  run<0, 12>("regs operands", out, p.multiProcessorCount, 1000000);
  run<1, 12>("smem operands", out, p.multiProcessorCount, 1000000);
  run<3, 12>("smem operands + epilogue every 2 steps", out, p.multiProcessorCount, 2);
  run<3, 12>("smem operands + epilogue every 3 steps", out, p.multiProcessorCount, 3);
  run<2, 12>("regs operands + epilogue every 2 steps", out, p.multiProcessorCount, 2);
  run<0, 8>("regs operands", out, p.multiProcessorCount, 1000000);
  run<1, 8>("smem operands", out, p.multiProcessorCount, 1000000);
  run<0, 4>("regs operands", out, p.multiProcessorCount, 1000000);
  run<1, 4>("smem operands", out, p.multiProcessorCount, 1000000);
  return 0;
}

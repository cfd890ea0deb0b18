// Dependent-chain latencies and single-warp issue rates of the instructions on the solver's pivot chain (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
__global__ void k_dfma_dep(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_indep(double* out, long long* cyc, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_dep(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp_dep(double* out, long long* cyc, double a) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_div_dep(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) x = b / x;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_dep(double* out, long long* cyc) {
  __shared__ int idx[64];
  for (int i = threadIdx.x; i < 64; i += 32) idx[i] = (i + 1) & 63;
  __syncwarp();
  int x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = idx[x];
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_stslds_dep(double* out, long long* cyc, double a) {
  __shared__ double buf[64];
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N_IT; ++i) { buf[threadIdx.x] = x; __syncwarp(); x = buf[(threadIdx.x + 1) & 31]; __syncwarp(); }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_imad_dep(double* out, long long* cyc, int a, int b) {
  int x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = x * b + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma_dep(double* out, long long* cyc, float a, float b) {
  float x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = fmaf(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  long long h;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock rate attr %d kHz\n", clk);
#define RUN(name, launch, per) for (int r = 0; r < 2; ++r) { launch; cudaDeviceSynchronize(); } cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %8.2f cycles per op\n", name, (double)h / (N_IT * (per)));
  RUN("DFMA dependent (1 warp)", (k_dfma_dep<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999)), 1)
  RUN("DFMA 8 independent (1 warp)", (k_dfma_indep<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999)), 8)
  RUN("DFMA 8 indep (4 warps/SM)", (k_dfma_indep<<<1, 128>>>(out, cyc, 1.0000001, 0.9999999)), 8)
  RUN("DFMA 8 indep (16 warps/SM)", (k_dfma_indep<<<1, 512>>>(out, cyc, 1.0000001, 0.9999999)), 8)
  RUN("SHFL.64 dependent", (k_shfl_dep<<<1, 32>>>(out, cyc, 1.0)), 1)
  RUN("rcp.approx.f64 dependent", (k_rcp_dep<<<1, 32>>>(out, cyc, 1.5)), 1)
  RUN("fp64 divide dependent", (k_div_dep<<<1, 32>>>(out, cyc, 1.5, 2.5)), 1)
  RUN("LDS dependent", (k_lds_dep<<<1, 32>>>(out, cyc)), 1)
  RUN("STS+sync+LDS+sync round trip", (k_stslds_dep<<<1, 32>>>(out, cyc, 1.0)), 1)
  RUN("IMAD dependent", (k_imad_dep<<<1, 32>>>(out, cyc, 3, 5)), 1)
  RUN("FFMA dependent", (k_ffma_dep<<<1, 32>>>(out, cyc, 1.0000001f, 0.9999999f)), 1)
  return 0;
}

// DFMA latency / throughput microbenchmark (B200): nvcc -gencode arch=compute_100a,code=sm_100a -O3 dfma.cu -o dfma && ./dfma
// Results used in DESIGN.md section 3: dependent DFMA issue interval 9.4 cycles, peak 1 DFMA / 2 cycles / SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double seed) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 123.456) out[blockIdx.x] = s;
}
// mixed: DFMA chain with an independent integer instruction stream (IMAD) to see co-issue cost
template <int ILP, int NINT>
__global__ void kmix(double* out, int iters, double seed) {
  double a[ILP];
  unsigned u = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
#pragma unroll
    for (int j = 0; j < NINT; ++j) u = u * 1664525u + 1013904223u;
  }
  double s = u;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 123.456) out[blockIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 1 << 20);
  int sms = 148, iters = 1 << 15;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock %d kHz\n", clk);
  printf("warps/SM ILP  TFLOP/s  cycles_per_dfma_per_warp(approx @1.965GHz)\n");
  int wps[] = {4, 8, 12, 16, 32, 64};
  for (int w : wps) {
#define RUN(I) { float ms = timeit([&]{ k<I><<<sms, w*32>>>(out, iters, 1.0); }); \
    double fl = (double)sms*w*32*(double)iters*I*2; double cyc = ms*1e-3*1.965e9/((double)iters*I)*(4.0/ (w<4?w:4)) ; \
    printf("%3d %3d  %7.2f   issue-interval %.2f cyc/DFMA/SMSP-warpslot\n", w, I, fl/(ms*1e-3)/1e12, ms*1e-3*1.965e9/((double)iters*I*((w+3)/4))); }
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(8)
  }
  printf("mix: 12 warps/SM, ILP2, + N int ops per DFMA pair\n");
  { float ms = timeit([&]{ kmix<2,0><<<sms, 12*32>>>(out, iters, 1.0); }); printf("  nint 0: %.3f ms\n", ms); }
  { float ms = timeit([&]{ kmix<2,2><<<sms, 12*32>>>(out, iters, 1.0); }); printf("  nint 2: %.3f ms\n", ms); }
  { float ms = timeit([&]{ kmix<2,4><<<sms, 12*32>>>(out, iters, 1.0); }); printf("  nint 4: %.3f ms\n", ms); }
  { float ms = timeit([&]{ kmix<2,8><<<sms, 12*32>>>(out, iters, 1.0); }); printf("  nint 8: %.3f ms\n", ms); }
  return 0;
}

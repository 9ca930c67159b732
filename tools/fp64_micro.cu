// FP64 pipe microbenchmarks for the fused-pass inner loop (developer tool).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_micro fp64_micro.cu && ./fp64_micro
// Reports warp-instructions per clock per SM for instruction mixes at several occupancies.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE, int ILP>
__global__ void mix_kernel(int iters, double* sink, double seed) {
  double a[ILP], b[ILP];
  unsigned c[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) { a[k] = seed + threadIdx.x * 1e-9 + k; b[k] = seed * 0.5 + k; c[k] = 0; }
  const double m = 1.0000000001, t = 3.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        if (MODE == 0) a[k] = fma(a[k], m, b[k]);                 // DFMA
        if (MODE == 1) a[k] = __dadd_rn(a[k], b[k]);              // DADD
        if (MODE == 2) a[k] = __dmul_rn(a[k], m);                 // DMUL
        if (MODE == 3) {                                          // DSETP + predicated int add
          asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(c[k]) : "d"(a[k]), "d"(t));
        }
        if (MODE == 4) {                                          // DADD + DSETP + select-DFMA (the classification tail)
          a[k] = __dadd_rn(a[k], m);
          asm volatile("{.reg .pred p; setp.ge.f64 p, %2, %3; @p fma.rn.f64 %0, %2, %2, %0; @p add.u32 %1, %1, 1;}"
                       : "+d"(b[k]), "+r"(c[k]) : "d"(a[k]), "d"(t));
        }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) r += a[k] + b[k] + c[k];
  if (r == 12345.678) sink[0] = r;
}

// the W=2 window loop of pair_hist_win_kernel, HRx i-points per thread, j from shared memory
template <int HRX, int UNROLL>
__global__ void pair_loop_kernel(int iters, double* sink, double seed) {
  __shared__ double2 sXY[256];
  __shared__ double sZ[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) { sXY[k] = make_double2(seed * k, seed * (k ^ 5)); sZ[k] = seed + k; }
  __syncthreads();
  double xi[HRX], yi[HRX], zi[HRX], T1[HRX], q0[HRX], q1[HRX];
  unsigned c1[HRX];
#pragma unroll
  for (int r = 0; r < HRX; ++r) {
    xi[r] = seed * (threadIdx.x + 64 * r); yi[r] = seed * (threadIdx.x * 3 + r); zi[r] = seed + r;
    T1[r] = seed * seed * 1e4 * (r + 1); q0[r] = q1[r] = 0; c1[r] = 0;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll UNROLL
    for (int jj = 0; jj < 256; ++jj) {
      const double2 pj = sXY[jj];
      const double zj = sZ[jj];
#pragma unroll
      for (int r = 0; r < HRX; ++r) {
        const double dx = xi[r] - pj.x;
        const double dy = yi[r] - pj.y;
        const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        const double dz = zi[r] - zj;
        q0[r] = fma(dz, dz, q0[r]);
        asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %3, %4;\n\t@p fma.rn.f64 %0, %2, %2, %0;\n\t@p add.u32 %1, %1, 1;\n\t}"
            : "+d"(q1[r]), "+r"(c1[r]) : "d"(dz), "d"(s), "d"(T1[r]));
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < HRX; ++k) r += q0[k] + q1[k] + c1[k];
  if (r == 12345.678) sink[0] = r;
}

template <class K>
static double run(K kern, int blocks, int threads, int iters, double* sink) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<<<blocks, threads>>>(iters / 4 + 1, sink, 1.25);
  CK(cudaEventRecord(e0));
  kern<<<blocks, threads>>>(iters, sink, 1.25);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  double* sink;
  CK(cudaMalloc(&sink, 8));
  const double clk = khz * 1e3;
  printf("SMs %d, clock %.0f MHz (nominal; rates below assume it)\n", sms, clk / 1e6);
  const char* names[5] = {"DFMA", "DADD", "DMUL", "DSETP+@IADD", "DADD+DSETP+selDFMA"};
  const int per[5] = {1, 1, 1, 1, 3};  // FP64-pipe instructions per (k, u) step
  for (int warps : {4, 8, 16, 32}) {
    const int threads = 128, blocks = sms * (warps * 32 / threads);
    const int iters = 2000;
    double ms[5];
    ms[0] = run(mix_kernel<0, 4>, blocks, threads, iters, sink);
    ms[1] = run(mix_kernel<1, 4>, blocks, threads, iters, sink);
    ms[2] = run(mix_kernel<2, 4>, blocks, threads, iters, sink);
    ms[3] = run(mix_kernel<3, 4>, blocks, threads, iters, sink);
    ms[4] = run(mix_kernel<4, 4>, blocks, threads, iters, sink);
    for (int m = 0; m < 5; ++m) {
      const double winst = (double)blocks * (threads / 32) * iters * 8.0 * 4 * per[m];
      printf("warps/SM %2d ILP4 %-20s %.3f ms  FP64 warp-inst/clk/SM %.3f\n", warps, names[m], ms[m],
             winst / (ms[m] * 1e-3 * clk) / sms);
    }
  }
  printf("pair loop (9 FP64-pipe instructions per pair; ideal 2.0 warp-inst/clk/SM = 7.11 pairs/clk/SM)\n");
  auto report = [&](const char* nm, double ms, int blocks, int threads, int hrx, int iters) {
    const double pairs = (double)blocks * threads * hrx * 256.0 * iters;
    printf("%-26s %.3f ms  pairs/clk/SM %.3f  (FP64 warp-inst/clk/SM %.3f)\n", nm, ms,
           pairs / (ms * 1e-3 * clk) / sms, pairs / 32 * 9 / (ms * 1e-3 * clk) / sms);
  };
  const int it = 40;
  for (int cta : {2, 4, 6, 8}) {
    char nm[64];
    snprintf(nm, 64, "64thr x4pts u2, %d CTA/SM", cta);
    report(nm, run(pair_loop_kernel<4, 2>, sms * cta, 64, it, sink), sms * cta, 64, 4, it);
    snprintf(nm, 64, "128thr x2pts u4, %d CTA/SM", cta);
    report(nm, run(pair_loop_kernel<2, 4>, sms * cta, 128, it, sink), sms * cta, 128, 2, it);
    snprintf(nm, 64, "256thr x1pt u8, %d CTA/SM", cta);
    report(nm, run(pair_loop_kernel<1, 8>, sms * cta, 256, it, sink), sms * cta, 256, 1, it);
  }
  return 0;
}

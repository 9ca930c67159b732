// Error plumbing, device queries and the DFMA roofline microbenchmark.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace skg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
    return n;
  }
  cudaGetLastError();  // clear
  return 148;          // B200; used only for sizing queries on a box without a GPU
}

// 8 independent DFMA chains per thread: saturates the FP64 pipe of every SMSP.
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678) sink[0] = r;
}

// Z-order key of every point: quantise each coordinate to `bits` bits over its range and
// interleave (dim 2: 31 bits each, dim 3: 21 bits each).
__global__ void morton_keys_kernel(const double* __restrict__ coords, int64_t n, int dim,
                                   double lo0, double lo1, double lo2, double sc0, double sc1,
                                   double sc2, unsigned long long* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bits = dim == 3 ? 21 : 31;
  const long long top = (1ll << bits) - 1;
  unsigned long long key = 0ull;
  for (int k = 0; k < dim; ++k) {
    const double x = coords[(int64_t)k * n + i];
    const double lo = k == 0 ? lo0 : (k == 1 ? lo1 : lo2);
    const double sc = k == 0 ? sc0 : (k == 1 ? sc1 : sc2);
    double q = (x - lo) * sc;
    long long v = (q == q) ? (long long)fmin(fmax(q, 0.0), (double)top) : 0ll;  // NaN -> cell 0
    unsigned long long u = (unsigned long long)v;
    if (dim == 2) {
      u = (u | (u << 16)) & 0x0000FFFF0000FFFFull;
      u = (u | (u << 8)) & 0x00FF00FF00FF00FFull;
      u = (u | (u << 4)) & 0x0F0F0F0F0F0F0F0Full;
      u = (u | (u << 2)) & 0x3333333333333333ull;
      u = (u | (u << 1)) & 0x5555555555555555ull;
    } else {
      u = (u | (u << 32)) & 0x1F00000000FFFFull;
      u = (u | (u << 16)) & 0x1F0000FF0000FFull;
      u = (u | (u << 8)) & 0x100F00F00F00F00Full;
      u = (u | (u << 4)) & 0x10C30C30C30C30C3ull;
      u = (u | (u << 2)) & 0x1249249249249249ull;
    }
    key |= u << k;
  }
  keys[i] = key;
}

}  // namespace skg

using namespace skg;

extern "C" {

int skg_morton_keys(const double* coords, int64_t n, int dim, const double* lo, const double* ext,
                    uint64_t* keys, void* stream) {
  if (!coords || !lo || !ext || !keys || n < 0) { set_error("NULL argument"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  if (n == 0) return 0;
  const double top = (double)((1ll << (dim == 3 ? 21 : 31)) - 1);
  double sc[3] = {0.0, 0.0, 0.0}, l[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < dim; ++k) {
    l[k] = lo[k];
    sc[k] = ext[k] > 0.0 ? top / ext[k] : 0.0;
  }
  const int threads = 256;
  morton_keys_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      coords, n, dim, l[0], l[1], l[2], sc[0], sc[1], sc[2],
      reinterpret_cast<unsigned long long*>(keys));
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int skg_abi_version(void) { return SKG_ABI_VERSION; }

const char* skg_last_error(void) { return g_err; }

int skg_device_sm_count(int* out) {
  if (!out) { set_error("NULL"); return SKG_E_BADARG; }
  int dev = 0, n = 0;
  SKG_CUDA(cudaGetDevice(&dev));
  SKG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return 0;
}

int skg_fp64_peak(int iters, double* ms, double* fma_count) {
  if (iters < 1 || !ms || !fma_count) { set_error("bad arguments"); return SKG_E_BADARG; }
  int sms = 0;
  int rc = skg_device_sm_count(&sms);
  if (rc) return rc;
  double* sink = nullptr;
  SKG_CUDA(cudaMalloc((void**)&sink, 8));
  cudaEvent_t e0, e1;
  SKG_CUDA(cudaEventCreate(&e0));
  SKG_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  dfma_peak_kernel<<<blocks, threads>>>(iters / 8 + 1, sink);  // warm-up
  SKG_CUDA(cudaEventRecord(e0));
  dfma_peak_kernel<<<blocks, threads>>>(iters, sink);
  SKG_CUDA(cudaEventRecord(e1));
  SKG_CUDA(cudaEventSynchronize(e1));
  float t = 0.f;
  SKG_CUDA(cudaEventElapsedTime(&t, e0, e1));
  *ms = t;
  *fma_count = (double)blocks * threads * (double)iters * 16.0 * 8.0;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

}  // extern "C"

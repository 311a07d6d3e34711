// kab_debug.h -- development builds only (-DKAB_BAND_TIMING, -DKAB_BANDP_TIMING, -DKAB_BANDQ_TIMING,
// -DKAB_BANDR_TIMING, -DKAB_WIDE_TIMING): each kernel's timing build writes cycle
// counters to KabParams::debug; the helpers below own that buffer for one launch (constructor:
// allocate once, clear, hook it into the parameters; destructor: wait for the stream and print to
// stderr).  In a normal build they are empty objects and the kernels never see a debug pointer.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "kab_common.cuh"

namespace kab_debug {
inline long long *buffer(long long *&slot, size_t n, cudaStream_t s) {
  if (!slot) cudaMalloc((void **)&slot, n * sizeof(long long));
  cudaMemsetAsync(slot, 0, n * sizeof(long long), s);
  return slot;
}
}  // namespace kab_debug

// ---- kab_band_kernel (single CTA)
struct KabBandTiming {
#ifdef KAB_BAND_TIMING
  long long *d; cudaStream_t s; int nw;
  KabBandTiming(KabParams &p, cudaStream_t stream, int band_nw) : s(stream), nw(band_nw) {
    static long long *slot = nullptr;
    d = kab_debug::buffer(slot, 32 * 8, s);
    p.debug = d;
  }
  ~KabBandTiming() {
    long long h[32 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int w = 0; w < nw; ++w)
      fprintf(stderr, "warp %2d: fast %lld cyc / %lld groups, slow %lld / %lld, epi %lld, bar %lld, post %lld, total %lld\n", w,
              h[w * 8 + 0], h[w * 8 + 1], h[w * 8 + 2], h[w * 8 + 3], h[w * 8 + 4], h[w * 8 + 5], h[w * 8 + 6], h[w * 8 + 7]);
  }
#else
  KabBandTiming(KabParams &, cudaStream_t, int) {}
#endif
};

// ---- kab_bandp_kernel
struct KabBandpTiming {
#ifdef KAB_BANDP_TIMING
  long long *d; cudaStream_t s; int nw;
  KabBandpTiming(KabParams &p, cudaStream_t stream, int n_warps) : s(stream), nw(n_warps) {
    static long long *slot = nullptr;
    d = kab_debug::buffer(slot, 32 * 16 + 8, s);
    p.debug = d;
  }
  ~KabBandpTiming() {
    long long h[32 * 16 + 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int w = 0; w < nw; ++w) {
      const long long *x = h + w * 16;
      const double n = (double)(x[7] ? x[7] : 1);
      fprintf(stderr, "warp %2d: per group: ghost %5.0f (guard %4.0f) emis %4.0f comp %5.0f pub %4.0f rel %4.0f bp %4.0f | total %lld cyc, %lld groups, need %lld, safe %lld, wait/need %.0f, first-try %lld, comp safe %.0f slow %.0f\n",
              w, x[0] / n, x[8] / n, x[1] / n, x[2] / n, x[3] / n, x[4] / n, x[5] / n, x[6], x[7], x[9], x[10], x[9] ? (double)x[11] / x[9] : 0.0, x[12],
              x[10] ? (double)(x[2] - x[13]) / x[10] : 0.0, x[7] - x[10] ? (double)x[13] / (x[7] - x[10]) : 0.0);
    }
    fprintf(stderr, "backtrack %lld cyc\n", h[32 * 16]);
  }
#else
  KabBandpTiming(KabParams &, cudaStream_t, int) {}
#endif
};

// ---- kab_bandq_kernel
struct KabBandqTiming {
#ifdef KAB_BANDQ_TIMING
  long long *d; cudaStream_t s; int nw;
  KabBandqTiming(KabParams &p, cudaStream_t stream, int n_warps) : s(stream), nw(n_warps) {
    static long long *slot = nullptr;
    d = kab_debug::buffer(slot, 64 * 28 + 8, s);
    p.debug = d;
  }
  ~KabBandqTiming() {
    static long long h[64 * 28 + 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long ct[4] = {0, 0, 0, 0}, cn[4] = {0, 0, 0, 0}, cw[4] = {0, 0, 0, 0};
    for (int w = 0; w < nw; ++w)
      for (int k = 0; k < 4; ++k) {
        const long long *c = h + 64 * 16 + 8 + w * 12;
        ct[k] += c[k]; cn[k] += c[4 + k]; cw[k] += c[8 + k];
      }
    const char *names[4] = {"free / head (no message, edge body)", "no message, safe body", "message + edge body", "message + safe body"};
    for (int k = 0; k < 4; ++k)
      fprintf(stderr, "groups [%s]: %lld, %.0f cycles each, of which waiting %.0f\n", names[k], cn[k],
              cn[k] ? (double)ct[k] / cn[k] : 0.0, cn[k] ? (double)cw[k] / cn[k] : 0.0);
    for (int w = 0; w < nw; w += 9) {
      const long long *x = h + w * 16;
      const double n = (double)(x[7] ? x[7] : 1);
      fprintf(stderr, "warp %2d: per group: ghost %5.0f emis %4.0f comp %5.0f pub %4.0f bp %4.0f rel %4.0f | total %lld cyc = %.0f / group, %lld groups, need %lld, safe %lld, wait/need %.0f, comp safe %.0f slow %.0f\n",
              w, x[0] / n, x[1] / n, x[2] / n, x[3] / n, x[5] / n, x[4] / n, x[6], x[6] / n, x[7], x[9], x[10], x[9] ? (double)x[11] / x[9] : 0.0,
              x[10] ? (double)(x[2] - x[13]) / x[10] : 0.0, x[7] - x[10] ? (double)x[13] / (x[7] - x[10]) : 0.0);
    }
  }
#else
  KabBandqTiming(KabParams &, cudaStream_t, int) {}
#endif
};

// ---- kab_bandr_kernel: per-warp cycle counters and, with KAB_TRACE_FILE=path, the per-pass trace
// (start, waits, frames of every two-group pass of every compute warp) as raw int64 for analysis
struct KabBandrTiming {
#ifdef KAB_BANDR_TIMING
  long long *d; cudaStream_t s; int nw;
  static constexpr size_t N = 64 * 26 + (size_t)64 * 16384 * 2;
  KabBandrTiming(KabParams &p, cudaStream_t stream, int n_warps) : s(stream), nw(n_warps) {
    static long long *slot = nullptr;
    d = kab_debug::buffer(slot, N, s);
    p.debug = d;
  }
  ~KabBandrTiming() {
    static long long h[64 * 26];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (const char *tf = getenv("KAB_TRACE_FILE")) {
      std::vector<long long> tr((size_t)64 * 16384 * 2);
      cudaMemcpy(tr.data(), d + 64 * 26, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      if (FILE *f = fopen(tf, "wb")) { fwrite(tr.data(), sizeof(long long), tr.size(), f); fclose(f); }
    }
    long long a[16] = {0};
    for (int w = 0; w < nw; ++w)
      for (int k = 0; k < 15; ++k) a[k] += h[w * 16 + k];
    const double n = (double)(a[6] ? a[6] : 1);
    fprintf(stderr, "compute warps, per group: tile %.0f msg %.0f (waiting %.0f) frames %.0f pub %.0f bp %.0f | total %.0f cycles / group\n",
            a[0] / n, a[1] / n, a[8] / n, a[2] / n, a[3] / n, a[4] / n, a[5] / n);
    long long ft = 0, fn = 0;
    for (int w = 0; w < nw; ++w) { ft += h[64 * 24 + w * 2]; fn += h[64 * 24 + w * 2 + 1]; }
    fprintf(stderr, "  common-path groups: %lld of %lld, %.0f cycles each; the others %.0f cycles each\n", fn, a[6], fn ? (double)ft / fn : 0.0,
            a[6] - fn ? (double)(a[5] - ft) / (a[6] - fn) : 0.0);
    long long b[8] = {0};
    for (int w = 0; w < nw; ++w)
      for (int k = 0; k < 6; ++k) b[k] += h[64 * 16 + w * 8 + k];
    const double m = (double)(b[5] ? b[5] : 1);
    fprintf(stderr, "prep warps, per group: waiting for the slot %.0f, tile %.0f, stage / backpointers %.0f | total %.0f, safe groups %.0f %%\n",
            b[0] / m, b[1] / m, b[2] / m, b[3] / m, 100.0 * b[4] / m);
  }
#else
  KabBandrTiming(KabParams &, cudaStream_t, int) {}
#endif
};

// ---- kab_wide_kernel
struct KabWideTiming {
#ifdef KAB_WIDE_TIMING
  long long *d; cudaStream_t s;
  KabWideTiming(KabParams &p, cudaStream_t stream) : s(stream) {
    static long long *slot = nullptr;
    d = kab_debug::buffer(slot, 64, s);
    p.debug = d;
  }
  ~KabWideTiming() {
    long long h[64];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[6] = {"w0", "w1", "w2", "w3", "mid", "last"};
    for (int w = 0; w < 6; ++w) {
      const long long *x = h + w * 8;
      const double n = (double)(x[7] ? x[7] : 1);
      fprintf(stderr, "%4s: per group: ghost %6.0f emis %5.0f comp %6.0f pub %6.0f rest %5.0f | prefetch misses %lld of %lld groups, total %lld cyc\n",
              nm[w], x[0] / n, x[1] / n, x[2] / n, x[3] / n, x[4] / n, x[5], x[7], x[6]);
    }
  }
#else
  KabWideTiming(KabParams &, cudaStream_t) {}
#endif
};

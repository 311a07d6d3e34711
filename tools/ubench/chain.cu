// Latency / throughput of the instructions on the band kernels' recurrence chain (sm_100a):
// SHFL.UP, FADD, FADD2 (add.rn.f32x2), FMNMX, FMNMX3, and the chain SHFL -> FADD2 -> FMNMX3 itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/chain.bin tools/ubench/chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../kokoro-align_b200/csrc/kab_common.cuh"

#define N 4096

__device__ __forceinline__ void add2(float lo, float hi, float e, float &olo, float &ohi) {
  asm volatile("{\n\t.reg .b64 u, v, w;\n\tmov.b64 u, {%2, %3};\n\tmov.b64 v, {%4, %4};\n\tadd.rn.f32x2 w, u, v;\n\tmov.b64 {%0, %1}, w;\n\t}"
               : "=f"(olo), "=f"(ohi) : "f"(lo), "f"(hi), "f"(e));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int MODE>
__global__ void k(float *out, long long *cyc, float e, int active_warps) {
  const int warp = threadIdx.x >> 5;
  float s0 = threadIdx.x * 0.001f, s1 = s0 + 1.f, t0 = 0, t1 = 0;
  uint32_t bw = 0, acc = 0;
  const uint32_t one = (uint32_t)active_warps > 0 ? 1u : 0u;
  __syncthreads();
  long long c0 = clock64();
  if (warp < active_warps) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
      if (MODE == 0) {  // SHFL.UP chain
        s0 = __shfl_up_sync(0xffffffffu, s0, 1);
      } else if (MODE == 1) {  // FADD chain
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s0) : "f"(e));
      } else if (MODE == 2) {  // FADD2 chain
        add2(s0, s1, e, s0, s1);
      } else if (MODE == 3) {  // FMNMX chain
        asm volatile("max.f32 %0, %0, %1;" : "+f"(s0) : "f"(s1));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(s1) : "f"(s0));
      } else if (MODE == 4) {  // FMNMX3 chain
        s0 = max3(s0, s1, e);
        s1 = max3(s1, s0, e);
      } else if (MODE == 5) {  // the recurrence chain: SHFL -> FADD2 -> FMNMX3
        const float h = __shfl_up_sync(0xffffffffu, s1, 1);
        add2(s0, h, e, t0, t1);
        s1 = max3(t0, t1, s1);
      } else if (MODE == 6) {  // two frames of the 2-states-per-lane cell update (values only)
        const float h1 = __shfl_up_sync(0xffffffffu, s1, 1);
        const float h2 = __shfl_up_sync(0xffffffffu, s0, 1);
        const float h3 = __shfl_up_sync(0xffffffffu, s1, 2);
        float a0, a1, a2, a3, th3;
        add2(s0, h1, e, t0, t1);
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(th3) : "f"(h3), "f"(e));
        add2(s0, s1, e, a1, a0);
        add2(h2, h1, e, a3, a2);
        s0 = max3(t0, t1, th3);
        float m01;
        asm volatile("max.f32 %0, %1, %2;" : "=f"(m01) : "f"(a0), "f"(a1));
        s1 = max3(m01, a2, a3);
      } else if (MODE == 7) {  // 8 independent SHFL.UP per iteration (throughput)
        s0 = __shfl_up_sync(0xffffffffu, s0, 1);
        s1 = __shfl_up_sync(0xffffffffu, s1, 1);
        t0 = __shfl_up_sync(0xffffffffu, t0, 1);
        t1 = __shfl_up_sync(0xffffffffu, t1, 1);
      } else if (MODE == 9) {  // the real two-state frame of kab_bandq / kab_bandr: values + backpointer bits
        const float h1 = __shfl_up_sync(0xffffffffu, s1, 1);
        const float h2 = __shfl_up_sync(0xffffffffu, s0, 1);
        const float h3 = __shfl_up_sync(0xffffffffu, s1, 2);
        float a0, a1, a2, a3, x0, x1;
        kab_add2(s0, h1, e, x0, x1);
        const float th3 = __fadd_rn(h3, e);
        kab_add2(s0, s1, e, a1, a0);
        kab_add2(h2, h1, e, a3, a2);
        const int sh = 4 * (i & 7);
        s0 = kab_blank_sel(x0, x1, th3, bw, 1u << (sh + 0), 2u << (sh + 0), one);
        s1 = kab_label_sel(a0, a1, a2, a3, bw, 1u << (sh + 2), 2u << (sh + 2), one);
        if ((i & 7) == 7) { acc ^= bw; bw = 0; }
      } else if (MODE == 8) {  // shared-memory exchange instead of SHFL: STS -> (warp sync) -> LDS
        extern __shared__ float sm[];
        sm[threadIdx.x + 1] = s1;
        __syncwarp();
        const float h = sm[threadIdx.x];
        __syncwarp();
        add2(s0, h, e, t0, t1);
        s1 = max3(t0, t1, s1);
      }
    }
  }
  long long c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + t0 + t1 + (float)acc;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + warp] = c1 - c0;
}

template <int MODE>
void run(const char *name, int per_iter, int warps, int active) {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  k<MODE><<<1, warps * 32, 8192>>>(out, cyc, 0.5f, active);
  cudaDeviceSynchronize();
  k<MODE><<<1, warps * 32, 8192>>>(out, cyc, 0.5f, active);
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
  printf("%-44s warps %2d: %7.1f cycles / iteration (%d dependent steps)  [%s]\n", name, active, (double)h[0] / N, per_iter, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

// One frame warp per scheduler (warps 8..11, the highest ids) next to two helper warps per scheduler
// (warps 0..7) that poll a shared-memory word the way the prep warps of kab_bandr.cuh wait for a tile
// slot: NOISE 0 = helpers exit, 1 = tight LDS poll, 2 = poll + __nanosleep(100), 3 = poll + __nanosleep(1000),
// 4 = helpers build tiles continuously (16 LDS + 8 STS.64 per iteration)
template <int NOISE>
__global__ void kn(float *out, long long *cyc, float e, volatile int *flag, int ninf_lanes = 0) {
  __shared__ int word;
  __shared__ __align__(16) float buf[4096];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) word = 0;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
  __syncthreads();
  if (warp >= 8) {
    float s0 = threadIdx.x * 0.001f, s1 = s0 + 1.f;
    uint32_t bw = 0, acc = 0;
    const uint32_t one = e > 0 ? 1u : 0u;
    // lanes below ninf_lanes: states outside the window (scores and masked emissions -inf), as at the
    // band's edges and in a warp whose ring slots are outside the window
    if ((int)(threadIdx.x & 31) < ninf_lanes) { s0 = s1 = e = -__int_as_float(0x7f800000); }
    const long long c0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
      const float h1 = __shfl_up_sync(0xffffffffu, s1, 1);
      const float h2 = __shfl_up_sync(0xffffffffu, s0, 1);
      const float h3 = __shfl_up_sync(0xffffffffu, s1, 2);
      float a0, a1, a2, a3, x0, x1;
      kab_add2(s0, h1, e, x0, x1);
      const float th3 = __fadd_rn(h3, e);
      kab_add2(s0, s1, e, a1, a0);
      kab_add2(h2, h1, e, a3, a2);
      const int sh = 4 * (i & 7);
      s0 = kab_blank_sel(x0, x1, th3, bw, 1u << (sh + 0), 2u << (sh + 0), one);
      s1 = kab_label_sel(a0, a1, a2, a3, bw, 1u << (sh + 2), 2u << (sh + 2), one);
      if ((i & 7) == 7) { acc ^= bw; bw = 0; }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = s0 + s1 + (float)acc;
    if ((threadIdx.x & 31) == 0) cyc[warp - 8] = c1 - c0;
    __syncwarp();
    if (threadIdx.x == 8 * 32) { __threadfence_block(); *(volatile int *)&word = 1; }
  } else if (NOISE > 0) {
    float acc = 0;
    if (NOISE == 4) {
      float2 *tile = reinterpret_cast<float2 *>(buf) + (threadIdx.x & 31) + (warp & 1) * 1024;
      while (*(volatile int *)&word == 0) {
#pragma unroll
        for (int f = 0; f < 8; ++f) tile[f * 32] = make_float2(buf[2048 + f * 39 + (threadIdx.x & 31)], buf[2048 + f * 39 + 7]);
        __syncwarp();
      }
    } else {
      while (*(volatile int *)&word == 0) {
        acc += buf[threadIdx.x];
        if (NOISE == 2) __nanosleep(100);
        if (NOISE == 3) __nanosleep(1000);
      }
    }
    out[threadIdx.x] = acc;
  }
}
template <int NOISE>
void runn(const char *name, int ninf_lanes = 0) {
  float *out; long long *cyc; int *flag;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64 * 8); cudaMalloc(&flag, 4);
  kn<NOISE><<<1, 12 * 32>>>(out, cyc, 0.5f, flag, ninf_lanes);
  cudaDeviceSynchronize();
  kn<NOISE><<<1, 12 * 32>>>(out, cyc, 0.5f, flag, ninf_lanes);
  cudaDeviceSynchronize();
  long long h[4];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("frame warp next to two helper warps per scheduler, helpers: %-34s %6.1f %6.1f %6.1f %6.1f cycles / frame  [%s]\n", name,
         (double)h[0] / N, (double)h[1] / N, (double)h[2] / N, (double)h[3] / N, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  runn<0>("exit at once");
  runn<0>("exit at once, 1 lane at -inf", 1);
  runn<0>("exit at once, 12 lanes at -inf", 12);
  runn<0>("exit at once, all lanes at -inf", 32);
  runn<4>("build tiles, 12 lanes at -inf", 12);
  runn<1>("tight shared-memory poll");
  runn<2>("poll + __nanosleep(100)");
  runn<3>("poll + __nanosleep(1000)");
  runn<4>("build tiles (16 LDS + 8 STS.64)");
  for (int w : {1, 4, 8, 16}) {
    printf("---- %d warp(s) in one CTA (%d per scheduler)\n", w, (w + 3) / 4);
    run<0>("SHFL.UP -> SHFL.UP", 1, w, w);
    run<1>("FADD -> FADD", 1, w, w);
    run<2>("FADD2 -> FADD2", 1, w, w);
    run<3>("FMNMX -> FMNMX (x2)", 2, w, w);
    run<4>("FMNMX3 -> FMNMX3 (x2)", 2, w, w);
    run<5>("SHFL.UP -> FADD2 -> FMNMX3", 3, w, w);
    run<6>("full 2-state frame (values only)", 3, w, w);
    run<7>("4 independent SHFL.UP", 0, w, w);
    run<8>("STS -> LDS -> FADD2 -> FMNMX3", 4, w, w);
    run<9>("full 2-state frame with backpointer bits", 3, w, w);
  }
  return 0;
}

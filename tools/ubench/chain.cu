// Latency / throughput of the instructions on the band kernels' recurrence chain (sm_100a):
// SHFL.UP, FADD, FADD2 (add.rn.f32x2), FMNMX, FMNMX3, and the chain SHFL -> FADD2 -> FMNMX3 itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/chain.bin tools/ubench/chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

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
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + t0 + t1;
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

int main() {
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
  }
  return 0;
}

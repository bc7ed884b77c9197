// FFMA vs FFMA2 issue / pipe throughput on sm_100a, and shared-memory float atomics vs plain RMW vs integer atomics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_fma(float *out, float a0, float b0, int iters) {
  float a = a0, b = b0;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
  } else {
    float2 *p = reinterpret_cast<float2 *>(x);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], a2, b2);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: atomicAdd(float) on shared; 1: plain LDS/FADD/STS; 2: atomicAdd(int) on shared.  30 lanes active, addresses bank-distinct.
template <int MODE>
__global__ void __launch_bounds__(256) k_atom(float *out, int iters, int stride) {
  __shared__ float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  float v = l * 0.5f + 1.f;
  if (l < 30) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float *p = sm + ((w * 1024 + l * stride + k * 32) & 8191);   // each warp its own region: no inter-warp contention
        if (MODE == 0) atomicAdd(p, v);
        else if (MODE == 1) *p += v;
        else atomicAdd(reinterpret_cast<int *>(p), (int)v);
      }
    }
  }
  __syncthreads();
  out[blockIdx.x * 256 + threadIdx.x] = sm[threadIdx.x];
}

int main() {
  float *out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_fma<0><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); else k_fma<1><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = 148.0 * 8 * 256 * (double)iters * 128;   // lane-FMAs
      printf("%s: %.3f ms  %.1f TFLOP/s  (%.1f lane-FMA/clk/SM at %.0f MHz max)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / (dev_clk * 1e3), dev_clk * 1e-3);
    }
  }
  for (int stride = 1; stride <= 33; stride += 32)
    for (int mode = 0; mode < 3; ++mode) {
      cudaEventRecord(e0);
      if (mode == 0) k_atom<0><<<148 * 2, 256>>>(out, 2000, stride);
      else if (mode == 1) k_atom<1><<<148 * 2, 256>>>(out, 2000, stride);
      else k_atom<2><<<148 * 2, 256>>>(out, 2000, stride);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops = 2.0 * 8 * 2000 * 16;   // warp-level ops per SM
      printf("smem %s stride %d: %.3f ms  %.2f clk per warp-op per SM\n", mode == 0 ? "atomicAdd(float)" : mode == 1 ? "plain RMW       " : "atomicAdd(int)  ", stride, ms, ms * 1e-3 * dev_clk * 1e3 / ops);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

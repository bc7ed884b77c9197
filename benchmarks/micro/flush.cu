// Cost of the P2G flush (21 sums per lane, lanes = 3 cells x 9 node columns, arena = eight [7][64] tiles) on sm_100a, by accumulation scheme.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o flush flush.cu && ./flush
// 8 warps per CTA flush into ONE shared arena (as in p2g_binned_kernel), `ctas` CTAs per SM; cells move every iteration.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0 atomicAdd(float), tile layout          1 plain RMW, tile layout (racy: cost reference only)
//      2 atomicAdd(int) of lrintf(v * scale)     3 atomicAdd(int), no conversion
//      4 atomicAdd(float), conflict-free layout  5 atomicAdd(int) no conversion, conflict-free layout
//      6 atomicAdd(unsigned long long): two channels per atomic, conflict-free layout (channel pairs adjacent)
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_flush(float *out, int iters) {
  __shared__ __align__(16) float sm[8 * 99 * 8];   // 6336 floats: tile layout needs 3584, padded layout 7 x 792 is emulated with ch stride 792 -> use ch < 7 only
  extern __shared__ float dyn[];
  float *arena = dyn;
  for (int i = threadIdx.x; i < 7 * 800; i += 256) arena[i] = 0.f;
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gi = l / 9, j = l % 9, ox = j / 3, oy = j % 3;
  const bool have = l < 27;
  float acc[7][3];
#pragma unroll
  for (int ch = 0; ch < 7; ++ch)
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[ch][k] = 0.001f * (l + 1) + ch + 0.1f * k;
  unsigned rng = w * 7919u + blockIdx.x * 104729u + 1u;
  for (int it = 0; it < iters; ++it) {
    rng = rng * 1664525u + 1013904223u;
    // a unit: three z-consecutive cells of one column; column (cx, cy) in [0,5]^2, z0 in [0,3]
    const int cx = (rng >> 8) % 6, cy = (rng >> 16) % 6, z0 = (rng >> 24) % 4;
    const int axn = cx + ox, ayn = cy + oy;
    if (have) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int azn = z0 + gi + k;
        if (MODE <= 3) {
          float *p = arena + (((axn >> 2) << 2) | ((ayn >> 2) << 1) | (azn >> 2)) * 448 + (((axn & 3) << 4) | ((ayn & 3) << 2) | (azn & 3));
#pragma unroll
          for (int ch = 0; ch < 7; ++ch) {
            if (MODE == 0) atomicAdd(p + ch * 64, acc[ch][k]);
            else if (MODE == 1) p[ch * 64] += acc[ch][k];
            else if (MODE == 2) atomicAdd(reinterpret_cast<int *>(p + ch * 64), __float2int_rn(acc[ch][k] * 1024.f));
            else atomicAdd(reinterpret_cast<int *>(p + ch * 64), __float_as_int(acc[ch][k]) & 0xff);
          }
        } else if (MODE <= 5) {
          float *p = arena + axn * 99 + ayn * 9 + azn;   // bank = 3 x + 9 y + z: the 27 lanes of one step hit 27 banks
#pragma unroll
          for (int ch = 0; ch < 7; ++ch) {
            if (MODE == 4) atomicAdd(p + ch * 792, acc[ch][k]);
            else atomicAdd(reinterpret_cast<int *>(p + ch * 792), __float_as_int(acc[ch][k]) & 0xff);
          }
        } else {
          unsigned long long *p = reinterpret_cast<unsigned long long *>(arena) + axn * 99 + ayn * 9 + azn;  // 8-byte slots: banks 2 (3x + 9y + z)
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) {
            const unsigned long long v = ((unsigned long long)(__float_as_int(acc[2 * c2][k]) & 0xff) << 32) | (unsigned)(__float_as_int(acc[(2 * c2 + 1) % 7][k]) & 0xff);
            atomicAdd(p + c2 * 792, v);
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 7; ++ch) acc[ch][it & 1] += 1.0f;   // keeps the values live / varying
  }
  __syncthreads();
  out[blockIdx.x * 256 + threadIdx.x] = arena[threadIdx.x] + sm[0];
}

template <int MODE> float run(float *out, int iters, int smem) {
  cudaFuncSetAttribute(k_flush<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_flush<MODE><<<148 * 4, 256, smem>>>(out, 10);
  cudaEventRecord(e0);
  k_flush<MODE><<<148 * 4, 256, smem>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float *out;
  cudaMalloc(&out, 148 * 4 * 256 * 4);
  int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
  const int iters = 4000;
  const char *names[] = {"atomicAdd(float) tile layout ", "plain RMW        tile layout ", "atomicAdd(int)+cvt tile layout", "atomicAdd(int)   tile layout ",
                         "atomicAdd(float) padded layout", "atomicAdd(int)   padded layout", "atomicAdd(u64) 2ch padded     "};
  for (int mode = 0; mode < 7; ++mode) {
    float ms = 0;
    const int smem = mode == 6 ? 4 * 800 * 8 : 7 * 800 * 4;
    switch (mode) {
      case 0: ms = run<0>(out, iters, smem); break; case 1: ms = run<1>(out, iters, smem); break; case 2: ms = run<2>(out, iters, smem); break;
      case 3: ms = run<3>(out, iters, smem); break; case 4: ms = run<4>(out, iters, smem); break; case 5: ms = run<5>(out, iters, smem); break;
      default: ms = run<6>(out, iters, smem); break;
    }
    // per SM: 4 CTAs x 8 warps x iters flushes
    const double flushes = 4.0 * 8 * iters;
    printf("%s: %.3f ms  %.1f clk per warp-flush per SM (21 values x 27 lanes)\n", names[mode], ms, ms * 1e-3 * dev_clk * 1e3 / flushes);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

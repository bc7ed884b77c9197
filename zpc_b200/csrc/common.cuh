// Shared helpers for the sm_100a kernels behind include/zpcb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/zpcb200.h"

#define ZPC_SM_COUNT 148  // B200: 2 dies x 74 SMs

extern std::atomic<int> g_zpc_launches;

#define ZPC_LAUNCHED() (g_zpc_launches.fetch_add(1, std::memory_order_relaxed))
#define ZPC_CHECK_LAUNCH()                       \
  do {                                           \
    ZPC_LAUNCHED();                              \
    cudaError_t e__ = cudaPeekAtLastError();     \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)
#define ZPC_CUDA(expr)                           \
  do {                                           \
    cudaError_t e__ = (expr);                    \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

static inline size_t zpc_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Element address of an iterator port (py_interop/GenericIterator.hpp:84-98).
template <typename T> struct PortAcc {
  T *base;
  uint32_t idx, bits, mask, chns;
  __host__ __device__ PortAcc() {}
  __host__ __device__ explicit PortAcc(const zpc_port &p)
      : base((T *)p.base), idx(p.idx), bits(p.numTileBits), mask(p.tileMask), chns(p.numChns) {}
  __host__ __device__ bool contiguous() const { return bits == 0 && chns == 1; }
  __device__ __forceinline__ T &operator[](size_t k) const {
    size_t i = (size_t)idx + k;
    if (chns == 1) return base[i];  // one channel: the tile formula is the identity (warp-uniform branch)
    return base[(((i >> bits) * chns) << bits) | (i & mask)];
  }
  // component d of a vector element (aosoa_iterator<T, N>: stride tileMask+1 between components)
  __device__ __forceinline__ T &at(size_t k, int d) const {
    size_t i = (size_t)idx + k;
    return base[((((i >> bits) * chns) << bits) | (i & mask)) + (size_t)d * (mask + 1)];
  }
};
template <typename T> struct PtrAcc {  // contiguous fast path
  T *base;
  __device__ __forceinline__ T &operator[](size_t k) const { return base[k]; }
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

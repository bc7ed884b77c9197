// reduce / inclusive_scan / exclusive_scan / radix_sort(_pair) for sm_100a.
//
// Replaces the CUB calls under zs::CudaExecutionPolicy (reference:
// include/zensim/cuda/execution/ExecutionPolicy.cuh:552-866).  All three are HBM-bound streaming
// kernels: reduce moves N*sizeof(T), scan 2*N*sizeof(T), a P-pass radix sort
// N*K + P*2*N*(K+V).
//   reduce : grid of k*148 CTAs, 128-bit loads, one partial per CTA, last CTA folds partials in a
//            fixed order (deterministic for floats given n).
//   scan   : single pass, decoupled look-back (tile descriptors = {flag,value} in one 64/128-bit word).
//   sort   : "onesweep" LSD radix sort, 8-bit digits: one histogram pass over the keys for all digit
//            positions, then one read+write pass per digit with decoupled look-back on per-tile digit
//            counts; stable (warp-synchronous match ranking), keys_out doubles as a ping-pong buffer
//            so scratch is ONE extra key/value buffer (the reference wraps CUB with 4 temp vectors and
//            4 copy kernels, ExecutionPolicy.cuh:794-820).
#include <cfloat>
#include <climits>

#include "common.cuh"

std::atomic<int> g_zpc_launches{0};

namespace {

// ------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------
struct OpSum { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpProd { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return a * b; } };
struct OpMin { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return b < a ? b : a; } };
struct OpMax { template <typename T> __device__ __forceinline__ T operator()(T a, T b) const { return b > a ? b : a; } };

template <typename T> __device__ __forceinline__ T shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <typename T> __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <typename T> __device__ __forceinline__ T shfl_idx(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

template <typename T, typename Op> __device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = op(v, shfl_down(v, d));
  return v;  // valid in lane 0
}

// block-wide reduce; result valid in thread 0.  NT threads, NT/32 <= 32.
template <int NT, typename T, typename Op> __device__ __forceinline__ T block_reduce(T v, Op op, T ident, T *smem) {
  v = warp_reduce(v, op);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = l < NT / 32 ? smem[l] : ident;
    v = warp_reduce(v, op);
  }
  return v;
}

// ------------------------------------------------------------------------------------------------
// reduce
// ------------------------------------------------------------------------------------------------
constexpr int RED_NT = 512;
constexpr int RED_CTAS = ZPC_SM_COUNT * 4;  // 4 resident CTAs of 512 threads per SM

template <typename T> struct VecOf;  // 16-byte vector of T
template <> struct VecOf<int32_t> { using type = int4; static constexpr int N = 4; };
template <> struct VecOf<uint32_t> { using type = uint4; static constexpr int N = 4; };
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecOf<int64_t> { using type = longlong2; static constexpr int N = 2; };
template <> struct VecOf<uint64_t> { using type = ulonglong2; static constexpr int N = 2; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

template <typename T, typename Op, bool CONTIG>
__global__ void __launch_bounds__(RED_NT) reduce_kernel(PortAcc<T> in, size_t n, T ident, T *partials,
                                                        unsigned *counter, PortAcc<T> out) {
  __shared__ T smem[32];
  __shared__ bool is_last;
  Op op;
  T acc = ident;
  const size_t tid = (size_t)blockIdx.x * RED_NT + threadIdx.x, nthreads = (size_t)gridDim.x * RED_NT;
  if constexpr (CONTIG) {
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::N;
    const T *p = in.base + in.idx;
    // peel to 16-byte alignment
    size_t head = ((16 - ((uintptr_t)p & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    if (tid < head) acc = op(acc, p[tid]);
    const V *pv = reinterpret_cast<const V *>(p + head);
    const size_t nv = (n - head) / VN;
    // 4 independent 128-bit loads in flight per thread
    size_t i = tid;
    for (; i + 3 * nthreads < nv; i += 4 * nthreads) {
      V a = pv[i], b = pv[i + nthreads], c = pv[i + 2 * nthreads], d = pv[i + 3 * nthreads];
      const T *e;
      e = reinterpret_cast<const T *>(&a);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc = op(acc, e[k]);
      e = reinterpret_cast<const T *>(&b);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc = op(acc, e[k]);
      e = reinterpret_cast<const T *>(&c);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc = op(acc, e[k]);
      e = reinterpret_cast<const T *>(&d);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc = op(acc, e[k]);
    }
    for (; i < nv; i += nthreads) {
      V a = pv[i];
      const T *e = reinterpret_cast<const T *>(&a);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc = op(acc, e[k]);
    }
    const size_t tail0 = head + nv * VN;
    if (tail0 + tid < n) acc = op(acc, p[tail0 + tid]);
  } else {
    for (size_t i = tid; i < n; i += nthreads) acc = op(acc, in[i]);
  }
  acc = block_reduce<RED_NT>(acc, op, ident, smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    unsigned done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    T v = ident;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += RED_NT) v = op(v, ((volatile T *)partials)[i]);
    __syncthreads();
    v = block_reduce<RED_NT>(v, op, ident, smem);
    if (threadIdx.x == 0) {
      out[0] = v;
      *counter = 0;  // leave scratch reusable
    }
  }
}

template <typename T, typename Op>
int reduce_impl(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, T ident, cudaStream_t s) {
  const size_t need = 256 + sizeof(T) * RED_CTAS;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  unsigned *counter = (unsigned *)temp;
  T *partials = (T *)((char *)temp + 256);
  PortAcc<T> pin(in), pout(out);
  ZPC_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), s));
  size_t per_cta = (size_t)RED_NT * 16;
  int grid = (int)((n + per_cta - 1) / per_cta);
  if (grid < 1) grid = 1;
  if (grid > RED_CTAS) grid = RED_CTAS;
  if (pin.contiguous())
    reduce_kernel<T, Op, true><<<grid, RED_NT, 0, s>>>(pin, n, ident, partials, counter, pout);
  else
    reduce_kernel<T, Op, false><<<grid, RED_NT, 0, s>>>(pin, n, ident, partials, counter, pout);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

// ------------------------------------------------------------------------------------------------
// scan: single-pass decoupled look-back
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_NT = 256;
constexpr int SCAN_VPT = 16;                           // 16-byte vectors per thread (4-byte T: 16 items)
enum : uint32_t { FLAG_EMPTY = 0, FLAG_AGG = 1, FLAG_INCL = 2 };

// tile descriptor: one word holding {flag, value}
template <typename T, int SZ = sizeof(T)> struct TileDesc;
template <typename T> struct TileDesc<T, 4> {
  using word = unsigned long long;
  static __device__ __forceinline__ void store(word *p, uint32_t flag, T v) {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    word w = ((word)flag << 32) | bits;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  }
  static __device__ __forceinline__ void load(const word *p, uint32_t &flag, T &v) {
    word w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    flag = (uint32_t)(w >> 32);
    uint32_t bits = (uint32_t)w;
    memcpy(&v, &bits, 4);
  }
};
template <typename T> struct TileDesc<T, 8> {
  struct alignas(16) word { unsigned long long flag, val; };
  static __device__ __forceinline__ void store(word *p, uint32_t flag, T v) {
    unsigned long long bits;
    memcpy(&bits, &v, 8);
    asm volatile("{ .reg .b128 t; mov.b128 t, {%1, %2}; st.relaxed.gpu.global.b128 [%0], t; }" ::"l"(p),
                 "l"((unsigned long long)flag), "l"(bits)
                 : "memory");
  }
  static __device__ __forceinline__ void load(const word *p, uint32_t &flag, T &v) {
    unsigned long long f, bits;
    asm volatile("{ .reg .b128 t; ld.relaxed.gpu.global.b128 t, [%2]; mov.b128 {%0, %1}, t; }"
                 : "=l"(f), "=l"(bits)
                 : "l"(p)
                 : "memory");
    flag = (uint32_t)f;
    memcpy(&v, &bits, 8);
  }
};

template <typename T, bool INCLUSIVE, bool CONTIG>
__global__ void __launch_bounds__(SCAN_NT) scan_kernel(PortAcc<T> in, PortAcc<T> out, size_t n,
                                                       typename TileDesc<T>::word *desc, unsigned *ticket) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::N;
  constexpr int TILE = SCAN_NT * SCAN_VPT * VN;
  constexpr int NW = SCAN_NT / 32;
  __shared__ T warp_tot[NW];
  __shared__ T tile_excl;
  __shared__ unsigned s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const size_t base = (size_t)tile * TILE;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  // element (j, lane, c) of warp w sits at  base + w*(32*VPT*VN) + j*(32*VN) + lane*VN + c
  const size_t wbase = base + (size_t)w * (32 * SCAN_VPT * VN);
  T x[SCAN_VPT][VN];
  const bool full = base + TILE <= n;
  bool vec_ok = false;
  if constexpr (CONTIG) vec_ok = full && ((((uintptr_t)(in.base + in.idx)) & 15) == 0);
  if (vec_ok) {
    const V *pv = reinterpret_cast<const V *>(in.base + in.idx + wbase);
#pragma unroll
    for (int j = 0; j < SCAN_VPT; ++j) {
      V v = pv[j * 32 + l];
      memcpy(x[j], &v, sizeof(V));
    }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_VPT; ++j)
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        size_t e = wbase + (size_t)j * (32 * VN) + l * VN + c;
        x[j][c] = e < n ? in[e] : (T)0;
      }
  }
  // thread-local inclusive prefix inside each vector, vector totals
  T vt[SCAN_VPT];
#pragma unroll
  for (int j = 0; j < SCAN_VPT; ++j) {
#pragma unroll
    for (int c = 1; c < VN; ++c) x[j][c] = x[j][c - 1] + x[j][c];
    vt[j] = x[j][VN - 1];
  }
  // warp scan of the vector totals per row j, then carry rows
  T row_carry = (T)0;
  T pre[SCAN_VPT];  // exclusive prefix of this thread's vector j inside the warp chunk
#pragma unroll
  for (int j = 0; j < SCAN_VPT; ++j) {
    T inc = vt[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T t = shfl_up(inc, d);
      if (l >= d) inc = t + inc;
    }
    T ex = shfl_up(inc, 1);
    if (l == 0) ex = (T)0;
    pre[j] = row_carry + ex;
    row_carry = row_carry + shfl_idx(inc, 31);
  }
  if (l == 0) warp_tot[w] = row_carry;
  __syncthreads();
  // warp 0: block aggregate + look-back
  if (w == 0) {
    T wt = l < NW ? warp_tot[l] : (T)0;
    T inc = wt;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      T t = shfl_up(inc, d);
      if (l >= d) inc = t + inc;
    }
    const T tile_total = shfl_idx(inc, NW - 1);
    T wex = shfl_up(inc, 1);
    if (l == 0) wex = (T)0;
    if (l < NW) warp_tot[l] = wex;  // exclusive warp offsets
    T excl = (T)0;
    if (tile == 0) {
      if (l == 0) TileDesc<T>::store(desc, FLAG_INCL, tile_total);
    } else {
      if (l == 0) TileDesc<T>::store(desc + tile, FLAG_AGG, tile_total);
      long long pred = (long long)tile - 1 - l;
      while (true) {
        uint32_t flag = FLAG_INCL;
        T val = (T)0;
        if (pred >= 0) {
          do { TileDesc<T>::load(desc + pred, flag, val); } while (flag == FLAG_EMPTY);
        }
        const unsigned incl = __ballot_sync(0xffffffffu, flag == FLAG_INCL);
        const int first = __ffs(incl) - 1;  // nearest predecessor holding an inclusive prefix
        T contrib = (incl == 0 || l <= first) ? val : (T)0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contrib = contrib + shfl_down(contrib, d);
        excl = excl + shfl_idx(contrib, 0);
        if (incl) break;
        pred -= 32;
      }
      if (l == 0) TileDesc<T>::store(desc + tile, FLAG_INCL, excl + tile_total);
    }
    if (l == 0) tile_excl = excl;
  }
  __syncthreads();
  const T off = tile_excl + warp_tot[w];
#pragma unroll
  for (int j = 0; j < SCAN_VPT; ++j) {
    const T p = off + pre[j];
    T y[VN];
    if constexpr (INCLUSIVE) {
#pragma unroll
      for (int c = 0; c < VN; ++c) y[c] = p + x[j][c];
    } else {
      y[0] = p;
#pragma unroll
      for (int c = 1; c < VN; ++c) y[c] = p + x[j][c - 1];
    }
    bool st_vec = false;
    if constexpr (CONTIG) st_vec = full && ((((uintptr_t)(out.base + out.idx)) & 15) == 0);
    if (st_vec) {
      V v;
      memcpy(&v, y, sizeof(V));
      reinterpret_cast<V *>(out.base + out.idx + wbase)[j * 32 + l] = v;
    } else {
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        size_t e = wbase + (size_t)j * (32 * VN) + l * VN + c;
        if (e < n) out[e] = y[c];
      }
    }
  }
}

template <typename T, bool INCLUSIVE>
int scan_impl(void *temp, size_t *temp_bytes, zpc_port in, zpc_port out, size_t n, cudaStream_t s) {
  constexpr int TILE = SCAN_NT * SCAN_VPT * VecOf<T>::N;
  using word = typename TileDesc<T>::word;
  const size_t tiles = (n + TILE - 1) / TILE;
  const size_t need = 256 + sizeof(word) * (tiles ? tiles : 1);
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (n == 0) return ZPCB200_OK;
  ZPC_CUDA(cudaMemsetAsync(temp, 0, need, s));
  unsigned *ticket = (unsigned *)temp;
  word *desc = (word *)((char *)temp + 256);
  PortAcc<T> pin(in), pout(out);
  if (pin.contiguous() && pout.contiguous())
    scan_kernel<T, INCLUSIVE, true><<<(unsigned)tiles, SCAN_NT, 0, s>>>(pin, pout, n, desc, ticket);
  else
    scan_kernel<T, INCLUSIVE, false><<<(unsigned)tiles, SCAN_NT, 0, s>>>(pin, pout, n, desc, ticket);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

// ------------------------------------------------------------------------------------------------
// radix sort: onesweep
// ------------------------------------------------------------------------------------------------
constexpr int RS_NT = 256;
constexpr int RS_NW = RS_NT / 32;
constexpr int RS_BINS = 256;
// look-back descriptor = 2 flag bits + a 30-bit count.  n <= 2^30 is safe: an aggregate is at most one tile, an inclusive prefix of
// tile t is at most n minus the keys of the later tiles, i.e. < 2^30 for every tile that HAS a successor; the one value that can
// reach 2^30 exactly (last tile, every key in one digit) wraps to 0 and is never read — nobody looks back at the last tile.
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_VAL_MASK = (1u << 30) - 1;

template <typename K> struct KeyBits;
template <> struct KeyBits<uint32_t> { using U = uint32_t; static constexpr U flip = 0; };
template <> struct KeyBits<int32_t> { using U = uint32_t; static constexpr U flip = 0x80000000u; };
template <> struct KeyBits<uint64_t> { using U = uint64_t; static constexpr U flip = 0; };

template <typename K> __device__ __forceinline__ unsigned digit_of(K k, int shift, unsigned mask) {
  using U = typename KeyBits<K>::U;
  return (unsigned)((((U)k) ^ KeyBits<K>::flip) >> shift) & mask;
}

// histogram of every digit position in one pass over the keys.  Shared-memory integer atomics (native ATOMS.ADD), one
// per key and digit position; a warp whose 32 keys share the digit (constant high bytes, already sorted ranges) adds
// once instead of serialising 32 same-address atomics.  Contiguous ranges are read with 128-bit loads.
template <typename K, int MAXP>
__device__ __forceinline__ void rs_hist_add(unsigned (*sh)[RS_BINS], K k, bool valid, int sbit, int ebit, int npass) {
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
    if (p < npass) {
      const int shift = sbit + 8 * p;
      const int bits = min(8, ebit - shift);
      const unsigned d = valid ? digit_of(k, shift, (1u << bits) - 1) : RS_BINS;
      const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
      if (__all_sync(0xffffffffu, d == d0)) {
        if ((threadIdx.x & 31) == 0 && d0 < RS_BINS) atomicAdd(&sh[p][d0], 32u);
      } else if (valid) {
        atomicAdd(&sh[p][d], 1u);
      }
    }
}
template <typename K, int MAXP>
__global__ void __launch_bounds__(512) rs_hist_kernel(PortAcc<K> keys, size_t n, int sbit, int ebit, int npass,
                                                      unsigned *ghist /*[npass][256]*/) {
  __shared__ unsigned sh[MAXP][RS_BINS];
  for (int i = threadIdx.x; i < MAXP * RS_BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);  // warp-uniform trip counts (votes inside)
  constexpr int V = 16 / (int)sizeof(K);
  size_t done = 0;  // keys [0, done) are covered by the vector loop
  if (keys.contiguous() && ((uintptr_t)(keys.base + keys.idx) & 15) == 0) {
    using VT = typename VecOf<K>::type;
    const VT *v = reinterpret_cast<const VT *>(keys.base + keys.idx);
    const size_t nv = n / V;
    for (size_t i0 = first; i0 < nv; i0 += stride) {
      const size_t i = i0 + (threadIdx.x & 31);
      const bool valid = i < nv;
      VT q = valid ? v[i] : VT();
      const K *kk = reinterpret_cast<const K *>(&q);
#pragma unroll
      for (int j = 0; j < V; ++j) rs_hist_add<K, MAXP>(sh, kk[j], valid, sbit, ebit, npass);
    }
    done = nv * V;
  }
  for (size_t i0 = done + first; i0 < n; i0 += stride) {
    const size_t i = i0 + (threadIdx.x & 31);
    const bool valid = i < n;
    const K k = valid ? keys[i] : (K)0;
    rs_hist_add<K, MAXP>(sh, k, valid, sbit, ebit, npass);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * RS_BINS; i += blockDim.x) {
    unsigned c = (&sh[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}
// exclusive scan of each 256-bin histogram (one CTA of 256 threads per pass)
__global__ void rs_scan_hist_kernel(unsigned *ghist) {
  __shared__ unsigned wsum[8];
  unsigned *h = ghist + blockIdx.x * RS_BINS;
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned v = h[threadIdx.x], inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
    if (l >= d) inc += t;
  }
  if (l == 31) wsum[w] = inc;
  __syncthreads();
  unsigned off = 0;
  for (int i = 0; i < w; ++i) off += wsum[i];
  h[threadIdx.x] = off + inc - v;
}

// element access of one tile: raw pointers when every range is contiguous (FAST), iterator ports otherwise
template <typename T, bool FAST> struct TileAcc;
template <typename T> struct TileAcc<T, true> {
  T *p;
  __device__ __forceinline__ TileAcc(const PortAcc<T> &a, size_t base) : p(a.base + a.idx + base) {}
  __device__ __forceinline__ T &operator[](long long i) const { return p[i]; }
};
template <typename T> struct TileAcc<T, false> {
  PortAcc<T> a;
  size_t base;
  __device__ __forceinline__ TileAcc(const PortAcc<T> &a_, size_t base_) : a(a_), base(base_) {}
  __device__ __forceinline__ T &operator[](long long i) const { return a[(size_t)((long long)base + i)]; }
};

template <typename K, bool PAIRS, int ITEMS, bool FAST>
#ifndef ZPC_RS_ITEMS4
#define ZPC_RS_ITEMS4 16
#endif
#ifndef ZPC_RS_MINB4   // CTAs per SM for 4-byte keys: 4 (64 registers) since the lane-mask ranking; 3 (80 registers) was right for the ballots
#define ZPC_RS_MINB4 4
#endif
#ifndef ZPC_RS_MATCH_OR   // 1: ranking through shared-memory lane masks (atomicOr); 0: one ballot per digit bit
#define ZPC_RS_MATCH_OR 1
#endif
__global__ void __launch_bounds__(RS_NT, sizeof(K) == 4 ? ZPC_RS_MINB4 : 2) rs_onesweep_kernel(PortAcc<K> kin, PortAcc<int> vin, PortAcc<K> kout,
                                                            PortAcc<int> vout, size_t n, int shift, unsigned mask, int nbits,
                                                            const unsigned *gbase /*[256] exclusive*/,
                                                            unsigned *lookback /*[tiles][256]*/, unsigned *ticket) {
  constexpr int TILE = RS_NT * ITEMS;
  __shared__ unsigned warp_hist[RS_NW][RS_BINS];  // per-warp digit counters -> exclusive warp offsets
  __shared__ unsigned tile_base[RS_BINS];         // exclusive scan of tile digit counts (position in tile)
  __shared__ int gofs[RS_BINS];                   // global position of tile-sorted slot j: j + gofs[digit]
  __shared__ unsigned wsum[RS_NW];
  __shared__ unsigned s_tile;
  __shared__ K skeys[TILE];
#if ZPC_RS_MATCH_OR
  // per-warp lane masks by digit, two items in flight; all zero between items.  They live in the staging buffer of the scatter phase,
  // which starts after the ranking (two barriers later)
  static_assert(sizeof(K) * TILE >= 2 * RS_NW * RS_BINS * sizeof(unsigned), "the match masks are overlaid on skeys");
  unsigned (*warp_match)[RS_NW][RS_BINS] = reinterpret_cast<unsigned (*)[RS_NW][RS_BINS]>(skeys);
#endif
  int *svals = reinterpret_cast<int *>(skeys);  // values staged after keys are written out (PAIRS, sizeof(K)>=4)

  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < RS_NW * RS_BINS; i += RS_NT) (&warp_hist[0][0])[i] = 0;
#if ZPC_RS_MATCH_OR
  for (int i = threadIdx.x; i < 2 * RS_NW * RS_BINS; i += RS_NT) (&warp_match[0][0][0])[i] = 0;
#endif
  __syncthreads();
  const unsigned tile = s_tile;
  const size_t base = (size_t)tile * TILE;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int wofs = w * (32 * ITEMS) + l;  // tile-local index of this thread's item i: wofs + 32 i
  const int cnt_tile = (int)min((size_t)TILE, n - base);
  const TileAcc<K, FAST> tk(kin, base);

  K key[ITEMS];
  unsigned rank[ITEMS];  // rank among same-digit keys of this warp (then of this tile)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) key[i] = (wofs + 32 * i) < cnt_tile ? tk[wofs + 32 * i] : (K)0;
  const bool full_tile = cnt_tile == TILE;  // CTA-uniform: every lane holds a key, no validity votes
#if ZPC_RS_MATCH_OR
  // Warp-synchronous stable ranking through shared-memory masks: every lane ORs its lane bit into the word of its digit (a native
  // integer atomic, one pass of the data pipe), reads the word back = the set of lanes holding that digit, and the lowest of them
  // clears the word and adds the group's size to the warp's digit counter.  About 20 instructions per key where one ballot per digit
  // bit took 85 (ncu: 51 % of the pass's instructions, the pass is issue-bound); two items are in flight per step (two mask arrays)
  // so that the shared-memory round trips of one overlap the other's.  MATCH.ANY would be one instruction, but its latency is long
  // and poorly pipelined on sm_100 (the histogram pass spent 7.5 ms in it at 2^28).
  {
    unsigned *wm0 = &warp_match[0][w][0], *wm1 = &warp_match[1][w][0];
    const unsigned lbit = 1u << l, lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < ITEMS; i += 2) {
      const bool va = full_tile || (wofs + 32 * i) < cnt_tile, vb = full_tile || (wofs + 32 * (i + 1)) < cnt_tile;
      const unsigned da = digit_of(key[i], shift, mask), db = digit_of(key[i + 1], shift, mask);
      if (va) atomicOr(&wm0[da], lbit);
      if (vb) atomicOr(&wm1[db], lbit);
      __syncwarp();
      const unsigned pa = va ? wm0[da] : 0u, pb = vb ? wm1[db] : 0u;
      __syncwarp();   // every lane has read the masks before a leader clears them
      const int la = __ffs(pa) - 1, lb = __ffs(pb) - 1;
      unsigned ca = 0, cb = 0;
      if (va && l == la) { wm0[da] = 0u; ca = atomicAdd(&warp_hist[w][da], (unsigned)__popc(pa)); }
      __syncwarp();   // item i's counter updates are ordered before item i + 1's (stability)
      if (vb && l == lb) { wm1[db] = 0u; cb = atomicAdd(&warp_hist[w][db], (unsigned)__popc(pb)); }
      ca = __shfl_sync(0xffffffffu, ca, la < 0 ? 0 : la);
      cb = __shfl_sync(0xffffffffu, cb, lb < 0 ? 0 : lb);
      rank[i] = ca + (unsigned)__popc(pa & lt);
      rank[i + 1] = cb + (unsigned)__popc(pb & lt);
      __syncwarp();
    }
  }
#else
  // warp-synchronous stable ranking: lanes holding the same digit are found with one ballot per digit bit
  // (MATCH.ANY has a long, poorly pipelined latency on sm_100), then one shared-memory counter per (warp, digit)
  // (1) peer masks of all items first: 16 independent vote chains the scheduler can overlap.  Packed into rank[i]:
  //     bits 0-4 = same-digit lanes before this one, 5-9 = leader lane, 10-15 = group size.
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const bool valid = full_tile || (wofs + 32 * i) < cnt_tile;
    const unsigned d = digit_of(key[i], shift, mask);
    unsigned peers = 0xffffffffu;
    if (!full_tile) {
      peers = __ballot_sync(0xffffffffu, valid);
      if (!valid) peers = ~peers;
    }
    if (nbits == 8) {  // the common pass: fully unrolled, one vote + one LOP3 per digit bit
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const unsigned bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= vote ^ (bit - 1u);  // bit ? vote : ~vote
      }
    } else {
      for (int b = 0; b < nbits; ++b) {
        const unsigned bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= vote ^ (bit - 1u);
      }
    }
    rank[i] = (unsigned)__popc(peers & lanemask_lt()) | ((unsigned)(__ffs(peers) - 1) << 5) | ((unsigned)__popc(peers) << 10);
  }
  // (2) one shared atomic per (item, digit group), in item order (stability); the group's base goes to its lanes
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const bool valid = full_tile || (wofs + 32 * i) < cnt_tile;
    const int leader = (int)((rank[i] >> 5) & 31u);
    unsigned c = 0;
    if (valid && l == leader) c = atomicAdd(&warp_hist[w][digit_of(key[i], shift, mask)], rank[i] >> 10);
    c = __shfl_sync(0xffffffffu, c, leader);
    rank[i] = c + (rank[i] & 31u);
    __syncwarp();  // item i's counter updates are ordered before item i+1's
  }
#endif
  __syncthreads();
  // per digit (thread d): exclusive prefix over warps, tile count
  {
    const int d = threadIdx.x;  // RS_NT == RS_BINS
    unsigned run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_NW; ++ww) {
      unsigned c = warp_hist[ww][d];
      warp_hist[ww][d] = run;
      run += c;
    }
    const unsigned tcount = run;
    // publish aggregate, then look back for this digit
    unsigned *lb = lookback + (size_t)tile * RS_BINS + d;
    unsigned excl = 0;
    if (tile == 0) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lb), "r"(RS_FLAG_INCL | tcount) : "memory");
    } else {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lb), "r"(RS_FLAG_AGG | tcount) : "memory");
      // walk back over the predecessors' descriptors, LB independent loads in flight per round (the walk is pure
      // L2 latency: one load at a time would make every tile wait ~#tiles-in-flight round trips)
      constexpr int LB = 8;
      long long pred = (long long)tile - 1;
      bool done = false;
      while (!done) {
        unsigned v[LB];
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          const long long pk = pred - k;
          v[k] = RS_FLAG_INCL;  // before tile 0: inclusive 0 terminates the walk
          if (pk >= 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[k]) : "l"(lookback + (size_t)pk * RS_BINS + d) : "memory");
        }
#pragma unroll
        for (int k = 0; k < LB; ++k) {
          if (done) break;
          if ((v[k] >> 30) == 0) {  // not published yet: spin on this one
            const unsigned *pp = lookback + (size_t)(pred - k) * RS_BINS + d;
            do {
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[k]) : "l"(pp) : "memory");
            } while ((v[k] >> 30) == 0);
          }
          excl += v[k] & RS_VAL_MASK;
          if ((v[k] >> 30) == 2) done = true;
        }
        pred -= LB;
      }
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(lb), "r"(RS_FLAG_INCL | ((excl + tcount) & RS_VAL_MASK))
                   : "memory");
    }
    // exclusive scan of tile counts across digits -> tile_base
    unsigned inc = tcount;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, inc, dd);
      if (l >= dd) inc += t;
    }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    unsigned off = 0;
#pragma unroll
    for (int ww = 0; ww < RS_NW; ++ww) off += (ww < w) ? wsum[ww] : 0;
    const unsigned tb = off + inc - tcount;
    tile_base[d] = tb;
    gofs[d] = (int)(gbase[d] + excl) - (int)tb;
  }
  __syncthreads();
  // scatter into shared memory in tile-sorted order
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if ((wofs + 32 * i) < cnt_tile) {
      const unsigned d = digit_of(key[i], shift, mask);
      rank[i] = tile_base[d] + warp_hist[w][d] + rank[i];
      skeys[rank[i]] = key[i];
    }
  }
  __syncthreads();
  // coalesced write-out: slot j goes to j + gofs[digit(key_j)]
  const TileAcc<K, FAST> ok(kout, 0);
  unsigned dj[ITEMS];  // digit of the key in tile slot threadIdx.x + i*RS_NT (reused for the values)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int j = threadIdx.x + i * RS_NT;
    dj[i] = 0;
    if (j < cnt_tile) {
      const K k = skeys[j];
      dj[i] = digit_of(k, shift, mask);
      ok[(long long)j + gofs[dj[i]]] = k;
    }
  }
  if constexpr (PAIRS) {
    const TileAcc<int, FAST> tv(vin, base), ov(vout, 0);
    int val[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) val[i] = (wofs + 32 * i) < cnt_tile ? tv[wofs + 32 * i] : 0;
    __syncthreads();  // every thread is done reading skeys
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if ((wofs + 32 * i) < cnt_tile) svals[rank[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int j = threadIdx.x + i * RS_NT;
      if (j < cnt_tile) ov[(long long)j + gofs[dj[i]]] = svals[j];
    }
  }
}

template <typename K> struct RsCfg { static constexpr int ITEMS = sizeof(K) == 4 ? ZPC_RS_ITEMS4 : 12; };

template <typename K, bool PAIRS>
int radix_sort_impl(void *temp, size_t *temp_bytes, zpc_port keys_in, zpc_port vals_in, zpc_port keys_out,
                    zpc_port vals_out, size_t n, int sbit, int ebit, cudaStream_t s) {
  constexpr int ITEMS = RsCfg<K>::ITEMS;
  constexpr int TILE = RS_NT * ITEMS;
  constexpr int MAXP = sizeof(K);
  if (sbit < 0 || ebit > (int)sizeof(K) * 8 || ebit < sbit) return ZPCB200_E_BADARG;
  if (n > ((size_t)1 << 30)) return ZPCB200_E_UNSUPPORTED;
  const int npass = (ebit - sbit + 7) / 8;
  const size_t tiles = (n + TILE - 1) / TILE;
  // layout: [ticket[MAXP] | pad to 256][ghist MAXP*256 u32][lookback npass*tiles*256 u32][key buf][val buf]
  const size_t off_hist = 256;
  const size_t off_lb = off_hist + sizeof(unsigned) * MAXP * RS_BINS;
  const size_t lb_bytes = sizeof(unsigned) * (size_t)(npass > 0 ? npass : 1) * (tiles ? tiles : 1) * RS_BINS;
  const size_t off_kbuf = zpc_align_up(off_lb + lb_bytes, 256);
  const size_t off_vbuf = zpc_align_up(off_kbuf + sizeof(K) * n, 256);
  const size_t need = off_vbuf + (PAIRS ? sizeof(int) * n : 0) + 256;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (n == 0) return ZPCB200_OK;
  char *t = (char *)temp;
  unsigned *ticket = (unsigned *)t;
  unsigned *ghist = (unsigned *)(t + off_hist);
  unsigned *lookback = (unsigned *)(t + off_lb);
  zpc_port kbuf = {t + off_kbuf, 0, 0, 0, 1}, vbuf = {t + off_vbuf, 0, 0, 0, 1};
  PortAcc<K> pkin(keys_in);
  // npass == 0 (ebit == sbit): a stable sort on no bits is the identity; run one zero-width pass
  ZPC_CUDA(cudaMemsetAsync(t, 0, off_lb + lb_bytes, s));
  const int passes = npass > 0 ? npass : 1;
  {
    int grid = (int)((n + 512 * 16 - 1) / (512 * 16));
    if (grid > ZPC_SM_COUNT * 4) grid = ZPC_SM_COUNT * 4;
    if (npass > 0) {
      rs_hist_kernel<K, MAXP><<<grid, 512, 0, s>>>(pkin, n, sbit, ebit, npass, ghist);
      ZPC_CHECK_LAUNCH();
      rs_scan_hist_kernel<<<npass, RS_BINS, 0, s>>>(ghist);
      ZPC_CHECK_LAUNCH();
    }
  }
  // ping-pong so that the last pass lands in keys_out: odd #passes: in->out->tmp->out..., even: in->tmp->out...
  zpc_port src_k = keys_in, src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    zpc_port dst_k = to_out ? keys_out : kbuf, dst_v = to_out ? vals_out : vbuf;
    const int shift = sbit + 8 * p;
    const int bits = npass > 0 ? ((ebit - shift) < 8 ? (ebit - shift) : 8) : 0;
    const unsigned mask = (1u << bits) - 1;
    const bool fast = src_k.numChns == 1 && dst_k.numChns == 1 && (!PAIRS || (src_v.numChns == 1 && dst_v.numChns == 1));
    if (fast)
      rs_onesweep_kernel<K, PAIRS, ITEMS, true><<<(unsigned)tiles, RS_NT, 0, s>>>(
          PortAcc<K>(src_k), PortAcc<int>(src_v), PortAcc<K>(dst_k), PortAcc<int>(dst_v), n, shift, mask, bits,
          ghist + p * RS_BINS, lookback + (size_t)p * tiles * RS_BINS, ticket + p);
    else
      rs_onesweep_kernel<K, PAIRS, ITEMS, false><<<(unsigned)tiles, RS_NT, 0, s>>>(
          PortAcc<K>(src_k), PortAcc<int>(src_v), PortAcc<K>(dst_k), PortAcc<int>(dst_v), n, shift, mask, bits,
          ghist + p * RS_BINS, lookback + (size_t)p * tiles * RS_BINS, ticket + p);
    ZPC_CHECK_LAUNCH();
    src_k = dst_k;
    src_v = dst_v;
  }
  return ZPCB200_OK;
}

// ------------------------------------------------------------------------------------------------
// merge_sort / merge_sort_pair (execution/ExecutionPolicy.hpp:285-455, cuda/execution/ExecutionPolicy.cuh:686-760):
// stable ascending sort under operator<, IN PLACE.  A stable sort under a strict weak order has one result, so it is
// produced with the radix machinery: keys are mapped to an order-preserving unsigned image (-0.0 and +0.0 compare
// equal and therefore share an image), (image, index) pairs are radix-sorted, then keys and values are gathered
// through the sorted indices and written back.  NaN keys (unordered under <) sort by their bit pattern.
// ------------------------------------------------------------------------------------------------
template <typename T> struct OrdKey;
template <> struct OrdKey<int32_t> {
  using U = uint32_t;
  static __device__ __forceinline__ U make(int32_t k) { return (uint32_t)k ^ 0x80000000u; }
};
template <> struct OrdKey<uint32_t> {
  using U = uint32_t;
  static __device__ __forceinline__ U make(uint32_t k) { return k; }
};
template <> struct OrdKey<int64_t> {
  using U = uint64_t;
  static __device__ __forceinline__ U make(int64_t k) { return (uint64_t)k ^ 0x8000000000000000ull; }
};
template <> struct OrdKey<uint64_t> {
  using U = uint64_t;
  static __device__ __forceinline__ U make(uint64_t k) { return k; }
};
template <> struct OrdKey<float> {
  using U = uint32_t;
  static __device__ __forceinline__ U make(float x) {
    uint32_t b = __float_as_uint(x);
    if (x == 0.f) b = 0u;
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
  }
};
template <> struct OrdKey<double> {
  using U = uint64_t;
  static __device__ __forceinline__ U make(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    if (x == 0.0) b = 0ull;
    return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
  }
};
template <typename T>
__global__ void ms_prepare_kernel(PortAcc<T> keys, size_t n, typename OrdKey<T>::U *img, int *idx) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  img[i] = OrdKey<T>::make(keys[i]);
  idx[i] = (int)i;
}
template <typename T, bool PAIRS>
__global__ void ms_gather_kernel(PortAcc<T> keys, PortAcc<int> vals, const int *__restrict__ idx, size_t n, T *ktmp, int *vtmp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t j = (size_t)idx[i];
  ktmp[i] = keys[j];
  if (PAIRS) vtmp[i] = vals[j];
}
template <typename T, bool PAIRS>
__global__ void ms_writeback_kernel(const T *__restrict__ ktmp, const int *__restrict__ vtmp, size_t n, PortAcc<T> keys,
                                    PortAcc<int> vals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ktmp[i];
  if (PAIRS) vals[i] = vtmp[i];
}
template <typename T, bool PAIRS>
int merge_sort_impl(void *temp, size_t *temp_bytes, zpc_port keys, zpc_port vals, size_t n, cudaStream_t s) {
  using U = typename OrdKey<T>::U;
  if (!temp_bytes) return ZPCB200_E_BADARG;
  if (n > ((size_t)1 << 30)) return ZPCB200_E_UNSUPPORTED;
  size_t sort_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = radix_sort_impl<U, true>(nullptr, &sort_bytes, none, none, none, none, n, 0, (int)sizeof(U) * 8, s);
  if (rc) return rc;
  const size_t o_img = 0, o_idx = zpc_align_up(o_img + sizeof(U) * n, 256), o_img2 = zpc_align_up(o_idx + 4 * n, 256),
               o_idx2 = zpc_align_up(o_img2 + sizeof(U) * n, 256), o_kt = zpc_align_up(o_idx2 + 4 * n, 256),
               o_vt = zpc_align_up(o_kt + sizeof(T) * n, 256), o_sort = zpc_align_up(o_vt + 4 * n, 256), need = o_sort + sort_bytes;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (!n) return ZPCB200_OK;
  char *t = (char *)temp;
  U *img = (U *)(t + o_img), *img2 = (U *)(t + o_img2);
  int *idx = (int *)(t + o_idx), *idx2 = (int *)(t + o_idx2), *vt = (int *)(t + o_vt);
  T *kt = (T *)(t + o_kt);
  const unsigned g = (unsigned)((n + 255) / 256);
  ms_prepare_kernel<T><<<g, 256, 0, s>>>(PortAcc<T>(keys), n, img, idx);
  ZPC_CHECK_LAUNCH();
  zpc_port pi = {img, 0, 0, 0, 1}, px = {idx, 0, 0, 0, 1}, pi2 = {img2, 0, 0, 0, 1}, px2 = {idx2, 0, 0, 0, 1};
  size_t sb = sort_bytes;
  rc = radix_sort_impl<U, true>(t + o_sort, &sb, pi, px, pi2, px2, n, 0, (int)sizeof(U) * 8, s);
  if (rc) return rc;
  ms_gather_kernel<T, PAIRS><<<g, 256, 0, s>>>(PortAcc<T>(keys), PortAcc<int>(vals), idx2, n, kt, vt);
  ZPC_CHECK_LAUNCH();
  ms_writeback_kernel<T, PAIRS><<<g, 256, 0, s>>>(kt, vt, n, PortAcc<T>(keys), PortAcc<int>(vals));
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

#define ZPC_DEF_REDUCE_SCAN(S, T, TMAX, TLOW)                                                                   \
  int zpcb200_reduce_sum_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n, zpc_stream_t st) {     \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return reduce_impl<T, OpSum>(temp, tb, in, out, n, (T)0, (cudaStream_t)st);                                 \
  }                                                                                                             \
  int zpcb200_reduce_prod_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n, zpc_stream_t st) {    \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return reduce_impl<T, OpProd>(temp, tb, in, out, n, (T)1, (cudaStream_t)st);                                \
  }                                                                                                             \
  int zpcb200_reduce_min_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n, zpc_stream_t st) {     \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return reduce_impl<T, OpMin>(temp, tb, in, out, n, (T)TMAX, (cudaStream_t)st);                              \
  }                                                                                                             \
  int zpcb200_reduce_max_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n, zpc_stream_t st) {     \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return reduce_impl<T, OpMax>(temp, tb, in, out, n, (T)TLOW, (cudaStream_t)st);                              \
  }                                                                                                             \
  int zpcb200_exclusive_scan_sum_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n,               \
                                     zpc_stream_t st) {                                                         \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return scan_impl<T, false>(temp, tb, in, out, n, (cudaStream_t)st);                                         \
  }                                                                                                             \
  int zpcb200_inclusive_scan_sum_##S(void *temp, size_t *tb, zpc_port in, zpc_port out, size_t n,               \
                                     zpc_stream_t st) {                                                         \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return scan_impl<T, true>(temp, tb, in, out, n, (cudaStream_t)st);                                          \
  }
ZPC_DEF_REDUCE_SCAN(i32, int32_t, INT32_MAX, INT32_MIN)
ZPC_DEF_REDUCE_SCAN(u32, uint32_t, UINT32_MAX, 0u)
ZPC_DEF_REDUCE_SCAN(i64, int64_t, INT64_MAX, INT64_MIN)
ZPC_DEF_REDUCE_SCAN(f32, float, FLT_MAX, -FLT_MAX)
ZPC_DEF_REDUCE_SCAN(f64, double, DBL_MAX, -DBL_MAX)

#define ZPC_DEF_SORT(S, K)                                                                                      \
  int zpcb200_radix_sort_pair_##S(void *temp, size_t *tb, zpc_port ki, zpc_port vi, zpc_port ko, zpc_port vo,   \
                                  size_t n, int sbit, int ebit, zpc_stream_t st) {                              \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    return radix_sort_impl<K, true>(temp, tb, ki, vi, ko, vo, n, sbit, ebit, (cudaStream_t)st);                 \
  }                                                                                                             \
  int zpcb200_radix_sort_##S(void *temp, size_t *tb, zpc_port ki, zpc_port ko, size_t n, int sbit, int ebit,    \
                             zpc_stream_t st) {                                                                 \
    if (!tb) return ZPCB200_E_BADARG;                                                                           \
    zpc_port none = {nullptr, 0, 0, 0, 1};                                                                      \
    return radix_sort_impl<K, false>(temp, tb, ki, none, ko, none, n, sbit, ebit, (cudaStream_t)st);            \
  }
ZPC_DEF_SORT(u32, uint32_t)
ZPC_DEF_SORT(i32, int32_t)
ZPC_DEF_SORT(u64, uint64_t)

#define ZPC_DEF_MERGE_SORT(S, T)                                                                                \
  int zpcb200_merge_sort_pair_##S(void *temp, size_t *tb, zpc_port keys, zpc_port vals, size_t n, zpc_stream_t st) { \
    return merge_sort_impl<T, true>(temp, tb, keys, vals, n, (cudaStream_t)st);                                 \
  }                                                                                                             \
  int zpcb200_merge_sort_##S(void *temp, size_t *tb, zpc_port keys, size_t n, zpc_stream_t st) {                 \
    zpc_port none = {nullptr, 0, 0, 0, 1};                                                                      \
    return merge_sort_impl<T, false>(temp, tb, keys, none, n, (cudaStream_t)st);                                \
  }
ZPC_DEF_MERGE_SORT(i32, int32_t)
ZPC_DEF_MERGE_SORT(u32, uint32_t)
ZPC_DEF_MERGE_SORT(i64, int64_t)
ZPC_DEF_MERGE_SORT(u64, uint64_t)
ZPC_DEF_MERGE_SORT(f32, float)
ZPC_DEF_MERGE_SORT(f64, double)

const char *zpcb200_version(void) { return "zpcb200 0.1 (sm_100a)"; }
int zpcb200_kernel_launch_count(void) { return g_zpc_launches.load(); }
}

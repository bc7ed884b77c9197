// Shared pieces of the hash-grid partition build (simulation/sparsity/SparsityOp.hpp:41-112 restated for the GPU):
// block codes, the scratch open-addressing set with its append list, the particle marking pass (block side 2^S) and
// EnlargeSparsity.  Used by the legacy HashTable build (mpm.cu) and the bht / SparseGrid build (sparsegrid.cu); each
// finishes with its own placement kernel.
#pragma once
#include <climits>

#include "common.cuh"
#include "mpm_math.cuh"

namespace {

// ---- partition build --------------------------------------------------------------------------------
// Block codes: 3 x 10 bits in a uint32_t (block coordinates in [-512, 511] per axis: |cell coord| < 2 048 on side-4 blocks) — the
// default, what the measured path sorts — or 3 x 21 bits in a uint64_t (the *_wide entries: +-2^20 blocks per axis, the same
// packing zpcb200_halo_codes uses).  Codes compare like (x, y, z) tuples, so the sorted code list is the lexicographic key order.
template <class CODE> struct CodeTraits;
template <> struct CodeTraits<unsigned> {
  static constexpr int BITS = 10, SORT_BITS = 30;
  static constexpr unsigned EMPTY = 0xffffffffu;
};
template <> struct CodeTraits<unsigned long long> {
  static constexpr int BITS = 21, SORT_BITS = 63;
  static constexpr unsigned long long EMPTY = 0xffffffffffffffffull;
};
constexpr unsigned CODE_EMPTY = CodeTraits<unsigned>::EMPTY;
template <class CODE> __device__ __forceinline__ bool code_pack(int bx, int by, int bz, CODE &code) {
  constexpr int B = CodeTraits<CODE>::BITS;
  constexpr int BIAS = 1 << (B - 1);
  const unsigned ux = (unsigned)(bx + BIAS), uy = (unsigned)(by + BIAS), uz = (unsigned)(bz + BIAS);
  code = ((CODE)ux << (2 * B)) | ((CODE)uy << B) | (CODE)uz;
  return (ux | uy | uz) < (1u << B);
}
template <class CODE> __device__ __forceinline__ void code_unpack(CODE code, int &bx, int &by, int &bz) {
  constexpr int B = CodeTraits<CODE>::BITS;
  constexpr int BIAS = 1 << (B - 1);
  constexpr unsigned M = (1u << B) - 1u;
  bx = (int)(unsigned)(code >> (2 * B)) - BIAS;
  by = (int)((unsigned)(code >> B) & M) - BIAS;
  bz = (int)((unsigned)code & M) - BIAS;
}
__device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ unsigned mix_code(unsigned c) { return mix32(c); }
__device__ __forceinline__ unsigned mix_code(unsigned long long c) { return mix32((unsigned)c ^ mix32((unsigned)(c >> 32))); }
// insert code into the scratch set; the first inserter appends it to list
template <class CODE>
__device__ __forceinline__ void set_insert(CODE code, CODE *set, unsigned set_mask, CODE *list,
                                           int list_cap, int *list_cnt, int *overflow) {
  constexpr CODE EMPTY = CodeTraits<CODE>::EMPTY;
  unsigned slot = mix_code(code) & set_mask;
  for (unsigned probes = 0; probes <= set_mask; ++probes) {
    CODE cur = set[slot];
    if (cur == code) return;
    if (cur == EMPTY) {
      cur = atomicCAS(&set[slot], EMPTY, code);
      if (cur == EMPTY) {
        const int i = atomicAdd(list_cnt, 1);
        if (i < list_cap) list[i] = code;
        else if (overflow) *overflow = 1;
        return;
      }
      if (cur == code) return;
    }
    slot = (slot + 1) & set_mask;
  }
  if (overflow) *overflow = 1;
}

// scratch of the partition build: the code set, the code list (padded with EMPTY so that it sorts to the end), counters
template <class CODE>
__global__ void part_clear_scratch_kernel(CODE *set, unsigned set_n, CODE *list, int list_cap, int *counters) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < set_n; i += stride) set[i] = CodeTraits<CODE>::EMPTY;
  for (size_t i = t0; i < (size_t)list_cap; i += stride) list[i] = CodeTraits<CODE>::EMPTY;
  if (t0 < 4) counters[t0] = 0;
}

template <int S, class CODE>  // block side 2^S
__global__ void part_mark_kernel(PortAcc<const float> x, size_t n, float dxinv, CODE *set, unsigned set_mask,
                                 CODE *list, int list_cap, int *list_cnt, int *overflow) {
  constexpr CODE EMPTY = CodeTraits<CODE>::EMPTY;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
  for (size_t i0 = first; i0 < n; i0 += stride) {  // warp-uniform trip count
    const size_t i = i0 + (threadIdx.x & 31);
    CODE code = EMPTY;
    if (i < n) {
      const int bx = zpcm::sparsity_coord(x.at(i, 0), dxinv) >> S;
      const int by = zpcm::sparsity_coord(x.at(i, 1), dxinv) >> S;
      const int bz = zpcm::sparsity_coord(x.at(i, 2), dxinv) >> S;
      if (!code_pack(bx, by, bz, code)) { code = EMPTY; if (overflow) *overflow = 1; }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, code);
    if (code != EMPTY && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
      set_insert(code, set, set_mask, list, list_cap, list_cnt, overflow);
  }
}
// EnlargeSparsity{lo, hi} (SparsityOp.hpp:88-112): every block present after the particle pass adds its
// neighbours at offsets [lo, hi)^3; the reference's substep uses {0, 2}
template <class CODE>
__global__ void part_enlarge_kernel(CODE *set, unsigned set_mask, CODE *list, int list_cap,
                                    const int *cnt_before, int *list_cnt, int lo, int hi, int *overflow) {
  const int n0 = min(*cnt_before, list_cap);
  const int w = hi - lo, w3 = w * w * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n0 * w3; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / w3), o = (int)(t % w3);
    const int ox = lo + o / (w * w), oy = lo + (o / w) % w, oz = lo + o % w;
    if ((ox | oy | oz) == 0) continue;
    int bx, by, bz;
    code_unpack(list[i], bx, by, bz);
    CODE code;
    if (code_pack(bx + ox, by + oy, bz + oz, code))
      set_insert(code, set, set_mask, list, list_cap, list_cnt, overflow);
    else if (overflow) *overflow = 1;
  }
}
__global__ void part_snapshot_kernel(const int *src, int *dst) { *dst = *src; }

template <class CODE> inline int sort_codes(void *temp, size_t *bytes, CODE *in, CODE *out, size_t n, cudaStream_t s);
template <> inline int sort_codes<unsigned>(void *temp, size_t *bytes, unsigned *in, unsigned *out, size_t n, cudaStream_t s) {
  zpc_port pl = {in, 0, 0, 0, 1}, ps = {out, 0, 0, 0, 1};
  return zpcb200_radix_sort_u32(temp, bytes, pl, ps, n, 0, CodeTraits<unsigned>::SORT_BITS, (zpc_stream_t)s);
}
template <> inline int sort_codes<unsigned long long>(void *temp, size_t *bytes, unsigned long long *in, unsigned long long *out, size_t n,
                                                       cudaStream_t s) {
  zpc_port pl = {in, 0, 0, 0, 1}, ps = {out, 0, 0, 0, 1};
  // the padding (EMPTY = all ones) must sort to the end: all 64 bits
  return zpcb200_radix_sort_u64(temp, bytes, pl, ps, n, 0, 64, (zpc_stream_t)s);
}

// sizes of the scratch the shared passes need; the sort scratch follows at off_sort
struct PartScratch {
  size_t set_n, off_set, off_list, off_sorted, off_sort, sort_bytes, need;
  int list_cap;
};
template <class CODE = unsigned>
inline int part_scratch_layout(size_t table_size, PartScratch &L) {
  L.set_n = 1;
  while (L.set_n < table_size / 2) L.set_n <<= 1;
  if (L.set_n < 1024) L.set_n = 1024;
  L.list_cap = table_size / 8 > 64 ? (int)(table_size / 8) : 64;
  L.sort_bytes = 0;
  const int rc = sort_codes<CODE>(nullptr, &L.sort_bytes, nullptr, nullptr, (size_t)L.list_cap, nullptr);
  if (rc) return rc;
  L.off_set = 256;
  L.off_list = zpc_align_up(L.off_set + sizeof(CODE) * L.set_n, 256);
  L.off_sorted = zpc_align_up(L.off_list + sizeof(CODE) * (size_t)L.list_cap, 256);
  L.off_sort = zpc_align_up(L.off_sorted + sizeof(CODE) * (size_t)L.list_cap, 256);
  L.need = L.off_sort + L.sort_bytes;
  return ZPCB200_OK;
}
// mark -> enlarge -> sort of the block codes; leaves the sorted codes at temp + off_sorted and their number in
// counters[0] (= (int*)temp); the caller has already cleared its table and launches its placement kernel afterwards
template <int S, class CODE = unsigned>
int part_collect_sorted(char *t, const PartScratch &L, zpc_port x, size_t n, float dx, int enlarge_lo, int enlarge_hi, int *overflow,
                        cudaStream_t s) {
  int *counters = (int *)t;  // [0] list count, [1] count before enlarge
  CODE *set = (CODE *)(t + L.off_set), *list = (CODE *)(t + L.off_list), *sorted = (CODE *)(t + L.off_sorted);
  const int G = ZPC_SM_COUNT * 8;
  part_clear_scratch_kernel<CODE><<<G, 256, 0, s>>>(set, (unsigned)L.set_n, list, L.list_cap, counters);
  ZPC_CHECK_LAUNCH();
  if (n) {
    part_mark_kernel<S, CODE><<<G, 256, 0, s>>>(PortAcc<const float>(x), n, 1.0f / dx, set, (unsigned)(L.set_n - 1), list, L.list_cap,
                                                 counters, overflow);
    ZPC_CHECK_LAUNCH();
  }
  part_snapshot_kernel<<<1, 1, 0, s>>>(counters, counters + 1);
  ZPC_CHECK_LAUNCH();
  if (enlarge_hi - enlarge_lo > 0) {
    part_enlarge_kernel<CODE><<<G, 256, 0, s>>>(set, (unsigned)(L.set_n - 1), list, L.list_cap, counters + 1, counters, enlarge_lo,
                                                 enlarge_hi, overflow);
    ZPC_CHECK_LAUNCH();
  }
  size_t sb = L.sort_bytes;
  return sort_codes<CODE>(t + L.off_sort, &sb, list, sorted, (size_t)L.list_cap, s);
}

}  // namespace

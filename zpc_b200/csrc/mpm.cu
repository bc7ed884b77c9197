// MPM path on the reference's own data layout (AoS Particles, HashTable<i32,3,int>, Grids<f32,3,4>).
//
//   partition_build : CleanSparsity + ComputeSparsity + EnlargeSparsity{0,2}
//                     (reference simulation/sparsity/SparsityOp.hpp:41-112) with DETERMINISTIC block
//                     numbering: dedupe block codes in a scratch hash set, radix-sort them (prims.cu),
//                     index = rank, then place every key in the legacy table so that the unmodified
//                     HashTableView::query (container/HashTable.hpp:447-456) finds it.
//   clean_grid / grid_update : simulation/grid/GridOp.hpp:54-110.
//   p2g_apic_fcr / g2p_apic  : simulation/transfer/{P2G,G2P}.hpp on AoS particles in ANY order — the
//                     drop-in for the reference's generic functors under cuda_c (one thread per particle,
//                     <= 8 hash queries per particle instead of 27, float REDs to L2).  The fast path for
//                     block-sorted AoSoA particles is mpm_binned.cu.
#include <climits>

#include "common.cuh"
#include "mpm_math.cuh"
#include "mpm_particle.cuh"
#include "mpm_kernels.cuh"
#include "partition.cuh"

namespace {

// ---- partition build (shared passes: partition.cuh) ------------------------------------------------------
__global__ void part_clear_table_kernel(int table_size, int *keys, int *indices, int *status) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)table_size * 3; i += stride) keys[i] = INT_MAX;
  for (size_t i = t0; i < (size_t)table_size; i += stride) { indices[i] = -1; status[i] = -1; }
}

// place key i (rank order) at the first free slot of ITS OWN probe sequence (any such placement is a valid table)
template <class CODE>
__global__ void part_place_kernel(const CODE *sorted, const int *list_cnt, int list_cap, int table_size, int *keys,
                                  int *indices, int *active_keys, int *cnt, int *overflow) {
  const int n = min(*list_cnt, list_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int bx, by, bz;
    code_unpack(sorted[i], bx, by, bz);
    int slot = zpcm::hash_slot0(bx, by, bz, table_size);
    int probes = 0;
    while (atomicCAS(&indices[slot], -1, i) != -1) {
      slot = (slot + 127) % table_size;
      if (++probes > table_size) { if (overflow) *overflow = 1; break; }
    }
    if (probes <= table_size) {
      keys[3 * (size_t)slot] = bx; keys[3 * (size_t)slot + 1] = by; keys[3 * (size_t)slot + 2] = bz;
    }
    active_keys[3 * (size_t)i] = bx; active_keys[3 * (size_t)i + 1] = by; active_keys[3 * (size_t)i + 2] = bz;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cnt = n;
}

// ---- index_buckets_for_particles (simulation/particle/Query.tpp:9-58) ------------------------------------------------
// cell of a particle: ComputeSparsity / SpatiallyCount / SpatiallyDistribute with blockLen 1, offset 0 (SparsityOp.hpp:71-76)
__device__ __forceinline__ int bucket_cell(float x, float dxinv, float displacement) { return (int)floorf(x * dxinv + displacement); }
template <class CODE>
__global__ void bucket_mark_kernel(PortAcc<const float> x, size_t n, float dxinv, float displacement, CODE *set, unsigned set_mask,
                                   CODE *list, int list_cap, int *list_cnt, int *overflow) {
  constexpr CODE EMPTY = CodeTraits<CODE>::EMPTY;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
  for (size_t i0 = first; i0 < n; i0 += stride) {  // warp-uniform trip count
    const size_t i = i0 + (threadIdx.x & 31);
    CODE code = EMPTY;
    if (i < n && !code_pack(bucket_cell(x.at(i, 0), dxinv, displacement), bucket_cell(x.at(i, 1), dxinv, displacement),
                            bucket_cell(x.at(i, 2), dxinv, displacement), code)) {
      code = EMPTY;
      if (overflow) *overflow = 1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, code);
    if (code != EMPTY && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
      set_insert(code, set, set_mask, list, list_cap, list_cnt, overflow);
  }
}
// key = bucket number of the particle's cell, value = particle id; counts the bucket (SpatiallyCount, SparsityOp.hpp:115-151)
__global__ void bucket_keys_kernel(PortAcc<const float> x, size_t n, float dxinv, float displacement, zpc_hashtable_view tb, unsigned *keys,
                                   int *vals, int *counts, int *overflow) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = zpcm::table_query(bucket_cell(x.at(i, 0), dxinv, displacement), bucket_cell(x.at(i, 1), dxinv, displacement),
                            bucket_cell(x.at(i, 2), dxinv, displacement), tb.tableSize, tb.keys, tb.indices);
  if (b < 0) { if (overflow) *overflow = 1; b = 0; }
  keys[i] = (unsigned)b;
  vals[i] = (int)i;
  atomicAdd(&counts[b], 1);
}

using zpcm::collide;  // Collider::resolveCollision over the analytic level sets: mpm_math.cuh (host + device)

// ApplyBoundaryConditionOnGridBlocks with a static analytic collider: one thread per (block, cell)
__global__ void __launch_bounds__(256) apply_boundary_kernel(float *tiles, const int *__restrict__ active_keys, const int *cnt,
                                                             int nch, size_t cap_blocks, float dx, zpc_collider col) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const int cell = threadIdx.x & 63;
  const int cx = (cell >> 4) & 3, cy = (cell >> 2) & 3, cz = cell & 3;
  for (size_t b = (size_t)blockIdx.x * 4 + (threadIdx.x >> 6); b < nb; b += (size_t)gridDim.x * 4) {
    float *t = tiles + b * (size_t)nch * 64;
    if (!(t[cell] > 0.f)) continue;
    const float px = ((float)active_keys[3 * b] * 4.f + (float)cx) * dx, py = ((float)active_keys[3 * b + 1] * 4.f + (float)cy) * dx,
                pz = ((float)active_keys[3 * b + 2] * 4.f + (float)cz) * dx;
    float vx = t[64 + cell], vy = t[128 + cell], vz = t[192 + cell];
    const float ox = vx, oy = vy, oz = vz;
    collide(col, px, py, pz, vx, vy, vz);
    if (vx != ox || vy != oy || vz != oz) { t[64 + cell] = vx; t[128 + cell] = vy; t[192 + cell] = vz; }
  }
}

// GridMomentumToVelocity (GridOp.hpp:184-214): v = mv * (1/m) where m != 0, max |v|^2; one thread per (block, cell)
__global__ void __launch_bounds__(256) grid_momentum_to_velocity_kernel(float *tiles, const int *cnt, int nch, size_t cap_blocks, int m_chn,
                                                                        int mv_chn, float *max_vel_sqr) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  float mx = 0.f;
  for (size_t gc = (size_t)blockIdx.x * 256 + threadIdx.x; gc < nb * 64; gc += (size_t)gridDim.x * 256) {
    float *t = tiles + (gc >> 6) * (size_t)nch * 64 + (gc & 63);
    float mass = t[m_chn * 64];
    if (mass != 0.f) {
      mass = 1.f / mass;
      const float vx = t[mv_chn * 64] * mass, vy = t[(mv_chn + 1) * 64] * mass, vz = t[(mv_chn + 2) * 64] * mass;
      t[mv_chn * 64] = vx; t[(mv_chn + 1) * 64] = vy; t[(mv_chn + 2) * 64] = vz;
      mx = fmaxf(mx, vx * vx + vy * vy + vz * vz);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  __shared__ float smx[8];
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, smx[i]);
    if (mx > 0.f) atomicMax((int *)max_vel_sqr, __float_as_int(mx));
  }
}

// GridAngularMomentum (GridOp.hpp:216-262): sum6 += {x cross mv, mv} over the cells with mass; the per-cell products in float
// without contraction (the host build of the functor), the sums in double: warp shuffle -> shared -> one atomicAdd(double) per
// CTA and component instead of the reference's six atomics per cell.
__global__ void __launch_bounds__(256) grid_angular_momentum_kernel(const float *__restrict__ tiles, const int *__restrict__ active_keys,
                                                                    const int *cnt, int nch, size_t cap_blocks, float dx, int m_chn,
                                                                    int mv_chn, double *sum6) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const int cell = threadIdx.x & 63;
  const int cx = (cell >> 4) & 3, cy = (cell >> 2) & 3, cz = cell & 3;
  double acc[6] = {0., 0., 0., 0., 0., 0.};
  for (size_t b = (size_t)blockIdx.x * 4 + (threadIdx.x >> 6); b < nb; b += (size_t)gridDim.x * 4) {
    const float *t = tiles + b * (size_t)nch * 64 + cell;
    if (t[m_chn * 64] == 0.f) continue;
    const float px = zpcm::rn_mul(zpcm::rn_add(zpcm::rn_mul((float)active_keys[3 * b], 4.f), (float)cx), dx),
                py = zpcm::rn_mul(zpcm::rn_add(zpcm::rn_mul((float)active_keys[3 * b + 1], 4.f), (float)cy), dx),
                pz = zpcm::rn_mul(zpcm::rn_add(zpcm::rn_mul((float)active_keys[3 * b + 2], 4.f), (float)cz), dx);
    const float mx = t[mv_chn * 64], my = t[(mv_chn + 1) * 64], mz = t[(mv_chn + 2) * 64];
    acc[0] += (double)zpcm::rn_sub(zpcm::rn_mul(py, mz), zpcm::rn_mul(pz, my));
    acc[1] += (double)zpcm::rn_sub(zpcm::rn_mul(pz, mx), zpcm::rn_mul(px, mz));
    acc[2] += (double)zpcm::rn_sub(zpcm::rn_mul(px, my), zpcm::rn_mul(py, mx));
    acc[3] += (double)mx; acc[4] += (double)my; acc[5] += (double)mz;
  }
  __shared__ double ssum[8][6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = acc[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += ssum[w][threadIdx.x];
    if (v != 0.) atomicAdd(sum6 + threadIdx.x, v);
  }
}

// ComputeGridBlockVelocity followed by ApplyBoundaryConditionOnGridBlocks for up to ZPCB200_MAX_COLLIDERS colliders in
// ONE pass over the grid: the velocity never leaves the registers between the two functors.
// max |v|^2 is taken before the projection, as the reference's sequence of functors does.
struct ColliderSet { zpc_collider c[ZPCB200_MAX_COLLIDERS]; int n; };
__global__ void __launch_bounds__(256) grid_update_bc_kernel(float *tiles, const int *__restrict__ active_keys, const int *cnt, int nch,
                                                             size_t cap_blocks, float dx, float dt, float ex, float ey, float ez, int mode,
                                                             ColliderSet cols, float *max_vel_sqr, zpc_halo_view halo) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const int cell = threadIdx.x & 63;
  const int cx = (cell >> 4) & 3, cy = (cell >> 2) & 3, cz = cell & 3;
  float mx = 0.f;
  for (size_t b = (size_t)blockIdx.x * 4 + (threadIdx.x >> 6); b < nb; b += (size_t)gridDim.x * 4) {
    float *t = tiles + b * (size_t)nch * 64;
    if (halo.peer) {
      // fused halo receive: add what the ranks sharing this block sent (ascending rank: fixed order), leave the complete sums in the
      // grid like a single-GPU run would, hand the slots back zeroed.  Every load is issued before the first store (one vector load
      // for the four peer ranks, one for their slots, then up to 4 x 7 + 7 independent loads): the kernel is latency-bound otherwise
      // (ncu: 3 dependent round trips per peer, 15 % of the DRAM bandwidth).
      const int4 pr = *reinterpret_cast<const int4 *>(halo.peer + b * ZPCB200_HALO_K);
      if (pr.x >= 0) {
        const int4 ps = *reinterpret_cast<const int4 *>(halo.pos + b * ZPCB200_HALO_K);
        const int rk[4] = {pr.x, pr.y, pr.z, pr.w}, pk[4] = {ps.x, ps.y, ps.z, ps.w};
        float *src[4];
        float v[4][7], own[7];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          src[k] = halo.recv + (((size_t)halo.half * halo.world + (rk[k] >= 0 ? rk[k] : 0)) * halo.seg + pk[k]) * 448 + cell;
#pragma unroll
          for (int c = 0; c < 7; ++c) v[k][c] = rk[k] >= 0 ? __ldcg(src[k] + c * 64) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) own[c] = t[c * 64 + cell];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (rk[k] >= 0) {
#pragma unroll
            for (int c = 0; c < 7; ++c) { own[c] += v[k][c]; src[k][c * 64] = 0.f; }
          }
#pragma unroll
        for (int c = 0; c < 7; ++c) t[c * 64 + cell] = own[c];
      }
    }
    float mass = t[cell];
    if (mass != 0.f) {
      float mvx = t[64 + cell], mvy = t[128 + cell], mvz = t[192 + cell];
      if (mode == 1) { mvx += t[256 + cell]; mvy += t[320 + cell]; mvz += t[384 + cell]; }
      const float minv = 1.f / mass;
      float vx = mvx * minv + ex * dt, vy = mvy * minv + ey * dt, vz = mvz * minv + ez * dt;
      mx = fmaxf(mx, vx * vx + vy * vy + vz * vz);
      if (mass > 0.f) {
        const float px = ((float)active_keys[3 * b] * 4.f + (float)cx) * dx, py = ((float)active_keys[3 * b + 1] * 4.f + (float)cy) * dx,
                    pz = ((float)active_keys[3 * b + 2] * 4.f + (float)cz) * dx;
        for (int k = 0; k < cols.n; ++k) collide(cols.c[k], px, py, pz, vx, vy, vz);
      }
      t[64 + cell] = vx; t[128 + cell] = vy; t[192 + cell] = vz;
    } else if (mode == 1) {
      t[64 + cell] += t[256 + cell]; t[128 + cell] += t[320 + cell]; t[192 + cell] += t[384 + cell];
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  __shared__ float smx[8];
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, smx[i]);
    if (mx > 0.f) atomicMax((int *)max_vel_sqr, __float_as_int(mx));
  }
}

// ---- halo pack / unpack ---------------------------------------------------------------------------------
template <int MODE>  // 0 pack, 1 unpack-add, 2 unpack-set
__global__ void halo_kernel(float *tiles, int nch_grid, const int *ids, int n, int chn0, int nch, float *buf) {
  const int per = nch * 64;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)n * per; t += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / per), r = (int)(t % per);
    float *g = tiles + ((size_t)ids[b] * nch_grid + chn0) * 64 + r;
    if (MODE == 0) buf[t] = *g;
    else if (MODE == 1) *g += buf[t];
    else *g = buf[t];
  }
}

}  // namespace

extern "C" {

}  // extern "C"
template <class CODE>
static int partition_build_impl(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, zpc_hashtable_view tb,
                                int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream) {
  if (!temp_bytes || tb.tableSize <= 0 || enlarge_hi < enlarge_lo || enlarge_hi - enlarge_lo > 8) return ZPCB200_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  // scratch set: power of two >= tableSize/2 ; list capacity = tableSize/8 block codes
  PartScratch L;
  int rc = part_scratch_layout<CODE>((size_t)tb.tableSize, L);
  if (rc) return rc;
  if (!temp) { *temp_bytes = L.need; return ZPCB200_OK; }
  if (*temp_bytes < L.need) return ZPCB200_E_TEMP_TOO_SMALL;
  char *t = (char *)temp;
  const int G = ZPC_SM_COUNT * 8;
  part_clear_table_kernel<<<G, 256, 0, s>>>(tb.tableSize, tb.keys, tb.indices, tb.status);
  ZPC_CHECK_LAUNCH();
  rc = part_collect_sorted<2, CODE>(t, L, x, n, dx, enlarge_lo, enlarge_hi, overflow, s);
  if (rc) return rc;
  part_place_kernel<CODE><<<G, 256, 0, s>>>((const CODE *)(t + L.off_sorted), (const int *)t, L.list_cap, tb.tableSize, tb.keys,
                                             tb.indices, tb.activeKeys, tb.cnt, overflow);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
extern "C" {
int zpcb200_partition_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, zpc_hashtable_view tb,
                            int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream) {
  return partition_build_impl<unsigned>(temp, temp_bytes, x, n, dx, tb, enlarge_lo, enlarge_hi, overflow, stream);
}
/* 64-bit block codes: block coordinates in [-2^20, 2^20) per axis instead of [-512, 511] */
int zpcb200_partition_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, zpc_hashtable_view tb,
                                 int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream) {
  return partition_build_impl<unsigned long long>(temp, temp_bytes, x, n, dx, tb, enlarge_lo, enlarge_hi, overflow, stream);
}

static int bit_length_u(size_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

}  // extern "C"
template <class CODE>
static int index_buckets_build_impl(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, float displacement,
                                    zpc_hashtable_view tb, int *counts, int *offsets, int *indices, int *overflow, zpc_stream_t stream) {
  if (!temp_bytes || tb.tableSize <= 0) return ZPCB200_E_BADARG;
  if (n > ((size_t)1 << 30)) return ZPCB200_E_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  PartScratch L;
  int rc = part_scratch_layout<CODE>((size_t)tb.tableSize, L);
  if (rc) return rc;
  const int ebit = bit_length_u(n ? n - 1 : 0) > 0 ? bit_length_u(n - 1) : 1;   // bucket numbers are < n
  size_t sort_bytes = 0, scan_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  rc = zpcb200_radix_sort_pair_u32(nullptr, &sort_bytes, none, none, none, none, n, 0, ebit, nullptr);
  if (rc) return rc;
  rc = zpcb200_exclusive_scan_sum_i32(nullptr, &scan_bytes, none, none, n + 1, nullptr);
  if (rc) return rc;
  const size_t o_keys = zpc_align_up(L.need, 256), o_vals = zpc_align_up(o_keys + 4 * n, 256), o_skeys = zpc_align_up(o_vals + 4 * n, 256),
               o_scan = zpc_align_up(o_skeys + 4 * n, 256), o_sort = zpc_align_up(o_scan + scan_bytes, 256), need = o_sort + sort_bytes;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (!tb.keys || !tb.indices || !tb.status || !tb.activeKeys || !tb.cnt || !counts || !offsets || (n && (!indices || !x.base))) return ZPCB200_E_BADARG;
  char *t = (char *)temp;
  const int G = ZPC_SM_COUNT * 8;
  // table of the occupied cells: clear -> mark -> sort the cell codes -> place (bucket number = rank of the cell key)
  part_clear_table_kernel<<<G, 256, 0, s>>>(tb.tableSize, tb.keys, tb.indices, tb.status);
  ZPC_CHECK_LAUNCH();
  int *counters = (int *)t;
  CODE *set = (CODE *)(t + L.off_set), *list = (CODE *)(t + L.off_list), *sorted = (CODE *)(t + L.off_sorted);
  part_clear_scratch_kernel<CODE><<<G, 256, 0, s>>>(set, (unsigned)L.set_n, list, L.list_cap, counters);
  ZPC_CHECK_LAUNCH();
  const float dxinv = 1.0f / dx;
  if (n) {
    bucket_mark_kernel<CODE><<<G, 256, 0, s>>>(PortAcc<const float>(x), n, dxinv, displacement, set, (unsigned)(L.set_n - 1), list, L.list_cap,
                                                counters, overflow);
    ZPC_CHECK_LAUNCH();
  }
  {
    size_t sb = L.sort_bytes;
    rc = sort_codes<CODE>(t + L.off_sort, &sb, list, sorted, (size_t)L.list_cap, s);
    if (rc) return rc;
  }
  part_place_kernel<CODE><<<G, 256, 0, s>>>(sorted, counters, L.list_cap, tb.tableSize, tb.keys, tb.indices, tb.activeKeys, tb.cnt, overflow);
  ZPC_CHECK_LAUNCH();
  // counts (n + 1 entries: buckets beyond table.size() stay 0), offsets = exclusive scan, indices = ids sorted by bucket (stable)
  ZPC_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (n + 1), s));
  unsigned *keys = (unsigned *)(t + o_keys), *skeys = (unsigned *)(t + o_skeys);
  int *vals = (int *)(t + o_vals);
  if (n) {
    bucket_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(PortAcc<const float>(x), n, dxinv, displacement, tb, keys, vals, counts, overflow);
    ZPC_CHECK_LAUNCH();
  }
  {
    zpc_port pi = {counts, 0, 0, 0, 1}, po = {offsets, 0, 0, 0, 1};
    size_t sb = scan_bytes;
    rc = zpcb200_exclusive_scan_sum_i32(t + o_scan, &sb, pi, po, n + 1, stream);
    if (rc) return rc;
  }
  if (n) {
    zpc_port pk = {keys, 0, 0, 0, 1}, pv = {vals, 0, 0, 0, 1}, psk = {skeys, 0, 0, 0, 1}, psv = {indices, 0, 0, 0, 1};
    size_t sb = sort_bytes;
    rc = zpcb200_radix_sort_pair_u32(t + o_sort, &sb, pk, pv, psk, psv, n, 0, ebit, stream);
    if (rc) return rc;
  }
  return ZPCB200_OK;
}
extern "C" {
int zpcb200_index_buckets_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, float displacement,
                                zpc_hashtable_view tb, int *counts, int *offsets, int *indices, int *overflow, zpc_stream_t stream) {
  return index_buckets_build_impl<unsigned>(temp, temp_bytes, x, n, dx, displacement, tb, counts, offsets, indices, overflow, stream);
}
/* 64-bit cell codes: cell coordinates in [-2^20, 2^20) per axis instead of [-512, 511] */
int zpcb200_index_buckets_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, float displacement,
                                     zpc_hashtable_view tb, int *counts, int *offsets, int *indices, int *overflow, zpc_stream_t stream) {
  return index_buckets_build_impl<unsigned long long>(temp, temp_bytes, x, n, dx, displacement, tb, counts, offsets, indices, overflow, stream);
}

int zpcb200_clean_grid(zpc_grids_view g, const int *cnt, zpc_stream_t stream) {
  if (!g.tiles || !cnt) return ZPCB200_E_BADARG;
  clean_grid_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((float4 *)g.tiles, cnt, g.numChannels, g.numBlocks, 64);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_grid_update(zpc_grids_view g, const int *cnt, float dt, const float extf[3], int mode, float *maxVelSqr,
                        zpc_stream_t stream) {
  if (!g.tiles || !cnt || !maxVelSqr || g.numChannels < (mode == 1 ? 7 : 4)) return ZPCB200_E_BADARG;
  grid_update_kernel<64><<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, cnt, g.numChannels, g.numBlocks, dt, extf[0],
                                                                        extf[1], extf[2], mode, maxVelSqr);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_apply_boundary(zpc_grids_view g, zpc_hashtable_view tb, zpc_collider col, zpc_stream_t stream) {
  if (!g.tiles || !tb.activeKeys || !tb.cnt || g.numChannels < 4 || (unsigned)col.geometry > 2u || (unsigned)col.type > 2u || !(col.s > 0.f))
    return ZPCB200_E_BADARG;
  apply_boundary_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, tb.activeKeys, tb.cnt, g.numChannels, g.numBlocks,
                                                                           g.dx, col);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_grid_momentum_to_velocity(zpc_grids_view g, const int *cnt, int mChn, int mvChn, float *maxVelSqr, zpc_stream_t stream) {
  if (!g.tiles || !cnt || !maxVelSqr || mChn < 0 || mvChn < 0 || mChn >= g.numChannels || mvChn + 3 > g.numChannels ||
      (mChn >= mvChn && mChn < mvChn + 3))
    return ZPCB200_E_BADARG;
  grid_momentum_to_velocity_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, cnt, g.numChannels, g.numBlocks, mChn, mvChn,
                                                                                      maxVelSqr);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_grid_angular_momentum(zpc_grids_view g, zpc_hashtable_view tb, int mChn, int mvChn, double *sum6, zpc_stream_t stream) {
  if (!g.tiles || !tb.activeKeys || !tb.cnt || !sum6 || mChn < 0 || mvChn < 0 || mChn >= g.numChannels || mvChn + 3 > g.numChannels)
    return ZPCB200_E_BADARG;
  grid_angular_momentum_kernel<<<ZPC_SM_COUNT * 4, 256, 0, (cudaStream_t)stream>>>(g.tiles, tb.activeKeys, tb.cnt, g.numChannels, g.numBlocks,
                                                                                  g.dx, mChn, mvChn, sum6);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

zpc_collider zpcb200_collider_static(int geometry, int type, const float origin[3], const float normal_or_radius[3]) {
  zpc_collider c = {};
  c.geometry = geometry;
  c.type = type;
  for (int d = 0; d < 3; ++d) { c.origin[d] = origin[d]; c.normal[d] = normal_or_radius[d]; }
  c.R[0] = c.R[4] = c.R[8] = 1.f;
  c.s = 1.f;
  return c;
}

int zpcb200_grid_update_bc(zpc_grids_view g, zpc_hashtable_view tb, float dt, const float extf[3], int mode,
                           const zpc_collider *colliders, int ncolliders, float *maxVelSqr, zpc_stream_t stream) {
  if (!g.tiles || !tb.activeKeys || !tb.cnt || !extf || !maxVelSqr || (mode != 0 && mode != 1) || g.numChannels < (mode ? 7 : 4) ||
      ncolliders < 0 || ncolliders > ZPCB200_MAX_COLLIDERS || (ncolliders && !colliders))
    return ZPCB200_E_BADARG;
  ColliderSet cs;
  cs.n = ncolliders;
  for (int k = 0; k < ncolliders; ++k) {
    if ((unsigned)colliders[k].geometry > 2u || (unsigned)colliders[k].type > 2u || !(colliders[k].s > 0.f)) return ZPCB200_E_BADARG;
    cs.c[k] = colliders[k];
  }
  grid_update_bc_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, tb.activeKeys, tb.cnt, g.numChannels, g.numBlocks, g.dx,
                                                                           dt, extf[0], extf[1], extf[2], mode, cs, maxVelSqr, zpc_halo_view{});
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_grid_update_halo(zpc_grids_view g, zpc_hashtable_view tb, float dt, const float extf[3], int mode,
                             const zpc_collider *colliders, int ncolliders, float *maxVelSqr, zpc_halo_view halo, zpc_stream_t stream) {
  if (!g.tiles || !tb.activeKeys || !tb.cnt || !extf || !maxVelSqr || (mode != 0 && mode != 1) || g.numChannels != 7 ||
      ncolliders < 0 || ncolliders > ZPCB200_MAX_COLLIDERS || (ncolliders && !colliders))
    return ZPCB200_E_BADARG;
  if (halo.peer && (!halo.pos || !halo.recv || halo.world < 1 || halo.world > ZPCB200_HALO_MAX_PEERS || (unsigned)halo.rank >= (unsigned)halo.world ||
                    halo.seg <= 0 || (halo.half != 0 && halo.half != 1)))
    return ZPCB200_E_BADARG;
  ColliderSet cs;
  cs.n = ncolliders;
  for (int k = 0; k < ncolliders; ++k) {
    if ((unsigned)colliders[k].geometry > 2u || (unsigned)colliders[k].type > 2u || !(colliders[k].s > 0.f)) return ZPCB200_E_BADARG;
    cs.c[k] = colliders[k];
  }
  grid_update_bc_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, tb.activeKeys, tb.cnt, g.numChannels, g.numBlocks, g.dx,
                                                                           dt, extf[0], extf[1], extf[2], mode, cs, maxVelSqr, halo);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_fcr(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_fixed_corotated model,
                         zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;  // empty range: nothing to launch (pointers of an empty container may be null)
  if (g.numChannels != 7 || !P.X || !P.V || !P.M || !P.C || !P.F) return ZPCB200_E_BADARG;
  float mu, lam;
  zpcm::lame_host(model.E, model.nu, mu, lam);
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt, model.volume, mu, lam);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_vonmises(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt,
                              zpc_vonmises_fixed_corotated model, zpc_stream_t stream) {
  if (g.numChannels != 7 || !g.tiles || !tb.keys || !tb.indices) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.M || !P.C || !P.F)) return ZPCB200_E_BADARG;
  if (!P.count) return ZPCB200_OK;
  float mu, lam;
  zpcm::lame_host(model.E, model.nu, mu, lam);
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_vm_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt, model.volume, mu,
                                                            lam, model.yieldStress);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_drucker_prager(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt,
                                    zpc_drucker_prager model, zpc_stream_t stream) {
  if (g.numChannels != 7 || !g.tiles || !tb.keys || !tb.indices) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.M || !P.C || !P.F || !P.logJp)) return ZPCB200_E_BADARG;
  if (!P.count) return ZPCB200_OK;
  float mu, lam;
  zpcm::lame_host(model.E, model.nu, mu, lam);
  const PlasticParams prm{model.cohesion, model.beta, model.yieldSurface, 0.f, model.volumeCorrection};
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_plastic_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt,
                                                                    model.volume, mu, lam, prm);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_nacc(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_nacc model,
                          zpc_stream_t stream) {
  if (g.numChannels != 7 || !g.tiles || !tb.keys || !tb.indices || model.dim != 3) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.M || !P.C || !P.F || !P.logJp)) return ZPCB200_E_BADARG;
  if (!P.count) return ZPCB200_OK;
  float mu, lam;
  zpcm::lame_host(model.E, model.nu, mu, lam);
  const PlasticParams prm{zpcm::nacc_bulk_host(model.E, model.nu), model.xi, model.beta, zpcm::nacc_msqr_host(model.fa, model.dim),
                          model.hardeningOn};
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_plastic_kernel<3><<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt,
                                                                    model.volume, mu, lam, prm);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_g2p2g_apic(zpc_particles_view P, zpc_hashtable_view tb, float dx, float dt, int model_kind, const void *model,
                       const float *gridv, float *gridr, zpc_stream_t stream) {
  if (!model || (unsigned)model_kind > 4u || !tb.keys || !tb.indices) return ZPCB200_E_BADARG;
  if (!P.count) return ZPCB200_OK;
  if (!P.X || !gridv || !gridr || (model_kind == ZPC_MODEL_EOS ? !P.J : !P.F) ||
      ((model_kind == ZPC_MODEL_DRUCKER_PRAGER || model_kind == ZPC_MODEL_NACC) && !P.logJp))
    return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
  const zpcp::LegacyGrid ga{tb};
  float mu = 0.f, lam = 0.f;
  PlasticParams prm{0.f, 0.f, 0.f, 0.f, 0};
  switch (model_kind) {
    case ZPC_MODEL_FIXED_COROTATED: {
      const auto &m = *(const zpc_fixed_corotated *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      g2p2g_aos_kernel<0><<<grid, 128, 0, s>>>(P, ga, gridv, gridr, dx, dt, m.volume, mu, lam, prm);
    } break;
    case ZPC_MODEL_VONMISES: {
      const auto &m = *(const zpc_vonmises_fixed_corotated *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      prm.a = m.yieldStress;
      g2p2g_aos_kernel<1><<<grid, 128, 0, s>>>(P, ga, gridv, gridr, dx, dt, m.volume, mu, lam, prm);
    } break;
    case ZPC_MODEL_DRUCKER_PRAGER: {
      const auto &m = *(const zpc_drucker_prager *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      prm = PlasticParams{m.cohesion, m.beta, m.yieldSurface, 0.f, m.volumeCorrection};
      g2p2g_aos_kernel<2><<<grid, 128, 0, s>>>(P, ga, gridv, gridr, dx, dt, m.volume, mu, lam, prm);
    } break;
    case ZPC_MODEL_NACC: {
      const auto &m = *(const zpc_nacc *)model;
      if (m.dim != 3) return ZPCB200_E_BADARG;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      prm = PlasticParams{zpcm::nacc_bulk_host(m.E, m.nu), m.xi, m.beta, zpcm::nacc_msqr_host(m.fa, m.dim), m.hardeningOn};
      g2p2g_aos_kernel<3><<<grid, 128, 0, s>>>(P, ga, gridv, gridr, dx, dt, m.volume, mu, lam, prm);
    } break;
    default: {
      const auto &m = *(const zpc_equation_of_state *)model;
      prm.a = m.bulk;
      prm.b = m.viscosity;
      g2p2g_aos_kernel<4><<<grid, 128, 0, s>>>(P, ga, gridv, gridr, dx, dt, m.volume, mu, lam, prm);
    } break;
  }
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_g2p_apic(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels < 4 || !P.X || !P.V || !P.C || !P.F) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  g2p_aos_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_eos(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_equation_of_state model,
                         zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels != 7 || !P.X || !P.V || !P.M || !P.C || !P.J) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_eos_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt, model.volume, model.bulk,
                                                             model.viscosity);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_g2p_apic_eos(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels < 4 || !P.X || !P.V || !P.C || !P.J) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  g2p_aos_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::LegacyGrid{tb}, g.tiles, g.numChannels, g.dx, dt);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

static int halo_launch(int mode, zpc_grids_view g, const int *ids, int n, int chn0, int nch, float *buf, zpc_stream_t stream) {
  if (n <= 0) return ZPCB200_OK;
  if (!g.tiles || !ids || !buf || chn0 < 0 || chn0 + nch > g.numChannels) return ZPCB200_E_BADARG;
  const size_t work = (size_t)n * nch * 64;
  int grid = (int)((work + 255) / 256);
  if (grid > ZPC_SM_COUNT * 8) grid = ZPC_SM_COUNT * 8;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) halo_kernel<0><<<grid, 256, 0, s>>>(g.tiles, g.numChannels, ids, n, chn0, nch, buf);
  else if (mode == 1) halo_kernel<1><<<grid, 256, 0, s>>>(g.tiles, g.numChannels, ids, n, chn0, nch, buf);
  else halo_kernel<2><<<grid, 256, 0, s>>>(g.tiles, g.numChannels, ids, n, chn0, nch, buf);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
int zpcb200_halo_pack(zpc_grids_view g, const int *ids, int n, int chn0, int nch, float *buf, zpc_stream_t st) {
  return halo_launch(0, g, ids, n, chn0, nch, buf, st);
}
int zpcb200_halo_unpack_add(zpc_grids_view g, const int *ids, int n, int chn0, int nch, const float *buf, zpc_stream_t st) {
  return halo_launch(1, g, ids, n, chn0, nch, (float *)buf, st);
}
int zpcb200_halo_unpack_set(zpc_grids_view g, const int *ids, int n, int chn0, int nch, const float *buf, zpc_stream_t st) {
  return halo_launch(2, g, ids, n, chn0, nch, (float *)buf, st);
}
}

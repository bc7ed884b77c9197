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

namespace {

// ---- partition build --------------------------------------------------------------------------------
constexpr unsigned CODE_EMPTY = 0xffffffffu;
constexpr int CODE_BIAS = 512;  // block coordinates in [-512, 511] per axis (|cell coord| < 2048)
__device__ __forceinline__ bool code_pack(int bx, int by, int bz, unsigned &code) {
  const unsigned ux = (unsigned)(bx + CODE_BIAS), uy = (unsigned)(by + CODE_BIAS), uz = (unsigned)(bz + CODE_BIAS);
  code = (ux << 20) | (uy << 10) | uz;
  return (ux | uy | uz) < 1024u;
}
__device__ __forceinline__ void code_unpack(unsigned code, int &bx, int &by, int &bz) {
  bx = (int)(code >> 20) - CODE_BIAS;
  by = (int)((code >> 10) & 1023u) - CODE_BIAS;
  bz = (int)(code & 1023u) - CODE_BIAS;
}
__device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// insert code into the scratch set; the first inserter appends it to list
__device__ __forceinline__ void set_insert(unsigned code, unsigned *set, unsigned set_mask, unsigned *list,
                                           int list_cap, int *list_cnt, int *overflow) {
  unsigned slot = mix32(code) & set_mask;
  for (unsigned probes = 0; probes <= set_mask; ++probes) {
    unsigned cur = set[slot];
    if (cur == code) return;
    if (cur == CODE_EMPTY) {
      cur = atomicCAS(&set[slot], CODE_EMPTY, code);
      if (cur == CODE_EMPTY) {
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

__global__ void part_clear_kernel(int table_size, int *keys, int *indices, int *status, unsigned *set, unsigned set_n,
                                  unsigned *list, int list_cap, int *counters) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < (size_t)table_size * 3; i += stride) keys[i] = INT_MAX;
  for (size_t i = t0; i < (size_t)table_size; i += stride) { indices[i] = -1; status[i] = -1; }
  for (size_t i = t0; i < set_n; i += stride) set[i] = CODE_EMPTY;
  for (size_t i = t0; i < (size_t)list_cap; i += stride) list[i] = CODE_EMPTY;  // pads sort to the end
  if (t0 < 4) counters[t0] = 0;
}

__global__ void part_mark_kernel(PortAcc<const float> x, size_t n, float dxinv, unsigned *set, unsigned set_mask,
                                 unsigned *list, int list_cap, int *list_cnt, int *overflow) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t first = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
  for (size_t i0 = first; i0 < n; i0 += stride) {  // warp-uniform trip count
    const size_t i = i0 + (threadIdx.x & 31);
    unsigned code = CODE_EMPTY;
    if (i < n) {
      const int bx = zpcm::floor_div4(zpcm::sparsity_coord(x.at(i, 0), dxinv));
      const int by = zpcm::floor_div4(zpcm::sparsity_coord(x.at(i, 1), dxinv));
      const int bz = zpcm::floor_div4(zpcm::sparsity_coord(x.at(i, 2), dxinv));
      if (!code_pack(bx, by, bz, code)) { code = CODE_EMPTY; if (overflow) *overflow = 1; }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, code);
    if (code != CODE_EMPTY && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
      set_insert(code, set, set_mask, list, list_cap, list_cnt, overflow);
  }
}
// EnlargeSparsity{lo, hi} (SparsityOp.hpp:88-112): every block present after the particle pass adds its
// neighbours at offsets [lo, hi)^3; the reference's substep uses {0, 2}
__global__ void part_enlarge_kernel(unsigned *set, unsigned set_mask, unsigned *list, int list_cap,
                                    const int *cnt_before, int *list_cnt, int lo, int hi, int *overflow) {
  const int n0 = min(*cnt_before, list_cap);
  const int w = hi - lo, w3 = w * w * w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n0 * w3; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / w3), o = (int)(t % w3);
    const int ox = lo + o / (w * w), oy = lo + (o / w) % w, oz = lo + o % w;
    if ((ox | oy | oz) == 0) continue;
    int bx, by, bz;
    code_unpack(list[i], bx, by, bz);
    unsigned code;
    if (code_pack(bx + ox, by + oy, bz + oz, code))
      set_insert(code, set, set_mask, list, list_cap, list_cnt, overflow);
    else if (overflow) *overflow = 1;
  }
}
__global__ void part_snapshot_kernel(const int *src, int *dst) { *dst = *src; }

// place key i (rank order) at the first free slot of ITS OWN probe sequence (any such placement is a valid table)
__global__ void part_place_kernel(const unsigned *sorted, const int *list_cnt, int list_cap, int table_size, int *keys,
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

// ---- grid -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clean_grid_kernel(float4 *tiles, const int *cnt, int nch, size_t cap_blocks) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const size_t n4 = nb * (size_t)nch * 16;  // float4 per tile = nch*64/4
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) tiles[i] = z;
}

// one thread per (block, cell); 4 blocks per 256-thread CTA
__global__ void __launch_bounds__(256) grid_update_kernel(float *tiles, const int *cnt, int nch, size_t cap_blocks, float dt,
                                                          float ex, float ey, float ez, int mode, float *max_vel_sqr) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const int cell = threadIdx.x & 63;
  float mx = 0.f;
  for (size_t b = (size_t)blockIdx.x * 4 + (threadIdx.x >> 6); b < nb; b += (size_t)gridDim.x * 4) {
    float *t = tiles + b * (size_t)nch * 64;
    float mass = t[cell];
    if (mass != 0.f) {
      float mvx = t[64 + cell], mvy = t[128 + cell], mvz = t[192 + cell];
      if (mode == 1) { mvx += t[256 + cell]; mvy += t[320 + cell]; mvz += t[384 + cell]; }
      mass = 1.f / mass;
      const float vx = mvx * mass + ex * dt, vy = mvy * mass + ey * dt, vz = mvz * mass + ez * dt;
      t[64 + cell] = vx; t[128 + cell] = vy; t[192 + cell] = vz;
      mx = fmaxf(mx, vx * vx + vy * vy + vz * vz);
    } else if (mode == 1) {
      // explicit mode folds rhs into mv for every cell (oracle: mv += rhs before the mass test)
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
    if (mx > 0.f) atomicMax((int *)max_vel_sqr, __float_as_int(mx));  // non-negative floats order as ints
  }
}

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
    const float d0 = px - col.origin[0], d1 = py - col.origin[1], d2 = pz - col.origin[2];
    float n0, n1, n2, dist;
    if (col.geometry == ZPC_GEOM_PLANE) {
      n0 = col.normal[0]; n1 = col.normal[1]; n2 = col.normal[2];
      dist = __fadd_rn(__fadd_rn(__fmul_rn(n0, d0), __fmul_rn(n1, d1)), __fmul_rn(n2, d2));  // no contraction: the sign decides
    } else {
      const float l2 = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
      const float len = sqrtf(l2);
      dist = len - col.normal[0];
      const bool tiny = l2 < 1e-7f;
      n0 = tiny ? 0.f : d0 / len; n1 = tiny ? 0.f : d1 / len; n2 = tiny ? 0.f : d2 / len;
    }
    if (dist < 0.f) {
      float vx = t[64 + cell], vy = t[128 + cell], vz = t[192 + cell];
      if (col.type == ZPC_COLLIDER_STICKY) {
        vx = vy = vz = 0.f;
      } else {
        const float proj = n0 * vx + n1 * vy + n2 * vz;
        if (col.type == ZPC_COLLIDER_SLIP || proj < 0.f) { vx -= proj * n0; vy -= proj * n1; vz -= proj * n2; }
      }
      t[64 + cell] = vx; t[128 + cell] = vy; t[192 + cell] = vz;
    }
  }
}

// ---- P2G / G2P on AoS particles, any order -----------------------------------------------------------------
__global__ void __launch_bounds__(128) p2g_aos_kernel(zpc_particles_view P, zpc_hashtable_view tb, float *tiles, int nch,
                                                      float dx, float dt, float volume, float mu, float lam) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  float pos[3], vel[3], C[9], F[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pos[d] = P.X[3 * p + d]; vel[d] = P.V[3 * p + d]; }
#pragma unroll
  for (int d = 0; d < 9; ++d) { C[d] = P.C[9 * p + d]; F[d] = P.F[9 * p + d]; }
  zpcp::p2g_scatter_particle(pos, vel, P.M[p], C, F, tb, tiles, nch, dx, dt, volume, mu, lam);
}

__global__ void __launch_bounds__(128) p2g_aos_eos_kernel(zpc_particles_view P, zpc_hashtable_view tb, float *tiles, int nch,
                                                          float dx, float dt, float volume, float bulk, float viscosity) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  float pos[3], vel[3], C[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pos[d] = P.X[3 * p + d]; vel[d] = P.V[3 * p + d]; }
#pragma unroll
  for (int d = 0; d < 9; ++d) C[d] = P.C[9 * p + d];
  zpcp::p2g_scatter_particle_eos(pos, vel, P.M[p], C, P.J[p], tb, tiles, nch, dx, dt, volume, bulk, viscosity);
}

template <bool EOS>
__global__ void __launch_bounds__(128) g2p_aos_kernel(zpc_particles_view P, zpc_hashtable_view tb, const float *tiles,
                                                      int nch, float dx, float dt) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float pos[3], vel[3] = {0.f, 0.f, 0.f}, C[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] = P.X[3 * p + d];
#pragma unroll
  for (int d = 0; d < 9; ++d) C[d] = 0.f;
  zpcm::Arena ar;
  zpcm::arena_init(ar, dx, pos);
  long long boff[8];
  zpcp::resolve_blocks(ar.corner, tb, nch, boff);
  const int lx0 = ar.corner[0] & 3, ly0 = ar.corner[1] & 3, lz0 = ar.corner[2] & 3;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long off = boff[((lx >> 2) << 2) | ((ly >> 2) << 1) | (lz >> 2)];
        if (off < 0) continue;
        const float *t = tiles + off + (((lx & 3) << 4) | ((ly & 3) << 2) | (lz & 3));
        const float xixp[3] = {(float)i * dx - ar.local[0], (float)j * dx - ar.local[1], (float)k * dx - ar.local[2]};
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
        const float vi[3] = {__ldg(t + 64), __ldg(t + 128), __ldg(t + 192)};
#pragma unroll
        for (int d = 0; d < 3; ++d) vel[d] += vi[d] * W;
#pragma unroll
        for (int d = 0; d < 9; ++d) C[d] += W * vi[d % 3] * xixp[d / 3] * D_inv;
      }
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] += vel[d] * dt;
  if constexpr (EOS) {  // G2P.hpp:69-73
    P.J[p] = (1 + (C[0] + C[4] + C[8]) * dt) * P.J[p];
  } else {
    float Fo[9], tmp[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) { Fo[d] = P.F[9 * p + d]; tmp[d] = C[d] * dt + ((d & 3) ? 0.f : 1.f); }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) P.F[9 * p + 3 * c + r] = tmp[r] * Fo[3 * c] + tmp[3 + r] * Fo[3 * c + 1] + tmp[6 + r] * Fo[3 * c + 2];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { P.X[3 * p + d] = pos[d]; P.V[3 * p + d] = vel[d]; }
#pragma unroll
  for (int d = 0; d < 9; ++d) P.C[9 * p + d] = C[d];
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

int zpcb200_partition_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, float dx, zpc_hashtable_view tb,
                            int enlarge_lo, int enlarge_hi, int *overflow, zpc_stream_t stream) {
  if (!temp_bytes || tb.tableSize <= 0 || enlarge_hi < enlarge_lo || enlarge_hi - enlarge_lo > 8) return ZPCB200_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  // scratch set: power of two >= tableSize/2 ; list capacity = tableSize/8 block codes
  size_t set_n = 1;
  while (set_n < (size_t)tb.tableSize / 2) set_n <<= 1;
  if (set_n < 1024) set_n = 1024;
  const int list_cap = tb.tableSize / 8 > 64 ? tb.tableSize / 8 : 64;
  size_t sort_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = zpcb200_radix_sort_u32(nullptr, &sort_bytes, none, none, (size_t)list_cap, 0, 30, nullptr);
  if (rc) return rc;
  const size_t off_set = 256;
  const size_t off_list = zpc_align_up(off_set + 4 * set_n, 256);
  const size_t off_sorted = zpc_align_up(off_list + 4 * (size_t)list_cap, 256);
  const size_t off_sort = zpc_align_up(off_sorted + 4 * (size_t)list_cap, 256);
  const size_t need = off_sort + sort_bytes;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  char *t = (char *)temp;
  int *counters = (int *)t;  // [0] list count, [1] count before enlarge
  unsigned *set = (unsigned *)(t + off_set), *list = (unsigned *)(t + off_list), *sorted = (unsigned *)(t + off_sorted);
  const int G = ZPC_SM_COUNT * 8;
  part_clear_kernel<<<G, 256, 0, s>>>(tb.tableSize, tb.keys, tb.indices, tb.status, set, (unsigned)set_n, list, list_cap,
                                       counters);
  ZPC_CHECK_LAUNCH();
  if (n) {
    part_mark_kernel<<<G, 256, 0, s>>>(PortAcc<const float>(x), n, 1.0f / dx, set, (unsigned)(set_n - 1), list, list_cap,
                                        counters, overflow);
    ZPC_CHECK_LAUNCH();
  }
  part_snapshot_kernel<<<1, 1, 0, s>>>(counters, counters + 1);
  ZPC_CHECK_LAUNCH();
  if (enlarge_hi - enlarge_lo > 0) {
    part_enlarge_kernel<<<G, 256, 0, s>>>(set, (unsigned)(set_n - 1), list, list_cap, counters + 1, counters, enlarge_lo,
                                           enlarge_hi, overflow);
    ZPC_CHECK_LAUNCH();
  }
  zpc_port pl = {list, 0, 0, 0, 1}, ps = {sorted, 0, 0, 0, 1};
  size_t sb = sort_bytes;
  rc = zpcb200_radix_sort_u32(t + off_sort, &sb, pl, ps, (size_t)list_cap, 0, 30, stream);
  if (rc) return rc;
  part_place_kernel<<<G, 256, 0, s>>>(sorted, counters, list_cap, tb.tableSize, tb.keys, tb.indices, tb.activeKeys, tb.cnt,
                                       overflow);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_clean_grid(zpc_grids_view g, const int *cnt, zpc_stream_t stream) {
  if (!g.tiles || !cnt) return ZPCB200_E_BADARG;
  clean_grid_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((float4 *)g.tiles, cnt, g.numChannels, g.numBlocks);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_grid_update(zpc_grids_view g, const int *cnt, float dt, const float extf[3], int mode, float *maxVelSqr,
                        zpc_stream_t stream) {
  if (!g.tiles || !cnt || !maxVelSqr || g.numChannels < (mode == 1 ? 7 : 4)) return ZPCB200_E_BADARG;
  grid_update_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, cnt, g.numChannels, g.numBlocks, dt, extf[0],
                                                                        extf[1], extf[2], mode, maxVelSqr);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_apply_boundary(zpc_grids_view g, zpc_hashtable_view tb, zpc_collider col, zpc_stream_t stream) {
  if (!g.tiles || !tb.activeKeys || !tb.cnt || g.numChannels < 4 || (unsigned)col.geometry > 1u || (unsigned)col.type > 2u)
    return ZPCB200_E_BADARG;
  apply_boundary_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(g.tiles, tb.activeKeys, tb.cnt, g.numChannels, g.numBlocks,
                                                                           g.dx, col);
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
  p2g_aos_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, tb, g.tiles, g.numChannels, g.dx, dt, model.volume, mu, lam);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_g2p_apic(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels < 4 || !P.X || !P.V || !P.C || !P.F) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  g2p_aos_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(P, tb, g.tiles, g.numChannels, g.dx, dt);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_eos(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_equation_of_state model,
                         zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels != 7 || !P.X || !P.V || !P.M || !P.C || !P.J) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_eos_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, tb, g.tiles, g.numChannels, g.dx, dt, model.volume, model.bulk,
                                                             model.viscosity);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_g2p_apic_eos(zpc_particles_view P, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  if (P.count == 0) return ZPCB200_OK;
  if (g.numChannels < 4 || !P.X || !P.V || !P.C || !P.J) return ZPCB200_E_BADARG;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  g2p_aos_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(P, tb, g.tiles, g.numChannels, g.dx, dt);
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

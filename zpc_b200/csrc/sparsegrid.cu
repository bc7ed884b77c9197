// SparseGrid<3,f32,8> variant of the MPM path: bht<i32,3,int,16> partition build readable by the
// reference's unmodified BHTView::query, the transfer / grid functors on side-8 blocks, and the SparseGridView
// accessors (valueOr, iCoord, wCoord) as bulk operations.
//
//   table   : container/Bht.hpp — buckets of 16 slots, three universal hashes drawn from std::mt19937(2) (:165-169),
//             16-byte padded key slots whose empty value is the byte pattern 0x3f (:108-112), keys = block ORIGINS in
//             cell coordinates (geometry/SparseGrid.hpp:305-309).
//   grid    : TileVector<f32,512>, tile b = [nch][512], cell offset (x*8+y)*8+z (SparseGrid.hpp:275-283).
//   kernels : the same per-particle scatter / gather as the Grids<f32,3,4> drop-in path (mpm_kernels.cuh), instantiated
//             for zpcp::SparseGrid8.
#include <climits>

#include "common.cuh"
#include "mpm_kernels.cuh"
#include "mpm_math.cuh"
#include "mpm_particle.cuh"
#include "partition.cuh"

namespace {

constexpr int KEY_EMPTY = 0x3f3f3f3f;

__global__ void sg_clear_table_kernel(zpc_bht_view tb) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int4 *k = reinterpret_cast<int4 *>(tb.keys);
  for (size_t i = t0; i < (size_t)tb.tableSize; i += stride) {
    k[i] = make_int4(KEY_EMPTY, KEY_EMPTY, KEY_EMPTY, KEY_EMPTY);  // Table::reset: every byte of the slot is 0x3f
    tb.indices[i] = -1;
    tb.status[i] = -1;
  }
}

// key i (rank order) goes to the first of its three candidate buckets with room: slots 0..14 of a bucket are claimed
// in order with a CAS on the index array (the reference's insert never uses slot 15: threshold = 14, Bht.hpp:41)
template <class CODE>
__global__ void sg_place_kernel(const CODE *sorted, const int *list_cnt, int list_cap, zpc_bht_view tb, int *overflow) {
  const int n = min(*list_cnt, list_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int bx, by, bz;
    code_unpack(sorted[i], bx, by, bz);
    const int kx = bx * 8, ky = by * 8, kz = bz * 8;
    bool placed = false;
    for (int it = 0; it < 3 && !placed; ++it) {
      const unsigned b = zpcm::bht_hash(tb.hf[2 * it], tb.hf[2 * it + 1], kx, ky, kz) % tb.numBuckets * 16u;
      for (int s = 0; s < 15; ++s) {
        if (tb.indices[b + s] != -1) continue;
        if (atomicCAS(&tb.indices[b + s], -1, i) == -1) {
          reinterpret_cast<int4 *>(tb.keys)[b + s] = make_int4(kx, ky, kz, KEY_EMPTY);
          placed = true;
          break;
        }
      }
    }
    if (!placed) {
      if (overflow) *overflow = 1;
      if (tb.success) *tb.success = 0;
    }
    tb.activeKeys[3 * (size_t)i] = kx; tb.activeKeys[3 * (size_t)i + 1] = ky; tb.activeKeys[3 * (size_t)i + 2] = kz;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *tb.cnt = n;
    if ((unsigned)n + 20u >= tb.tableSize) {  // the reference's "proximity" failure (Bht.hpp:522-527)
      if (overflow) *overflow = 1;
      if (tb.success) *tb.success = 0;
    }
  }
}

__global__ void sg_set_success_kernel(int *success) { *success = 1; }

__global__ void sg_value_or_kernel(zpc_sparsegrid_view sg, int chn, const int *__restrict__ coords, size_t n, float dflt,
                                   float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
  const int cx = x & 7, cy = y & 7, cz = z & 7;  // decomposeCoord (SparseGrid.hpp:305-309)
  const int bno = zpcm::bht_query(x - cx, y - cy, z - cz, sg.table);
  out[i] = bno == -1 ? dflt : sg.grid[((size_t)bno * sg.numChannels + chn) * 512 + ((cx * 8 + cy) * 8 + cz)];
}

__global__ void sg_cell_coords_kernel(zpc_sparsegrid_view sg, const int *__restrict__ bno, const int *__restrict__ cno, size_t n,
                                      int *icoord, float *wcoord) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = bno[i], c = cno[i];
  const int ic[3] = {sg.table.activeKeys[3 * (size_t)b] + ((c >> 6) & 7), sg.table.activeKeys[3 * (size_t)b + 1] + ((c >> 3) & 7),
                     sg.table.activeKeys[3 * (size_t)b + 2] + (c & 7)};
  if (icoord) { icoord[3 * i] = ic[0]; icoord[3 * i + 1] = ic[1]; icoord[3 * i + 2] = ic[2]; }
  if (wcoord) {
    const float *M = sg.transform;
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // (X, 1) * M, accumulated in the reference's order (no contraction)
      float s = __fmul_rn((float)ic[0], M[j]);
      s = __fadd_rn(s, __fmul_rn((float)ic[1], M[4 + j]));
      s = __fadd_rn(s, __fmul_rn((float)ic[2], M[8 + j]));
      wcoord[3 * i + j] = __fadd_rn(s, M[12 + j]);
    }
  }
}

// TileVector::reorderTiles (container/TileVector.hpp:641-691): float4 copies, one tile = tile_f4 float4
__global__ void reorder_tiles_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t tile_f4, const int *__restrict__ map,
                                     size_t ntiles, int scatter) {
  const size_t total = ntiles * tile_f4;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / tile_f4, r = t % tile_f4;
    const size_t j = (size_t)map[i];
    if (scatter) dst[j * tile_f4 + r] = src[t];
    else dst[t] = src[j * tile_f4 + r];
  }
}

// bht::reorder (container/Bht.hpp:343-400, ReorderBht): new key list + renumbered index of every key's slot
__global__ void bht_reorder_kernel(zpc_bht_view tb, const int *__restrict__ map, int n, int scatter, int *ordered_keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = map[i];
  const int src = scatter ? i : j, dst = scatter ? j : i;
  const int kx = tb.activeKeys[3 * (size_t)src], ky = tb.activeKeys[3 * (size_t)src + 1], kz = tb.activeKeys[3 * (size_t)src + 2];
  ordered_keys[3 * (size_t)dst] = kx; ordered_keys[3 * (size_t)dst + 1] = ky; ordered_keys[3 * (size_t)dst + 2] = kz;
  bool found = false;
  for (int it = 0; it < 3 && !found; ++it) {
    const unsigned b = zpcm::bht_hash(tb.hf[2 * it], tb.hf[2 * it + 1], kx, ky, kz) % tb.numBuckets * 16u;
    const int4 *k = reinterpret_cast<const int4 *>(tb.keys) + b;
    for (int s2 = 0; s2 < 16; ++s2) {
      const int4 c = k[s2];
      if (c.x == kx && c.y == ky && c.z == kz) { tb.indices[b + s2] = dst; found = true; break; }
    }
  }
  if (!found && tb.success) *tb.success = 0;
}

// 30-bit Morton code of a block (10 bits per axis of key/8 + 512) and its current index
__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void sg_morton_kernel(const int *__restrict__ active_keys, int n, unsigned *codes, int *ids, int *overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bx = (active_keys[3 * (size_t)i] >> 3) + 512, by = (active_keys[3 * (size_t)i + 1] >> 3) + 512,
            bz = (active_keys[3 * (size_t)i + 2] >> 3) + 512;
  if (((unsigned)bx | (unsigned)by | (unsigned)bz) >= 1024u && overflow) *overflow = 1;
  codes[i] = (spread10((unsigned)bx) << 2) | (spread10((unsigned)by) << 1) | spread10((unsigned)bz);
  ids[i] = i;
}

// dx of an axis-aligned uniform transform without translation; false otherwise
bool sg_uniform_dx(const zpc_sparsegrid_view &sg, float &dx) {
  const float *M = sg.transform;
  dx = M[0];
  if (!(dx > 0.f) || M[5] != dx || M[10] != dx) return false;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      if (i == j) continue;
      if (M[4 * i + j] != 0.f) return false;
    }
  return M[15] == 1.f;
}
bool sg_table_ok(const zpc_bht_view &t) {
  return t.keys && t.indices && t.status && t.activeKeys && t.cnt && t.numBuckets * 16u == t.tableSize;
}

}  // namespace

extern "C" {

// std::mt19937(2): MT19937 with the standard constants; universal_hash(rng): hashx = rng() % prime (>= 1), hashy = rng() % prime
void zpcb200_bht_params(uint32_t hf[6]) {
  uint32_t mt[624];
  mt[0] = 2u;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  for (int i = 0; i < 624; ++i) {
    const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
    mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  for (int k = 0; k < 6; ++k) {
    uint32_t y = mt[k];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    y %= 4294967291u;
    if ((k & 1) == 0 && y < 1) y = 1;
    hf[k] = y;
  }
}
size_t zpcb200_bht_table_size(size_t expected) {
  if (expected == 0) return 0;
  size_t p = 1;
  while (p < expected) p <<= 1;
  const size_t n = p * 2;
  return n + (16 - n % 16);
}

}  // extern "C"
template <class CODE>
static int sg_partition_build_impl(void *temp, size_t *temp_bytes, zpc_port x, size_t n, zpc_sparsegrid_view sg, int enlarge_lo,
                                   int enlarge_hi, int *overflow, zpc_stream_t stream) {
  if (!temp_bytes || sg.table.tableSize < 16 || enlarge_hi < enlarge_lo || enlarge_hi - enlarge_lo > 8) return ZPCB200_E_BADARG;
  float dx;
  if (!sg_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  PartScratch L;
  int rc = part_scratch_layout<CODE>((size_t)sg.table.tableSize * 4, L);  // the bht holds up to tableSize/2 keys: list of tableSize/2 codes
  if (rc) return rc;
  if (!temp) { *temp_bytes = L.need; return ZPCB200_OK; }
  if (*temp_bytes < L.need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (!sg_table_ok(sg.table)) return ZPCB200_E_BADARG;
  char *t = (char *)temp;
  const int G = ZPC_SM_COUNT * 8;
  sg_clear_table_kernel<<<G, 256, 0, s>>>(sg.table);
  ZPC_CHECK_LAUNCH();
  if (sg.table.success) {
    sg_set_success_kernel<<<1, 1, 0, s>>>(sg.table.success);
    ZPC_CHECK_LAUNCH();
  }
  rc = part_collect_sorted<3, CODE>(t, L, x, n, dx, enlarge_lo, enlarge_hi, overflow, s);
  if (rc) return rc;
  sg_place_kernel<CODE><<<G, 256, 0, s>>>((const CODE *)(t + L.off_sorted), (const int *)t, L.list_cap, sg.table, overflow);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
extern "C" {
int zpcb200_sg_partition_build(void *temp, size_t *temp_bytes, zpc_port x, size_t n, zpc_sparsegrid_view sg, int enlarge_lo,
                               int enlarge_hi, int *overflow, zpc_stream_t stream) {
  return sg_partition_build_impl<unsigned>(temp, temp_bytes, x, n, sg, enlarge_lo, enlarge_hi, overflow, stream);
}
/* 64-bit block codes: block coordinates in [-2^20, 2^20) per axis */
int zpcb200_sg_partition_build_wide(void *temp, size_t *temp_bytes, zpc_port x, size_t n, zpc_sparsegrid_view sg, int enlarge_lo,
                                    int enlarge_hi, int *overflow, zpc_stream_t stream) {
  return sg_partition_build_impl<unsigned long long>(temp, temp_bytes, x, n, sg, enlarge_lo, enlarge_hi, overflow, stream);
}

int zpcb200_tilevector_reorder_tiles(const float *src, float *dst, int numChannels, int tileLength, const int *map, size_t numTiles,
                                     int scatter, zpc_stream_t stream) {
  if (numChannels <= 0 || tileLength <= 0 || (tileLength & 3) || (numTiles && (!src || !dst || !map || src == dst))) return ZPCB200_E_BADARG;
  if (!numTiles) return ZPCB200_OK;
  reorder_tiles_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((const float4 *)src, (float4 *)dst,
                                                                          (size_t)numChannels * tileLength / 4, map, numTiles, scatter);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_bht_reorder(zpc_bht_view table, const int *map, int n, int scatter, int *orderedKeys, zpc_stream_t stream) {
  if (n < 0 || (n && (!map || !orderedKeys || !sg_table_ok(table) || orderedKeys == table.activeKeys))) return ZPCB200_E_BADARG;
  if (!n) return ZPCB200_OK;
  bht_reorder_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, map, n, scatter, orderedKeys);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_morton_order(void *temp, size_t *temp_bytes, zpc_sparsegrid_view sg, int n, int *map, int *overflow, zpc_stream_t stream) {
  if (!temp_bytes || n < 0) return ZPCB200_E_BADARG;
  size_t sort_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = zpcb200_radix_sort_pair_u32(nullptr, &sort_bytes, none, none, none, none, (size_t)n, 0, 30, nullptr);
  if (rc) return rc;
  const size_t o_codes = 0, o_ids = zpc_align_up(4 * (size_t)n, 256), o_codes2 = o_ids + zpc_align_up(4 * (size_t)n, 256),
               o_sort = o_codes2 + zpc_align_up(4 * (size_t)n, 256), need = o_sort + sort_bytes;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (!n) return ZPCB200_OK;
  if (!map || !sg.table.activeKeys) return ZPCB200_E_BADARG;
  char *t = (char *)temp;
  unsigned *codes = (unsigned *)(t + o_codes), *codes2 = (unsigned *)(t + o_codes2);
  int *ids = (int *)(t + o_ids);
  sg_morton_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sg.table.activeKeys, n, codes, ids, overflow);
  ZPC_CHECK_LAUNCH();
  zpc_port pc = {codes, 0, 0, 0, 1}, pi = {ids, 0, 0, 0, 1}, pc2 = {codes2, 0, 0, 0, 1}, pm = {map, 0, 0, 0, 1};
  size_t sb = sort_bytes;
  return zpcb200_radix_sort_pair_u32(t + o_sort, &sb, pc, pi, pc2, pm, (size_t)n, 0, 30, stream);
}

int zpcb200_sg_clean(zpc_sparsegrid_view sg, zpc_stream_t stream) {
  if (!sg.grid || !sg.table.cnt) return ZPCB200_E_BADARG;
  clean_grid_kernel<<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((float4 *)sg.grid, sg.table.cnt, sg.numChannels, sg.numBlocks, 512);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_p2g_apic_fcr(zpc_particles_view P, zpc_sparsegrid_view sg, float dt, zpc_fixed_corotated model, zpc_stream_t stream) {
  if (sg.numChannels != 7 || !sg.grid || !sg_table_ok(sg.table)) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.M || !P.C || !P.F)) return ZPCB200_E_BADARG;
  float dx;
  if (!sg_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (!P.count) return ZPCB200_OK;
  float mu, lam;
  zpcm::lame_host(model.E, model.nu, mu, lam);
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  p2g_aos_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::SparseGrid8{sg.table}, sg.grid, sg.numChannels, dx, dt, model.volume, mu, lam);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

// the other constitutive models of P2GTransfer on the SparseGrid: the any-order kernels of mpm_kernels.cuh instantiated for SparseGrid8
int zpcb200_sg_p2g_apic_model(zpc_particles_view P, zpc_sparsegrid_view sg, float dt, int model_kind, const void *model, zpc_stream_t stream) {
  if (!model || (unsigned)model_kind > 4u || sg.numChannels != 7 || !sg.grid || !sg_table_ok(sg.table)) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.M || !P.C || (model_kind == ZPC_MODEL_EOS ? !P.J : !P.F) ||
                  ((model_kind == ZPC_MODEL_DRUCKER_PRAGER || model_kind == ZPC_MODEL_NACC) && !P.logJp)))
    return ZPCB200_E_BADARG;
  float dx;
  if (!sg_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (!P.count) return ZPCB200_OK;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  cudaStream_t s = (cudaStream_t)stream;
  const zpcp::SparseGrid8 ga{sg.table};
  float mu = 0.f, lam = 0.f;
  switch (model_kind) {
    case ZPC_MODEL_FIXED_COROTATED: {
      const auto &m = *(const zpc_fixed_corotated *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      p2g_aos_kernel<<<grid, 128, 0, s>>>(P, ga, sg.grid, sg.numChannels, dx, dt, m.volume, mu, lam);
    } break;
    case ZPC_MODEL_VONMISES: {
      const auto &m = *(const zpc_vonmises_fixed_corotated *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      p2g_aos_vm_kernel<<<grid, 128, 0, s>>>(P, ga, sg.grid, sg.numChannels, dx, dt, m.volume, mu, lam, m.yieldStress);
    } break;
    case ZPC_MODEL_DRUCKER_PRAGER: {
      const auto &m = *(const zpc_drucker_prager *)model;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      p2g_aos_plastic_kernel<2><<<grid, 128, 0, s>>>(P, ga, sg.grid, sg.numChannels, dx, dt, m.volume, mu, lam,
                                                     PlasticParams{m.cohesion, m.beta, m.yieldSurface, 0.f, m.volumeCorrection});
    } break;
    case ZPC_MODEL_NACC: {
      const auto &m = *(const zpc_nacc *)model;
      if (m.dim != 3) return ZPCB200_E_BADARG;
      zpcm::lame_host(m.E, m.nu, mu, lam);
      p2g_aos_plastic_kernel<3><<<grid, 128, 0, s>>>(P, ga, sg.grid, sg.numChannels, dx, dt, m.volume, mu, lam,
                                                     PlasticParams{zpcm::nacc_bulk_host(m.E, m.nu), m.xi, m.beta, zpcm::nacc_msqr_host(m.fa, m.dim), m.hardeningOn});
    } break;
    default: {
      const auto &m = *(const zpc_equation_of_state *)model;
      p2g_aos_eos_kernel<<<grid, 128, 0, s>>>(P, ga, sg.grid, sg.numChannels, dx, dt, m.volume, m.bulk, m.viscosity);
    } break;
  }
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
int zpcb200_sg_g2p_apic_eos(zpc_particles_view P, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream) {
  if (sg.numChannels < 4 || !sg.grid || !sg_table_ok(sg.table)) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.C || !P.J)) return ZPCB200_E_BADARG;
  float dx;
  if (!sg_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (!P.count) return ZPCB200_OK;
  g2p_aos_kernel<true><<<(unsigned)((P.count + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P, zpcp::SparseGrid8{sg.table}, sg.grid, sg.numChannels, dx, dt);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_grid_update(zpc_sparsegrid_view sg, float dt, const float extf[3], int mode, float *maxVelSqr, zpc_stream_t stream) {
  if (!sg.grid || !sg.table.cnt || !extf || !maxVelSqr || (mode != 0 && mode != 1) || sg.numChannels < (mode ? 7 : 4)) return ZPCB200_E_BADARG;
  grid_update_kernel<512><<<ZPC_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(sg.grid, sg.table.cnt, sg.numChannels, sg.numBlocks, dt, extf[0],
                                                                              extf[1], extf[2], mode, maxVelSqr);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_g2p_apic(zpc_particles_view P, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream) {
  if (sg.numChannels < 4 || !sg.grid || !sg_table_ok(sg.table)) return ZPCB200_E_BADARG;
  if (P.count && (!P.X || !P.V || !P.C || !P.F)) return ZPCB200_E_BADARG;
  float dx;
  if (!sg_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (!P.count) return ZPCB200_OK;
  const unsigned grid = (unsigned)((P.count + 127) / 128);
  g2p_aos_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(P, zpcp::SparseGrid8{sg.table}, sg.grid, sg.numChannels, dx, dt);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_value_or(zpc_sparsegrid_view sg, int chn, const int *coords, size_t n, float dflt, float *out, zpc_stream_t stream) {
  if (!sg.grid || !sg_table_ok(sg.table) || chn < 0 || chn >= sg.numChannels || (n && (!coords || !out))) return ZPCB200_E_BADARG;
  if (!n) return ZPCB200_OK;
  sg_value_or_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sg, chn, coords, n, dflt, out);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_sg_cell_coords(zpc_sparsegrid_view sg, const int *bno, const int *cno, size_t n, int *icoord, float *wcoord,
                           zpc_stream_t stream) {
  if (!sg.table.activeKeys || (n && (!bno || !cno))) return ZPCB200_E_BADARG;
  if (!n) return ZPCB200_OK;
  sg_cell_coords_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sg, bno, cno, n, icoord, wcoord);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
}

/* TEST INFRASTRUCTURE — NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of the reference's bucketed hash table bht<i32,3,int,16> (container/Bht.hpp) and of the
 * SparseGrid<3,f32,8> accessors (geometry/SparseGrid.hpp) that the side-8 variant of the MPM path uses.
 * Pinned against the reference's own containers by tests/test_oracle_vs_ref.py (oracle/_ref/libzpcref.so).
 * Paths are relative to /root/reference/include/zensim/. */
#include <math.h>
#include <string.h>

#include "oracle.h"

/* std::mt19937 (the generator the reference seeds with 2, Bht.hpp:165): MT19937, 32-bit, standard constants */
typedef struct { uint32_t mt[624]; int idx; } zo_mt;
static void mt_seed(zo_mt *g, uint32_t s) {
  g->mt[0] = s;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(zo_mt *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}

#define ZO_PRIME 4294967291u /* universal_hash_base::prime_divisor, py_interop/HashUtils.hpp:12 */

/* bht constructor, Bht.hpp:165-169: three universal_hash(rng) in a row; universal_hash(std::mt19937&),
 * container/Bcht.hpp:39-43: hashx = rng() % prime (forced >= 1), hashy = rng() % prime */
void zo_bht_params(uint32_t hf[6]) {
  zo_mt g;
  mt_seed(&g, 2u);
  for (int k = 0; k < 3; ++k) {
    uint32_t hx = mt_next(&g) % ZO_PRIME;
    if (hx < 1) hx = 1;
    hf[2 * k] = hx;
    hf[2 * k + 1] = mt_next(&g) % ZO_PRIME;
  }
}

/* evaluateTableSize, Bht.hpp:154-158: 2 * next_2pow(n), then ALWAYS one more (partial) bucket: n + (16 - n % 16) */
int zo_bht_table_size(int expected) {
  if (expected == 0) return 0;
  const size_t n = (size_t)zo_next_2pow(expected) * 2;
  return (int)(n + (16 - n % 16));
}

/* universal_hash_base::operator()(vec key), HashUtils.hpp:22-43: sub(k) = ((hashx ^ (u32)k) + hashy) % prime in
 * 32-bit unsigned arithmetic, combined with the 32-bit hash_combine (math/Hash.hpp: seed ^= v + 0x9e3779b9 +
 * (seed << 6) + (seed >> 2)) */
static uint32_t sub_hash(uint32_t hx, uint32_t hy, int k) { return ((hx ^ (uint32_t)k) + hy) % ZO_PRIME; }
uint32_t zo_bht_hash(uint32_t hx, uint32_t hy, const int key[3]) {
  uint32_t h = sub_hash(hx, hy, key[0]);
  for (int d = 1; d < 3; ++d) h ^= sub_hash(hx, hy, key[d]) + 0x9e3779b9u + (h << 6) + (h >> 2);
  return h;
}

/* bht::Table::reset(false), Bht.hpp:108-112: keys byte-filled with 0x3f (all 16 bytes of a slot), status -1; the
 * indices are left alone by the reference — cleared to -1 here so that dumps are comparable */
void zo_bht_clear(int table_size, int *keys16, int *indices, int *status, int *cnt) {
  memset(keys16, 0x3f, (size_t)table_size * 16);
  for (int i = 0; i < table_size; ++i) { indices[i] = -1; status[i] = -1; }
  *cnt = 0;
}

/* BHTView::insert (host), Bht.hpp:609-664, executed serially: returns the new index, -1 when the key is present,
 * INT_MIN (failure_token_v) when the three candidate buckets are full (load > threshold = 14, Bht.hpp:41) */
int zo_bht_insert(const int key[3], int table_size, const uint32_t hf[6], int *keys16, int *indices, int *active_keys,
                  int *cnt) {
  const int nb = table_size / 16;
  if (nb == 0) return (int)0x80000000;
  for (int iter = 0; iter < 3; ++iter) {
    const int b = (int)(zo_bht_hash(hf[2 * iter], hf[2 * iter + 1], key) % (uint32_t)nb) * 16;
    int load = 0;
    for (; load != 16; ++load) {
      const int *k = keys16 + 4 * (size_t)(b + load);
      if (k[0] == key[0] && k[1] == key[1] && k[2] == key[2]) return -1;
      if (k[0] == 0x3f3f3f3f && k[1] == 0x3f3f3f3f && k[2] == 0x3f3f3f3f) break;
    }
    if (load <= 14) {
      int *k = keys16 + 4 * (size_t)(b + load);
      k[0] = key[0]; k[1] = key[1]; k[2] = key[2];
      const int no = (*cnt)++;
      indices[b + load] = no;
      active_keys[3 * no] = key[0]; active_keys[3 * no + 1] = key[1]; active_keys[3 * no + 2] = key[2];
      return no;
    }
  }
  return (int)0x80000000;
}

/* BHTView::query, Bht.hpp:666-700: all 16 slots of the hf0 bucket, then hf1's, then hf2's; -1 when absent */
int zo_bht_query(const int key[3], int table_size, const uint32_t hf[6], const int *keys16, const int *indices) {
  const int nb = table_size / 16;
  if (nb == 0) return -1;
  for (int iter = 0; iter < 3; ++iter) {
    const int b = (int)(zo_bht_hash(hf[2 * iter], hf[2 * iter + 1], key) % (uint32_t)nb) * 16;
    for (int loc = 0; loc != 16; ++loc) {
      const int *k = keys16 + 4 * (size_t)(b + loc);
      if (k[0] == key[0] && k[1] == key[1] && k[2] == key[2]) return indices[b + loc];
    }
  }
  return -1;
}

/* SparseGridView::decomposeCoord, SparseGrid.hpp:305-309: cell = coord & 7, block key = coord - cell (block ORIGIN in
 * cell coordinates), cell offset = (x*8 + y)*8 + z (local_coord_to_offset, :275-283) */
void zo_sg_decompose(const int coord[3], int block_key[3], int *cellno) {
  int c[3];
  for (int d = 0; d < 3; ++d) { c[d] = coord[d] & 7; block_key[d] = coord[d] - c[d]; }
  *cellno = (c[0] * 8 + c[1]) * 8 + c[2];
}

/* SparseGridView::valueOr(false_c, chn, indexCoord, default), SparseGrid.hpp:340-351; grid = TileVector<f32,512>:
 * element (chn, block, cell) at ((block * nch + chn) * 512 + cell) (container/TileVector.hpp:108) */
float zo_sg_value_or(int chn, const int coord[3], float dflt, int table_size, const uint32_t hf[6], const int *keys16,
                     const int *indices, const float *grid, int nch) {
  int bk[3], cno;
  zo_sg_decompose(coord, bk, &cno);
  const int bno = zo_bht_query(bk, table_size, hf, keys16, indices);
  return bno == -1 ? dflt : grid[((size_t)bno * nch + chn) * 512 + cno];
}

/* iCoord(bno, cno) = activeKeys[bno] + local_offset_to_coord(cno) (SparseGrid.hpp:266-272, 407-409);
 * wCoord = indexToWorld(iCoord) = X * transform (row vector times the 4x4 index-to-world matrix, :256-258,
 * math/Vec operator* with homogeneous w = 1: out_j = sum_i X_i M[i][j] + M[3][j]) */
void zo_sg_coords(int bno, int cno, const int *active_keys, const float m16[16], int icoord[3], float wcoord[3]) {
  int loc[3], off = cno;
  for (int d = 2; d >= 0; --d, off /= 8) loc[d] = off % 8;
  for (int d = 0; d < 3; ++d) icoord[d] = active_keys[3 * bno + d] + loc[d];
  for (int j = 0; j < 3; ++j) {
    float s = 0.f;
    for (int i = 0; i < 3; ++i) s += (float)icoord[i] * m16[4 * i + j];
    wcoord[j] = s + m16[12 + j];
  }
}

/* The side-8 partition the SparseGrid variant of the MPM path uses: the ComputeSparsity / EnlargeSparsity{0,2}
 * convention of simulation/sparsity/SparsityOp.hpp:58-112 with blockLen = 8 and keys stored as block-origin cell
 * coordinates (what SparseGrid's table holds, SparseGrid.hpp:305-309).  Serial order. */
int zo_sg_partition_build(int n, const float *x, float dx, int table_size, const uint32_t hf[6], int *keys16, int *indices,
                          int *status, int *active_keys, int *cnt) {
  zo_bht_clear(table_size, keys16, indices, status, cnt);
  const float dxinv = (float)1.0 / dx; /* SparsityOp.hpp:66 */
  for (int p = 0; p < n; ++p) {
    int key[3];
    for (int d = 0; d < 3; ++d) {
      const int c = (int)floorf(x[3 * p + d] * dxinv + 0.5f) + (-2); /* :73-74 */
      const int b = c + (c < 0 ? -8 + 1 : 0);                        /* :76 with blockLen = 8 */
      key[d] = (b / 8) * 8;                                          /* :77, then block index -> block origin */
    }
    zo_bht_insert(key, table_size, hf, keys16, indices, active_keys, cnt);
  }
  const int n0 = *cnt;
  for (int b = 0; b < n0; ++b)
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          const int key[3] = {active_keys[3 * b] + 8 * i, active_keys[3 * b + 1] + 8 * j, active_keys[3 * b + 2] + 8 * k};
          zo_bht_insert(key, table_size, hf, keys16, indices, active_keys, cnt);
        }
  return *cnt;
}

/* TileVector::reorderTiles, container/TileVector.hpp:641-691 (TileVectorTileReorder): scatter: ordered tile map[i] <- tile i;
 * gather: ordered tile i <- tile map[i].  tile_floats = numChannels * tile length. */
void zo_tilevector_reorder_tiles(const float *src, float *dst, int tile_floats, const int *map, int ntiles, int scatter) {
  for (int i = 0; i < ntiles; ++i) {
    const int j = map[i];
    const float *s = src + (size_t)(scatter ? i : j) * tile_floats;
    float *d = dst + (size_t)(scatter ? j : i) * tile_floats;
    memcpy(d, s, sizeof(float) * (size_t)tile_floats);
  }
}

/* bht::reorder, container/Bht.hpp:343-400 (ReorderBht): the ordered key list and the index stored with every key are
 * renumbered through map (scatter: old i -> new map[i]; gather: new i <- old map[i]); ordered_keys replaces activeKeys. */
void zo_bht_reorder(int table_size, const uint32_t hf[6], const int *keys16, int *indices, const int *active_keys, int n,
                    const int *map, int scatter, int *ordered_keys) {
  const int nb = table_size / 16;
  for (int i = 0; i < n; ++i) {
    const int j = map[i];
    const int *key = active_keys + 3 * (scatter ? i : j);
    int *ok = ordered_keys + 3 * (scatter ? j : i);
    ok[0] = key[0]; ok[1] = key[1]; ok[2] = key[2];
    /* query(key, false_c): the slot */
    for (int iter = 0; iter < 3; ++iter) {
      const int b = (int)(zo_bht_hash(hf[2 * iter], hf[2 * iter + 1], key) % (uint32_t)nb) * 16;
      int loc = 0;
      for (; loc != 16; ++loc) {
        const int *k = keys16 + 4 * (size_t)(b + loc);
        if (k[0] == key[0] && k[1] == key[1] && k[2] == key[2]) break;
      }
      if (loc != 16) { indices[b + loc] = scatter ? j : i; break; }
    }
  }
}

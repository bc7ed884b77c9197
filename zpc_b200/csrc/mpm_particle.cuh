// Per-particle scatter / gather straight against the hashed grid in global memory — the reference's
// own formulation (simulation/transfer/P2G.hpp:107-125, G2P.hpp:56-66) with the 27 hash probes of
// unpack_coord_in_grid (simulation/Utils.hpp:20-28) reduced to the <= 8 distinct blocks of the stencil.
// Used by the any-order AoS kernels (mpm.cu) and as the stray path of the binned kernels.
#pragma once
#include "mpm_math.cuh"

namespace zpcp {

// Grid accessors: block side 2^S, tile = [nch][8^S] floats, cell offset (x << 2S) | (y << S) | z.
// LegacyGrid  = HashTable<i32,3,int> + Grids<f32,3,4>   (container/HashTable.hpp, geometry/Structure.hpp:49-61,851-859)
// SparseGrid8 = bht<i32,3,int,16> + TileVector<f32,512> (container/Bht.hpp, geometry/SparseGrid.hpp:275-309); the
//               table is keyed by the block ORIGIN in cell coordinates.
struct LegacyGrid {
  static constexpr int S = 2;
  zpc_hashtable_view tb;
  int *missing = nullptr;  // optional device status word: ORed with `missing_bit` when a stencil block is absent (binned callers)
  int missing_bit = 0;
  ZPC_HD int query(int bx, int by, int bz) const {
    return zpcm::table_query(bx, by, bz, tb.tableSize, tb.keys, tb.indices);
  }
};
struct SparseGrid8 {
  static constexpr int S = 3;
  zpc_bht_view tb;
  int *missing = nullptr;
  int missing_bit = 0;
  ZPC_HD int query(int bx, int by, int bz) const { return zpcm::bht_query(bx << 3, by << 3, bz << 3, tb); }
};
template <class G> ZPC_HD int cell_offset(int lx, int ly, int lz) {
  constexpr int M = (1 << G::S) - 1;
  return ((lx & M) << (2 * G::S)) | ((ly & M) << G::S) | (lz & M);
}

// tile offsets (in floats) of the 2x2x2 blocks around the stencil; -1 where not needed / absent
template <class G>
ZPC_HD void resolve_blocks(const int (&corner)[3], const G &g, int nch, long long (&off)[8]) {
  constexpr int S = G::S, M = (1 << S) - 1;
  const int b0x = corner[0] >> S, b0y = corner[1] >> S, b0z = corner[2] >> S;
  const bool nx = (corner[0] & M) >= M - 1, ny = (corner[1] & M) >= M - 1, nz = (corner[2] & M) >= M - 1;  // stencil spills over
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ox = i >> 2, oy = (i >> 1) & 1, oz = i & 1;
    const bool need = (!ox || nx) && (!oy || ny) && (!oz || nz);
    int id = -1;
    if (need) id = g.query(b0x + ox, b0y + oy, b0z + oz);
    off[i] = id < 0 ? -1 : ((long long)id * nch) << (3 * S);
  }
}

// scatter of one particle given its (already scaled: * -dt * D_inv) stress contribution — P2G.hpp:104-125
template <class G>
static __host__ __device__ __noinline__ void p2g_scatter_core(const float (&pos)[3], const float (&vel)[3], float mass, const float (&C)[9],
                                                     const float (&contrib)[9], const G &tb, float *tiles, int nch,
                                                     float dx) {
  constexpr int S = G::S, M = (1 << S) - 1, CS = 1 << (3 * S);  // CS = cells per block = channel stride
  zpcm::Arena ar;
  zpcm::arena_init(ar, dx, pos);
  long long off[8];
  resolve_blocks(ar.corner, tb, nch, off);
  const int lx0 = ar.corner[0] & M, ly0 = ar.corner[1] & M, lz0 = ar.corner[2] & M;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long o = off[((lx >> S) << 2) | ((ly >> S) << 1) | (lz >> S)];
        if (o < 0) {  // block absent from the partition: cannot happen after partition_build on these positions
#ifdef __CUDA_ARCH__
          if (tb.missing) atomicOr(tb.missing, tb.missing_bit);
#endif
          continue;
        }
        float *t = tiles + o + cell_offset<G>(lx, ly, lz);
        const float x0 = (float)i * dx - ar.local[0], x1 = (float)j * dx - ar.local[1], x2 = (float)k * dx - ar.local[2];
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
        zpcm::grid_add(t, mass * W);
        const float Wm = W * mass;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          zpcm::grid_add(t + (1 + d) * CS, Wm * (vel[d] + (C[d] * x0 + C[3 + d] * x1 + C[6 + d] * x2)));
          zpcm::grid_add(t + (4 + d) * CS, (contrib[d] * x0 + contrib[3 + d] * x1 + contrib[6 + d] * x2) * W);
        }
      }
}

// FixedCorotatedConfig (P2G.hpp:88-91)
template <class G>
static ZPC_HD void p2g_scatter_particle(const float (&pos)[3], const float (&vel)[3], float mass, const float (&C)[9],
                                                  const float (&F)[9], const G &tb, float *tiles, int nch,
                                                  float dx, float dt, float volume, float mu, float lam) {
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float contrib[9];
  zpcm::stress_fcr(volume, mu, lam, F, contrib);
#pragma unroll
  for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
  p2g_scatter_core(pos, vel, mass, C, contrib, tb, tiles, nch, dx);
}

// VonMisesFixedCorotatedConfig (P2G.hpp:89-90)
template <class G>
static ZPC_HD void p2g_scatter_particle_vm(const float (&pos)[3], const float (&vel)[3], float mass,
                                                               const float (&C)[9], const float (&F)[9], const G &tb, float *tiles,
                                                               int nch, float dx, float dt, float volume, float mu, float lam,
                                                               float yield_stress) {
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float contrib[9];
  zpcm::stress_vonmises(volume, mu, lam, yield_stress, F, contrib);
#pragma unroll
  for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
  p2g_scatter_core(pos, vel, mass, C, contrib, tb, tiles, nch, dx);
}

// EquationOfStateConfig (P2G.hpp:66-87): weakly compressible fluid, J instead of F; gamma is fixed to 7 by the reference
template <class G>
static ZPC_HD void p2g_scatter_particle_eos(const float (&pos)[3], const float (&vel)[3], float mass,
                                                                const float (&C)[9], float J, const G &tb,
                                                                float *tiles, int nch, float dx, float dt, float volume, float bulk,
                                                                float viscosity) {
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float contrib[9];
  zpcm::eos_contrib(C, J, volume, bulk, viscosity, contrib);
#pragma unroll
  for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
  p2g_scatter_core(pos, vel, mass, C, contrib, tb, tiles, nch, dx);
}

// vel = sum W v_i ; G[r + 3e] = sum W v_i[r] * o_e   (o = stencil offset 0..2), so that
// C[r + 3e] = D_inv * (dx * G[r+3e] - local_e * vel[r])  ==  sum W v_i[r] * xixp[e] * D_inv  (G2P.hpp:65)
template <class GA>
static __host__ __device__ __noinline__ void g2p_gather_particle(const zpcm::Arena &ar, const GA &tb, const float *tiles, int nch,
                                                 float (&vel)[3], float (&G)[9]) {
  constexpr int S = GA::S, M = (1 << S) - 1, CS = 1 << (3 * S);
  long long off[8];
  resolve_blocks(ar.corner, tb, nch, off);
  const int lx0 = ar.corner[0] & M, ly0 = ar.corner[1] & M, lz0 = ar.corner[2] & M;
#pragma unroll
  for (int d = 0; d < 3; ++d) vel[d] = 0.f;
#pragma unroll
  for (int d = 0; d < 9; ++d) G[d] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long o = off[((lx >> S) << 2) | ((ly >> S) << 1) | (lz >> S)];
        if (o < 0) {
#ifdef __CUDA_ARCH__
          if (tb.missing) atomicOr(tb.missing, tb.missing_bit);
#endif
          continue;
        }
        const float *t = tiles + o + cell_offset<GA>(lx, ly, lz);
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
        const float oe[3] = {(float)i, (float)j, (float)k};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float wv = W * zpcm::grid_load(t + (1 + r) * CS);
          vel[r] += wv;
#pragma unroll
          for (int e = 0; e < 3; ++e) G[r + 3 * e] = fmaf(wv, oe[e], G[r + 3 * e]);
        }
      }
}

}  // namespace zpcp

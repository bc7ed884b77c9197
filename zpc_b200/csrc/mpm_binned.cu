// MPM fast path on block-binned AoSoA particles (sm_100a).
//
// Particles live in a 25-channel TileVector<f32,32> (AoSoA, reference container/TileVector.hpp:108),
// sorted by HOME BLOCK = the block ComputeSparsity assigns (reference SparsityOp.hpp:68-79, the block
// holding base_node-1).  With side-4 blocks the whole 3^3 stencil of such a particle lies in the 2x2x2
// block arena {b, b+1}^3, and a particle that has since drifted by one cell in any
// direction still lies inside it.  One CTA per bin:
//
//   P2G  (a) group the bin's particles by (column, z) of their CURRENT home cell: read the cell-order cache the last
//            binned G2P left behind, or counting-sort in shared memory (first step after a re-bin);
//        (b) per chunk of 256 particles: lanes act as particles (coalesced AoSoA loads through the cell order, SVD
//            stress, 28-float record -> shared), then warps take three non-empty cells at a time from a shared work
//            counter — lanes = 3 cells x 9 (x,y) node columns, three z-nodes x 7 channels per lane in registers —
//            and add the sums to the arena (eight [7][64] grid tiles in shared memory) with shared float atomics;
//        (c) the eight arena tiles are added to HBM with eight 1792-byte TMA bulk reductions
//            (cp.reduce.async.bulk.global.shared::cta.add.f32) instead of 27*7 REDs per particle;
//        (d) particles that left the arena's reach since the last re-bin take the per-particle RED path.
//   G2P  64-thread CTAs; the arena's three velocity channels (8 x 768 contiguous bytes) AND the particle channels G2P
//        reads (x, F: contiguous inside a TileVector tile) are staged with TMA bulk copies on mbarriers (two-stage
//        ring), then one thread per particle gathers with a separable (z, then y, then x) contraction: 240 FMAs
//        instead of 27*16, and leaves the bin grouped by new home cell for the next P2G.
//
// Results equal P2G.hpp / G2P.hpp up to fp32 re-association (tests: <= 1e-5 relative).
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "mpm_math.cuh"
#include "mpm_particle.cuh"
#include "p2g_sweep.cuh"

namespace {

constexpr int TS = 32;                  // particle tile length
constexpr int NCH = ZPC_PB_NCH;         // 25 channels
constexpr int BIN_MAX = ZPCB200_BIN_MAX;
#ifndef ZPC_P2G_NT
#define ZPC_P2G_NT 256
#endif
#ifndef ZPC_P2G_MINB
#define ZPC_P2G_MINB 4
#endif
#ifndef ZPC_P2G_EARLY   // 1: sweep variants >= 4 skip the Jacobi sweeps a whole warp has converged on (zpcm::stress_fcr_lean<true>)
#define ZPC_P2G_EARLY 1
#endif
#ifndef ZPC_P2G_PIPE    // 1: sweep variants >= 4 compute the next chunk's records before the end-of-sweep barrier (measured: no gain, see below)
#define ZPC_P2G_PIPE 0
#endif
constexpr int P2G_NT = ZPC_P2G_NT, P2G_NW = P2G_NT / 32;
constexpr int CHUNK = P2G_NT;           // particles staged per pass (one record per thread)
constexpr int NCOL6 = 36;               // (x,y) columns of home cells in [-1,4]^2: 16 nominal + 20 ring
constexpr int NGRP = NCOL6 * 6 + 1;     // (column, z in [-1,4]) groups + far-stray group
constexpr int GRP_FAR = NCOL6 * 6;
constexpr int REC_F = 28;               // floats per particle record

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ size_t pslot(size_t i) { return ((i >> 5) * NCH) * TS + (i & 31); }  // channel 0 of particle i

// ---- grid adaptors of the binned kernels --------------------------------------------------------------------------------------
// The binned kernels think in VIRTUAL 4^3 blocks: a bin is the particles whose home cell lies in one, the arena is the 2x2x2 virtual
// blocks {b, b+1}^3 staged as eight [7][64] tiles.  BinGridLegacy: a virtual block IS a block of Grids<f32,3,4> (HashTable keys =
// block coordinates, tile = 1 792 contiguous bytes: TMA bulk copies / reductions).  BinGridSparse (round 2, SURVEY §8 a12 / (f) rank
// 3): a virtual block is one octant of a side-8 block of SparseGrid<3,f32,8> (bht keys = block origins in cells, tile = [nch][512],
// cell offset (x*8+y)*8+z, geometry/SparseGrid.hpp:275-309); id = (block number << 3) | octant.  An octant's row of four z-cells is 16
// contiguous bytes, so the arena moves with 128-bit loads and 128-bit vector reductions (red.global.add.v4.f32) instead of TMA.
struct BinGridLegacy {
  static constexpr bool SPARSE = false;
  zpc_hashtable_view tb;
  __device__ __forceinline__ int query(int vx, int vy, int vz) const { return zpcm::table_query(vx, vy, vz, tb.tableSize, tb.keys, tb.indices); }
  __device__ __forceinline__ int count() const { return *tb.cnt; }
  __device__ __forceinline__ void key_of(int id, int &kx, int &ky, int &kz) const {
    kx = tb.activeKeys[3 * (size_t)id]; ky = tb.activeKeys[3 * (size_t)id + 1]; kz = tb.activeKeys[3 * (size_t)id + 2];
  }
  __device__ __forceinline__ zpcp::LegacyGrid accessor(int *status) const { return zpcp::LegacyGrid{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}; }
};
struct BinGridSparse {
  static constexpr bool SPARSE = true;
  zpc_bht_view tb;
  __device__ __forceinline__ int query(int vx, int vy, int vz) const {
    const int bno = zpcm::bht_query((vx >> 1) << 3, (vy >> 1) << 3, (vz >> 1) << 3, tb);   // arithmetic shifts: floor for negative coordinates
    return bno < 0 ? -1 : (bno << 3) | ((vx & 1) << 2) | ((vy & 1) << 1) | (vz & 1);
  }
  __device__ __forceinline__ int count() const { return *tb.cnt * 8; }
  __device__ __forceinline__ void key_of(int id, int &kx, int &ky, int &kz) const {
    const size_t b = (size_t)(id >> 3);
    kx = (tb.activeKeys[3 * b] >> 2) + ((id >> 2) & 1); ky = (tb.activeKeys[3 * b + 1] >> 2) + ((id >> 1) & 1); kz = (tb.activeKeys[3 * b + 2] >> 2) + (id & 1);
  }
  __device__ __forceinline__ zpcp::SparseGrid8 accessor(int *status) const { return zpcp::SparseGrid8{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}; }
  // float offset of row (x, y) (four z-cells) of channel ch of octant id in a [nch][512] tile array
  __device__ __forceinline__ static size_t row_offset(int id, int nch, int ch, int x, int y) {
    return ((size_t)(id >> 3) * nch + ch) * 512 + ((((id >> 2) & 1) * 4 + x) * 8 + (((id >> 1) & 1) * 4 + y)) * 8 + (id & 1) * 4;
  }
};


// fused halo send (include/zpcb200.h: zpc_halo_view): the arena tile in shared memory at `tile_smem` was just added to grid block `id`;
// if other ranks hold that block too, the same 1 792 bytes are reduce-added into its slot of their receive buffers (peer-mapped)
__device__ __forceinline__ bool halo_send_tile(const zpc_halo_view &halo, int id, const float *tile_smem) {
  bool sent = false;
  if (halo.peer) {
#pragma unroll
    for (int k = 0; k < ZPCB200_HALO_K; ++k) {
      const int r = halo.peer[(size_t)id * ZPCB200_HALO_K + k];
      if (r < 0) break;
      float *dst = halo.peers[r] + (((size_t)halo.half * halo.world + halo.rank) * halo.seg + halo.pos[(size_t)id * ZPCB200_HALO_K + k]) * 448;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(tile_smem)), "r"(1792)
                   : "memory");
      sent = true;
    }
  }
  return sent;
}

struct P2GSmem {
  float4 rec4[CHUNK * 7 + CHUNK / 8];  // 29184 B at CHUNK = 256: 7 float4 per record (+1 pad granule per 8 records, used by VAR 4)
  float out[8 * 448];               // 14336 B: the arena as eight [7][64] grid tiles, accumulated with shared atomics
  unsigned short order[BIN_MAX];    // fallback grouping only (no cell-order cache)
  unsigned char grp_of[BIN_MAX];
  int cnt[NGRP + 3];
  int gstart[NGRP + 3];
  int tile_id[8];
  int next_unit;
  int ncells[2];                    // v4 sweep: non-empty cells of the current / next chunk
  unsigned char cells[2][NGRP + 7];
};
static_assert(sizeof(P2GSmem) <= (227 * 1024) / ZPC_P2G_MINB - 1024, "ZPC_P2G_MINB CTAs per SM (227 KB, 1 KB reserved per CTA)");

// sweep units in scheduling order: the 16 nominal columns first, then the 20 ring columns; c6 = (x+1)*6 + (y+1)
__constant__ unsigned char c_unit_c6[NCOL6] = {7,  8,  9,  10, 13, 14, 15, 16, 19, 20, 21, 22, 25, 26, 27, 28,
                                              0,  1,  2,  3,  4,  5,  6,  11, 12, 17, 18, 23, 24, 29, 30, 31, 32, 33, 34, 35};

// MODEL 0 = FixedCorotatedConfig, 1 = VonMisesFixedCorotatedConfig (yield_stress; P2G.hpp:89-90), 2 = DruckerPragerConfig,
// 3 = NACCConfig (pp; the per-particle logJp lives in `scalar`, one float per particle in BIN order, read and written back by the
// record phase like P2G.hpp:93,101), 4 = EquationOfStateConfig (pp.a = bulk, pp.b = viscosity; `scalar` = J, read only; the F
// channels of the bins are not touched, P2G.hpp:66-87) — the model only enters
// the records phase (and the stray path), the sweep and the write-back are the same
template <int VAR, int MODEL = 0, class BG = BinGridLegacy>
__global__ void __launch_bounds__(P2G_NT, ZPC_P2G_MINB)
p2g_binned_kernel(const float *__restrict__ pars, const int *__restrict__ binStart, const int *__restrict__ binKey,
                  const int *__restrict__ numBins, const unsigned short *__restrict__ cellOrder,
                  const unsigned short *__restrict__ cellStart, const int *__restrict__ cellOrderValid, BG bg,
                  float *__restrict__ tiles, float dx, float dt, float volume, float mu, float lam, int prefetch,
                  float yield_stress, float *__restrict__ scalar, zpcm::PlasticPrm pp, int *__restrict__ status, zpc_halo_view halo,
                  int nch = 7) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  P2GSmem &S = *reinterpret_cast<P2GSmem *>(smem_raw);
  const int bin = blockIdx.x;
  if (bin >= *numBins) return;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int p0 = binStart[bin], np = min(binStart[bin + 1] - p0, BIN_MAX);
  const int kx = binKey[3 * bin], ky = binKey[3 * bin + 1], kz = binKey[3 * bin + 2];
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  if (prefetch && np > 0) {
    // every 128-byte line of the particle tiles this bin overlaps (25 channel rows per 32-particle tile) is requested
    // into L2 now: the record phases read them through the cell order (a gather), after the lookups below
    // (7.62 -> 7.36 ms at C3; requesting the lines of the bin one wave of CTAs ahead instead was slower, 7.45 ms)
    const int t0 = p0 >> 5, nlines = (((p0 + np - 1) >> 5) - t0 + 1) * NCH;
    for (int i = tid; i < nlines; i += P2G_NT) {
      const float *a = pars + ((size_t)t0 * NCH + i) * TS;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }

  // ---- (0) arena blocks, zero the accumulation tiles ---------------------------------------------------
  if (tid < 8) S.tile_id[tid] = bg.query(kx + (tid >> 2), ky + ((tid >> 1) & 1), kz + (tid & 1));
  {
    float4 *z = reinterpret_cast<float4 *>(S.out);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 8 * 448 / 4; i += P2G_NT) z[i] = zero;
  }
  if (tid == 0) S.next_unit = 0;

  // ---- (a) group the bin's particles by (column, z) of their CURRENT home cell ----------------------------
  // either read the grouping the last binned G2P left behind, or counting-sort here
  const bool pre = cellOrder != nullptr && *cellOrderValid != 0;
  const unsigned short *gorder = pre ? cellOrder + p0 : S.order;
  if (pre) {
    for (int i = tid; i <= NGRP; i += P2G_NT) S.gstart[i] = cellStart[(size_t)bin * ZPCB200_CELL_GROUPS_PAD + i];
    __syncthreads();
  } else {
    for (int i = tid; i < NGRP + 3; i += P2G_NT) S.cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < np; i += P2G_NT) {
      const size_t s = pslot((size_t)p0 + i);
      // current base node (division form, as LocalArena does), minus one = home cell, relative to the bin's block origin
      const int cx = (int)floorf(pars[s + (ZPC_PB_X + 0) * TS] / dx - 0.5f) - 1 - 4 * kx;
      const int cy = (int)floorf(pars[s + (ZPC_PB_X + 1) * TS] / dx - 0.5f) - 1 - 4 * ky;
      const int cz = (int)floorf(pars[s + (ZPC_PB_X + 2) * TS] / dx - 0.5f) - 1 - 4 * kz;
      const int g = ((unsigned)(cx + 1) < 6u && (unsigned)(cy + 1) < 6u && (unsigned)(cz + 1) < 6u)
                        ? ((cx + 1) * 6 + (cy + 1)) * 6 + (cz + 1)
                        : GRP_FAR;
      S.grp_of[i] = (unsigned char)g;
      atomicAdd(&S.cnt[g], 1);
    }
    __syncthreads();
    if (w == 0) {  // exclusive scan of the 217 counters, 7 per lane
      int c[7], sum = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { const int g = l * 7 + k; c[k] = g < NGRP ? S.cnt[g] : 0; sum += c[k]; }
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (l >= d) inc += t; }
      int run = inc - sum;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int g = l * 7 + k;
        if (g <= NGRP) { S.gstart[g] = run; S.cnt[g] = run; }
        run += c[k];
      }
    }
    __syncthreads();
    for (int i = tid; i < np; i += P2G_NT) S.order[atomicAdd(&S.cnt[S.grp_of[i]], 1)] = (unsigned short)i;
    __syncthreads();
  }

  // ---- (b) per chunk: records (one thread per particle), then cell sweeps ------------------------------------
  const int n_fast = S.gstart[GRP_FAR];
  // VAR 3: lanes = the 27 stencil offsets of one cell.  VAR 4: lanes = 3 cells x 9 (ox,oy) node columns.
  const bool lane_on = l < 27;
  const int lc = lane_on ? l : 0;  // VAR 3: lanes 27..31 shadow lane 0 and never write
  const int ox = VAR >= 4 ? (lc % 9) / 3 : lc / 9, oy = VAR >= 4 ? lc % 3 : (lc / 3) % 3, oz = lc % 3;
  const int gi = l / 9;            // VAR 4: cell slot of this lane (3 = idle lanes 27..31)
  zpcs::LaneCoef L;
  {
    // quadratic B-spline as a polynomial in d0 (InterpolationKernel.hpp:105-113):
    //   o=0: .5 d^2 - 1.5 d + 1.125 ; o=1: -d^2 + 2 d - .25 ; o=2: .5 d^2 - .5 d + .125
    L.ax = ox == 1 ? -1.0f : 0.5f; L.bx = ox == 0 ? -1.5f : (ox == 1 ? 2.0f : -0.5f); L.cx = ox == 0 ? 1.125f : (ox == 1 ? -0.25f : 0.125f);
    L.ay = oy == 1 ? -1.0f : 0.5f; L.by = oy == 0 ? -1.5f : (oy == 1 ? 2.0f : -0.5f); L.cy = oy == 0 ? 1.125f : (oy == 1 ? -0.25f : 0.125f);
    L.az = oz == 1 ? -1.0f : 0.5f; L.bz = oz == 0 ? -1.5f : (oz == 1 ? 2.0f : -0.5f); L.cz = oz == 0 ? 1.125f : (oz == 1 ? -0.25f : 0.125f);
    L.fx = (float)ox; L.fy = (float)oy; L.fz = (float)oz;
  }
  // VAR 4: list of the cells that have particles in the chunk [cb, cb+CHUNK), in group order; run by one whole warp.
  // (Ordering the list by particle count — so that the three cells a warp sweeps in lockstep have equal length — was
  // measured slower: 8.0 ms vs 7.8 ms at C3; consecutive cells of one column share arena nodes and smem banks better.)
  auto build_cell_list = [&](int buf, int cb) {
    const int ce0 = min(cb + CHUNK, n_fast);
    unsigned char *out = S.cells[buf];
    int base = 0;
#pragma unroll 1
    for (int g = l; g < GRP_FAR + 31 - (GRP_FAR + 31) % 32; g += 32) {
      const bool ne = g < GRP_FAR && max(S.gstart[g], cb) < min(S.gstart[g + 1], ce0);
      const unsigned m = __ballot_sync(0xffffffffu, ne);
      if (ne) out[base + __popc(m & lanemask_lt())] = (unsigned char)g;
      base += __popc(m);
    }
    if (l == 0) S.ncells[buf] = base;
  };
  if (VAR >= 4 && w == P2G_NW - 1 && n_fast > 0) build_cell_list(0, 0);
  // One record per thread, computed into registers and stored to shared memory at the top of the chunk.  The stress and the affine
  // coefficients need nothing from shared memory, so with ZPC_P2G_PIPE=1 the NEXT chunk's records are computed right after a warp runs
  // out of sweep units, in the time it would otherwise wait at the end-of-sweep barrier (ncu, C3: 15 % of all warp samples sit at
  // that barrier; 11 units for 8 warps).  Measured at C3 (benchmarks/r2_s2_call4.sh): 6.81 ms against 6.78 without — the other three
  // CTAs of the SM already fill those slots, and carrying the record across the barrier costs 20 local stores; off by default.
  float4 R[7];
  auto compute_record = [&](int cb) {
    const int pos = cb + tid;
    if (pos < n_fast) {
      const size_t s = pslot((size_t)p0 + gorder[pos]);
      float F[9], K[9];
      if constexpr (MODEL != 4) {
#pragma unroll
        for (int d = 0; d < 9; ++d) F[d] = pars[s + (ZPC_PB_F + d) * TS];
      }
      if constexpr (MODEL == 4) {
        // stress from C and J (P2G.hpp:66-83); C is loaded again below for the affine part — the compiler merges the loads
        const float J = scalar[(size_t)p0 + gorder[pos]];
        float Cc[9];
#pragma unroll
        for (int d = 0; d < 9; ++d) Cc[d] = pars[s + (ZPC_PB_C + d) * TS];
        zpcm::eos_contrib(Cc, J, volume, pp.a, pp.b, K);
      } else if constexpr (MODEL == 1) zpcm::stress_vonmises(volume, mu, lam, yield_stress, F, K);
      else if constexpr (MODEL == 2 || MODEL == 3) {
        float *lj = scalar + (size_t)p0 + gorder[pos];
        float logJp = *lj;
        if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, pp.a, pp.b, pp.c, pp.flag != 0, logJp, F, K);
        else zpcm::stress_nacc(volume, mu, pp.a, pp.b, pp.c, pp.d, pp.flag != 0, logJp, F, K);
        *lj = logJp;
      } else zpcm::stress_fcr_lean<(VAR >= 4 && ZPC_P2G_EARLY != 0)>(volume * (-dt * D_inv), mu, lam, F, K);
      if constexpr (MODEL != 0) {
#pragma unroll
        for (int d = 0; d < 9; ++d) K[d] = K[d] * -dt * D_inv;
      }
      float d0[3], loc[3], vel[3], C[9];
      const float mass = pars[s + ZPC_PB_M * TS];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float X = zpcm::div_exact(pars[s + (ZPC_PB_X + d) * TS], dx, dx_inv);
        const float lp = X - floorf(X - 0.5f);
        d0[d] = lp;
        loc[d] = lp * dx;
        vel[d] = pars[s + (ZPC_PB_V + d) * TS];
      }
#pragma unroll
      for (int d = 0; d < 9; ++d) C[d] = pars[s + (ZPC_PB_C + d) * TS];
      // mv_d = W (A_d + sum_e B_de o_e), rhs_d = W (a_d + sum_e K_de o_e), o = stencil offset (0,1,2)^3
      float A[3], a[3], B[9], Kd[9];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        A[d] = mass * (vel[d] - (C[d] * loc[0] + C[3 + d] * loc[1] + C[6 + d] * loc[2]));
        a[d] = -(K[d] * loc[0] + K[3 + d] * loc[1] + K[6 + d] * loc[2]);
#pragma unroll
        for (int e = 0; e < 3; ++e) { B[3 * d + e] = mass * C[d + 3 * e] * dx; Kd[3 * d + e] = K[d + 3 * e] * dx; }
      }
      zpcs::write_record<VAR>(R, d0, mass, A, a, B, Kd);
    }
  };
  constexpr bool PIPE = VAR >= 4 && ZPC_P2G_PIPE != 0;
  if (PIPE && n_fast > 0) compute_record(0);
  for (int cb = 0; cb < n_fast; cb += CHUNK) {
    const int buf = (cb / CHUNK) & 1;
    if (!PIPE) compute_record(cb);
    if (cb + tid < n_fast) {
      float4 *dst = S.rec4 + zpcs::rec_at<VAR>(tid);
#pragma unroll
      for (int q = 0; q < 7; ++q) dst[q] = R[q];
    }
    __syncthreads();
    const int ce = min(cb + CHUNK, n_fast);
    if (VAR >= 4) {
      // cell triples are handed out dynamically (one shared counter); each lane group sweeps its own cell, the sums
      // of the lane's three z-nodes go into the arena tiles with shared-memory float atomics
      const int ncells = S.ncells[buf];
      const unsigned char *cells = S.cells[buf];
      zpcs::ColCoef Lc = {L.ax, L.bx, L.cx, L.ay, L.by, L.cy, L.fx, L.fy};
      while (true) {
        int u = 0;
        if (l == 0) u = atomicAdd(&S.next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (3 * u >= ncells) {
          // the first warp to run out of work prepares the next chunk's list while the others finish their sweeps
          if (3 * (u - 1) < ncells && cb + CHUNK < n_fast) build_cell_list(buf ^ 1, cb + CHUNK);
          break;
        }
        const int ci = 3 * u + gi;
        const bool have = gi < 3 && ci < ncells;
        const int g = have ? (int)cells[ci] : 0;
        const int lo = have ? max(S.gstart[g], cb) - cb : 0, hi = have ? min(S.gstart[g + 1], ce) - cb : 0;
        const int nmax = __reduce_max_sync(0xffffffffu, hi - lo);
        float acc[7][3];
        if constexpr (VAR == 5) {
          float accm[3] = {0.f, 0.f, 0.f};
          float2 accp[3][3];
#pragma unroll
          for (int q = 0; q < 3; ++q) { accp[q][0] = make_float2(0.f, 0.f); accp[q][1] = make_float2(0.f, 0.f); accp[q][2] = make_float2(0.f, 0.f); }
          zpcs::sweep_cells3_packed(S.rec4, lo, hi, nmax, Lc, accm, accp);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            acc[0][k] = accm[k];
#pragma unroll
            for (int q = 0; q < 3; ++q) { acc[1 + 2 * q][k] = accp[q][k].x; acc[2 + 2 * q][k] = accp[q][k].y; }
          }
        } else {
#pragma unroll
          for (int ch = 0; ch < 7; ++ch) { acc[ch][0] = 0.f; acc[ch][1] = 0.f; acc[ch][2] = 0.f; }
          zpcs::sweep_cells3(S.rec4, lo, hi, nmax, Lc, acc);
        }
        if (have) {
          const int c6 = g / 6, zc = g - 6 * c6;                     // g = (cx+1)*36 + (cy+1)*6 + (cz+1)
          const int axn = c6 / 6 + ox, ayn = c6 % 6 + oy;            // arena node (cx+1+ox, cy+1+oy, cz+1+k)
          const int xy_off = (((axn >> 2) << 2) | ((ayn >> 2) << 1)) * 448 + (((axn & 3) << 4) | ((ayn & 3) << 2));
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int azn = zc + k;
            float *dstn = S.out + xy_off + (azn >> 2) * 448 + (azn & 3);
#pragma unroll
            for (int ch = 0; ch < 7; ++ch) atomicAdd(dstn + ch * 64, acc[ch][k]);
          }
        }
      }
    } else {
      // columns are handed out dynamically (one shared counter): warps that finish early take the next column, the
      // per-cell sums go straight into the arena tiles with shared-memory float atomics (no private arenas, no merge)
      while (true) {
        int u = 0;
        if (l == 0) u = atomicAdd(&S.next_unit, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= NCOL6) break;
        const int c6 = c_unit_c6[u];
        const int g0 = c6 * 6;
        if (S.gstart[g0 + 6] <= cb || S.gstart[g0] >= ce) continue;  // nothing of this column in the chunk
        const int axn = c6 / 6 + ox, ayn = c6 % 6 + oy;              // arena node (x+1+ox, y+1+oy, zc+oz)
        const int xy_off = (((axn >> 2) << 2) | ((ayn >> 2) << 1)) * 448 + (((axn & 3) << 4) | ((ayn & 3) << 2));
  #pragma unroll 1
        for (int zc = 0; zc < 6; ++zc) {
          const int lo = max(S.gstart[g0 + zc], cb), hi = min(S.gstart[g0 + zc + 1], ce);
          if (lo >= hi) continue;
          float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          zpcs::sweep_cell(S.rec4 - 7 * cb, lo, hi, L, acc);
          if (lane_on) {
            const int azn = zc + oz;
            float *dstn = S.out + xy_off + (azn >> 2) * 448 + (azn & 3);
  #pragma unroll
            for (int ch = 0; ch < 7; ++ch) atomicAdd(dstn + ch * 64, acc[ch]);
          }
        }
      }
    }
    if (PIPE && cb + CHUNK < n_fast) compute_record(cb + CHUNK);   // registers only: overlaps the other warps' sweeps
    __syncthreads();
    if (tid == 0) S.next_unit = 0;  // ordered before the next sweep by the barrier after the next records phase
  }

  // ---- (c) add the eight arena tiles to the grid ------------------------------------------------------------------
  if constexpr (BG::SPARSE) {
    // SparseGrid: 8 tiles x 7 channels x 16 (x,y) rows of four z-cells, one 128-bit vector reduction each; rows nothing was added to are skipped
    for (int q = tid; q < 8 * 112; q += P2G_NT) {
      const int t = q / 112, r = q - 112 * t, ch = r >> 4, row = r & 15;
      const int id = S.tile_id[t];
      const float4 v = *reinterpret_cast<const float4 *>(S.out + t * 448 + ch * 64 + row * 4);
      if (id >= 0 && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)) {
        float *g = tiles + BG::row_offset(id, nch, ch, row >> 2, row & 3);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
  __syncthreads();
  bool remote = false;
  if (tid < 8) {
    const int id = S.tile_id[tid];
    if (BG::SPARSE) {
      if (id < 0 && status) {
        bool any = false;
        for (int c = 0; c < 64; ++c) any |= S.out[tid * 448 + c] != 0.f;
        if (any) atomicOr(status, ZPC_BINS_STENCIL_BLOCK_MISSING);
      }
    } else if (id >= 0) {
      float *g = tiles + (size_t)id * 448;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g),
                   "r"(smem_u32(S.out + tid * 448)), "r"(1792)
                   : "memory");
      remote = halo_send_tile(halo, id, S.out + tid * 448);
    } else if (status) {  // an arena block the partition does not hold: flag it if any mass was headed there
      bool any = false;
      for (int c = 0; c < 64; ++c) any |= S.out[tid * 448 + c] != 0.f;
      if (any) atomicOr(status, ZPC_BINS_STENCIL_BLOCK_MISSING);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }

  // ---- (d) far strays (moved more than one cell since the re-bin): per-particle scatter with REDs --------------
  for (int t = n_fast + tid; t < np; t += P2G_NT) {
    const size_t s = pslot((size_t)p0 + gorder[t]);
    float pos[3], vel[3], C[9], F[9];
    const float mass = pars[s + ZPC_PB_M * TS];
#pragma unroll
    for (int d = 0; d < 3; ++d) { pos[d] = pars[s + (ZPC_PB_X + d) * TS]; vel[d] = pars[s + (ZPC_PB_V + d) * TS]; }
#pragma unroll
    for (int d = 0; d < 9; ++d) { C[d] = pars[s + (ZPC_PB_C + d) * TS]; F[d] = pars[s + (ZPC_PB_F + d) * TS]; }
    if constexpr (MODEL == 1) zpcp::p2g_scatter_particle_vm(pos, vel, mass, C, F, bg.accessor(status), tiles, nch, dx, dt, volume, mu, lam, yield_stress);
    else if constexpr (MODEL == 4) {
      zpcp::p2g_scatter_particle_eos(pos, vel, mass, C, scalar[(size_t)p0 + gorder[t]], bg.accessor(status), tiles, nch, dx, dt, volume, pp.a, pp.b);
    } else if constexpr (MODEL >= 2) {
      float *lj = scalar + (size_t)p0 + gorder[t];
      float logJp = *lj, contrib[9];
      if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, pp.a, pp.b, pp.c, pp.flag != 0, logJp, F, contrib);
      else zpcm::stress_nacc(volume, mu, pp.a, pp.b, pp.c, pp.d, pp.flag != 0, logJp, F, contrib);
      *lj = logJp;
#pragma unroll
      for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
      zpcp::p2g_scatter_core(pos, vel, mass, C, contrib, bg.accessor(status), tiles, nch, dx);
    } else zpcp::p2g_scatter_particle(pos, vel, mass, C, F, bg.accessor(status), tiles, nch, dx, dt, volume, mu, lam);
  }
  if (tid < 8) {
    // smem must outlive the bulk reads; a tile that also went to a peer waits for the writes themselves, so that the kernel's
    // completion (and the barrier the caller puts after it) orders them before the peer's grid update
    if (remote) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Binned P2G, plane sweep (round 2; zpcs::sweep_plane): same bins, same TMA write-back as p2g_binned_kernel, but
//   * a lane of the sweep owns one x-plane of a cell's stencil (9 nodes x 7 channels in registers), ten cells per warp round:
//     a record is pulled out of shared memory by three lanes instead of nine — the data path, not issue, bounded the column sweep;
//   * ATOMIC-FREE accumulation: the bin's 6 x 6 cell columns are split into eight fixed regions (four x-slabs x two y-halves), one
//     per warp; a warp adds its sums with plain read-modify-writes to a PRIVATE copy of the part of the arena its region can reach
//     (4 x 5 x 8 nodes x 7 channels = 4 480 B) — lanes of one flush instruction hold different cells or different planes, so their
//     addresses differ, and nobody else writes there.  (Shared-memory float atomics are CAS loops on sm_100a: measured 5.9 clk per
//     warp-op on an idle SM against 2.4 for a plain RMW, benchmarks/micro/ffma2.cu; in the kernel, 7.7 wavefronts + 1.8 rounds each.)
//     After the last chunk the eight private copies are summed node by node — fixed order — straight into the eight [7][64] grid
//     tiles the TMA write-back reads;
//   * chunks of 512 particles (a whole nominal bin), two records per thread, the fixed-corotated stress in its lean form
//     (zpcm::stress_fcr_lean); two CTAs per SM at <= 128 registers.
constexpr int PL_NT = 256, PL_NW = PL_NT / 32, PL_CHUNK = 512, PL_RPT = PL_CHUNK / PL_NT, PL_UNIT = 10;
static_assert(PL_NW == 8, "eight warp regions");
// warp w = 2 a + h owns the cell columns cx' = cx + 1 in xlo(a) .. xlo(a) + ncx(a) - 1, cy' = cy + 1 in 3 h .. 3 h + 2 (all cz):
// x-slabs {0,1} {2} {3} {4,5} — the ring columns ride with the outer nominal slabs — and the arena nodes X0 .. X0 + nX - 1, 3 h .. 3 h + 4
__host__ __device__ constexpr int pl_xlo(int a) { return a == 0 ? 0 : a + 1; }
__host__ __device__ constexpr int pl_ncx(int a) { return (a == 0 || a == 3) ? 2 : 1; }
constexpr int PL_PRIV = 7 * 160;   // floats per private region: [channel][lx 0..3][ly 0..4][8 z], z rotated by 4 ly (bank spreading)
__device__ __forceinline__ int priv_idx(int lx, int ly, int z) { return lx * 40 + ly * 8 + ((z + 4 * ly) & 7); }
struct P2GPlaneSmem {
  float4 rec4[PL_CHUNK * 7 + PL_CHUNK / 8];  // 58368 B; reused as the eight [7][64] grid tiles of the write-back
  float priv[PL_NW * PL_PRIV];               // 35840 B
  unsigned short order[BIN_MAX];             // fallback grouping only (no cell-order cache)
  unsigned short rank[BIN_MAX];              // slot of the bin -> position in the (column, z) order (inverse of the cell order)
  unsigned char grp_of[BIN_MAX];
  int cnt[NGRP + 3];
  int gstart[NGRP + 3];
  int tile_id[8];
  unsigned char wcells[PL_NW][40];           // per warp: the non-empty cells of its region in the current chunk
};
static_assert(sizeof(P2GPlaneSmem) <= 112 * 1024, "two CTAs per SM");

// the 28 numbers of one particle's record from its 25 channels pd (ZPC_PB_* order): stress of MODEL, then
// mv_d = W (A_d + B_d. o), rhs_d = W (a_d + Kd_d. o)
template <int MODEL>
__device__ __forceinline__ void particle_record(const float (&pd)[NCH], float *__restrict__ sc, bool live, float dx, float dx_inv, float dt,
                                                float D_inv, float volume, float mu, float lam, float yield_stress,
                                                const zpcm::PlasticPrm &pp, float (&d0)[3], float &mass, float (&A)[3], float (&a)[3],
                                                float (&B)[9], float (&Kd)[9]) {
  float F[9], K[9], C[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) { C[d] = pd[ZPC_PB_C + d]; F[d] = pd[ZPC_PB_F + d]; }
  if constexpr (MODEL == 4) zpcm::eos_contrib(C, *sc, volume, pp.a, pp.b, K);
  else if constexpr (MODEL == 1) zpcm::stress_vonmises(volume, mu, lam, yield_stress, F, K);
  else if constexpr (MODEL == 2 || MODEL == 3) {
    float logJp = *sc;
    if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, pp.a, pp.b, pp.c, pp.flag != 0, logJp, F, K);
    else zpcm::stress_nacc(volume, mu, pp.a, pp.b, pp.c, pp.d, pp.flag != 0, logJp, F, K);
    if (live) *sc = logJp;
  } else zpcm::stress_fcr_lean<false>(volume * (-dt * D_inv), mu, lam, F, K);
  if constexpr (MODEL != 0) {
#pragma unroll
    for (int d = 0; d < 9; ++d) K[d] = K[d] * -dt * D_inv;
  }
  float loc[3], vel[3];
  mass = pd[ZPC_PB_M];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float X = zpcm::div_exact(pd[ZPC_PB_X + d], dx, dx_inv);
    const float lp = X - floorf(X - 0.5f);
    d0[d] = lp;
    loc[d] = lp * dx;
    vel[d] = pd[ZPC_PB_V + d];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    A[d] = mass * (vel[d] - (C[d] * loc[0] + C[3 + d] * loc[1] + C[6 + d] * loc[2]));
    a[d] = -(K[d] * loc[0] + K[3 + d] * loc[1] + K[6 + d] * loc[2]);
#pragma unroll
    for (int e = 0; e < 3; ++e) { B[3 * d + e] = mass * C[d + 3 * e] * dx; Kd[3 * d + e] = K[d + 3 * e] * dx; }
  }
}

template <int MODEL>
__global__ void __launch_bounds__(PL_NT, 2)
p2g_plane_kernel(const float *__restrict__ pars, const int *__restrict__ binStart, const int *__restrict__ binKey,
                 const int *__restrict__ numBins, const unsigned short *__restrict__ cellOrder,
                 const unsigned short *__restrict__ cellStart, const int *__restrict__ cellOrderValid, zpc_hashtable_view tb,
                 float *__restrict__ tiles, float dx, float dt, float volume, float mu, float lam, float yield_stress,
                 float *__restrict__ scalar, zpcm::PlasticPrm pp, int prefetch_ahead, int *__restrict__ status, zpc_halo_view halo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  P2GPlaneSmem &S = *reinterpret_cast<P2GPlaneSmem *>(smem_raw);
  const int bin = blockIdx.x;
  if (bin >= *numBins) return;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int p0 = binStart[bin], np = min(binStart[bin + 1] - p0, BIN_MAX);
  const int kx = binKey[3 * bin], ky = binKey[3 * bin + 1], kz = binKey[3 * bin + 2];
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  // the records are computed in STORAGE order — thread t owns slots t and t + 256 of every block of 512 — and written to the shared
  // record buffer at the particle's position in the cell order: every global load is a full 128-byte line, nothing is gathered, and
  // the loads of the first block are issued before anything else so that the lookups and the barriers below hide their latency
  float pd[PL_RPT][NCH];
#pragma unroll
  for (int r = 0; r < PL_RPT; ++r) {
    const int i = tid + r * PL_NT;
    const size_t s = pslot((size_t)p0 + i);
#pragma unroll
    for (int c = 0; c < NCH; ++c) pd[r][c] = i < np ? pars[s + c * TS] : 0.f;
  }
  if (prefetch_ahead > 0 && bin + prefetch_ahead < *numBins) {
    // the CTA that will run about one wave from now (two CTAs per SM are resident): ask for its particle lines and its cell order now,
    // so that its entry loads hit L2 instead of HBM
    const int q0 = binStart[bin + prefetch_ahead], qn = min(binStart[bin + prefetch_ahead + 1] - q0, BIN_MAX);
    if (qn > 0) {
      const int t0 = q0 >> 5, nlines = (((q0 + qn - 1) >> 5) - t0 + 1) * NCH;
      for (int i = tid; i < nlines; i += PL_NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(pars + ((size_t)t0 * NCH + i) * TS));
      if (cellOrder && tid < (qn + 63) / 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(cellOrder + q0 + 64 * tid));
    }
  }
  if (tid < 8) S.tile_id[tid] = zpcm::table_query(kx + (tid >> 2), ky + ((tid >> 1) & 1), kz + (tid & 1), tb.tableSize, tb.keys, tb.indices);
  {
    float4 *z = reinterpret_cast<float4 *>(S.priv);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < PL_NW * PL_PRIV / 4; i += PL_NT) z[i] = zero;
  }

  // ---- (a) grouping by (column, z) of the current home cell: the cache the last binned G2P left, or a counting sort here -------
  const bool pre = cellOrder != nullptr && *cellOrderValid != 0;
  const unsigned short *gorder = pre ? cellOrder + p0 : S.order;
  if (pre) {
    for (int i = tid; i <= NGRP; i += PL_NT) S.gstart[i] = cellStart[(size_t)bin * ZPCB200_CELL_GROUPS_PAD + i];
    for (int i = tid; i < np; i += PL_NT) S.rank[gorder[i]] = (unsigned short)i;
    __syncthreads();
  } else {
    for (int i = tid; i < NGRP + 3; i += PL_NT) S.cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < np; i += PL_NT) {
      const size_t s = pslot((size_t)p0 + i);
      const int cx = (int)floorf(pars[s + (ZPC_PB_X + 0) * TS] / dx - 0.5f) - 1 - 4 * kx;
      const int cy = (int)floorf(pars[s + (ZPC_PB_X + 1) * TS] / dx - 0.5f) - 1 - 4 * ky;
      const int cz = (int)floorf(pars[s + (ZPC_PB_X + 2) * TS] / dx - 0.5f) - 1 - 4 * kz;
      const int g = ((unsigned)(cx + 1) < 6u && (unsigned)(cy + 1) < 6u && (unsigned)(cz + 1) < 6u)
                        ? ((cx + 1) * 6 + (cy + 1)) * 6 + (cz + 1)
                        : GRP_FAR;
      S.grp_of[i] = (unsigned char)g;
      atomicAdd(&S.cnt[g], 1);
    }
    __syncthreads();
    if (w == 0) {
      int c[7], sum = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { const int g = l * 7 + k; c[k] = g < NGRP ? S.cnt[g] : 0; sum += c[k]; }
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (l >= d) inc += t; }
      int run = inc - sum;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int g = l * 7 + k;
        if (g <= NGRP) { S.gstart[g] = run; S.cnt[g] = run; }
        run += c[k];
      }
    }
    __syncthreads();
    for (int i = tid; i < np; i += PL_NT) {
      const int r = atomicAdd(&S.cnt[S.grp_of[i]], 1);
      S.order[r] = (unsigned short)i;
      S.rank[i] = (unsigned short)r;
    }
    __syncthreads();
  }

  // ---- (b) per chunk: records (two per thread), then plane sweeps ------------------------------------------------------------
  const int n_fast = S.gstart[GRP_FAR];
  const int gi = l / 3, pi = l - 3 * gi;     // cell slot of the round (10 = idle lanes 30, 31) and x-plane of this lane
  const zpcs::PlaneCoef Lp = {pi == 1 ? -1.0f : 0.5f, pi == 0 ? -1.5f : (pi == 1 ? 2.0f : -0.5f), pi == 0 ? 1.125f : (pi == 1 ? -0.25f : 0.125f),
                              (float)pi};
  const int ra = w >> 1, rh = w & 1, rxlo = pl_xlo(ra), rncx = pl_ncx(ra);   // this warp's region: rncx x-slabs of 3 columns x 6 cells
  float *Pw = S.priv + w * PL_PRIV;
  for (int cb = 0; cb < n_fast; cb += PL_CHUNK) {
    const int ce = min(cb + PL_CHUNK, n_fast);
#pragma unroll 1
    for (int sb = 0; sb < np; sb += PL_CHUNK) {   // blocks of 512 slots; a bin of at most 512 particles has one block and one chunk
      if (sb + cb > 0) {                          // (the first block of the first chunk was loaded at kernel entry)
#pragma unroll
        for (int r = 0; r < PL_RPT; ++r) {
          const int i = sb + tid + r * PL_NT;
          const size_t s = pslot((size_t)p0 + i);
#pragma unroll
          for (int c = 0; c < NCH; ++c) pd[r][c] = i < np ? pars[s + c * TS] : 0.f;
        }
      }
      float d0[PL_RPT][3], mass[PL_RPT], A[PL_RPT][3], a[PL_RPT][3], B[PL_RPT][9], Kd[PL_RPT][9];
      int slot[PL_RPT];
#pragma unroll
      for (int r = 0; r < PL_RPT; ++r) {
        const int i = sb + tid + r * PL_NT;
        const int pos = i < np ? (int)S.rank[i] : INT_MAX;
        slot[r] = (pos >= cb && pos < ce) ? pos - cb : -1;   // strays (pos >= n_fast) and other chunks' particles: not this pass
        particle_record<MODEL>(pd[r], scalar ? scalar + (size_t)p0 + min(i, np - 1) : nullptr, slot[r] >= 0, dx, dx_inv, dt, D_inv, volume,
                               mu, lam, yield_stress, pp, d0[r], mass[r], A[r], a[r], B[r], Kd[r]);
      }
#pragma unroll
      for (int r = 0; r < PL_RPT; ++r)
        if (slot[r] >= 0) zpcs::write_plane_record(S.rec4 + zpcs::prec_at(slot[r]), d0[r], mass[r], A[r], a[r], B[r], Kd[r]);
    }
    __syncthreads();   // records of the chunk complete
    // x-slab by x-slab (the outer regions hold a ring slab next to the nominal one): lanes of one flush instruction must differ in
    // (cy', cz') or in the plane while sharing cx' — two slabs in one round would meet at cx' + i
#pragma unroll 1
    for (int sl = 0; sl < rncx; ++sl) {
      // the slab's cells that have particles in this chunk, in (column, z) order — gstart is stable since the prologue
      const int col = l / 6, zc0 = l - 6 * col;
      const int g0 = ((rxlo + sl) * 6 + 3 * rh + col) * 6 + zc0;
      const bool ne = l < 18 && max(S.gstart[g0], cb) < min(S.gstart[g0 + 1], ce);
      const unsigned m = __ballot_sync(0xffffffffu, ne);
      if (ne) S.wcells[w][__popc(m & lanemask_lt())] = (unsigned char)g0;
      const int n_w = __popc(m);
      __syncwarp();
      // cells per round: spread evenly over the rounds, at most ten (30 lanes)
      const int rounds = (n_w + PL_UNIT - 1) / PL_UNIT;
      const int U = rounds > 0 ? (n_w + rounds - 1) / rounds : 0;
#pragma unroll 1
      for (int r = 0; r < rounds; ++r) {
        const int ci = r * U + gi;
        const bool have = gi < U && ci < n_w;
        const int g = have ? (int)S.wcells[w][ci] : 0;
        const int lo = have ? max(S.gstart[g], cb) - cb : 0, hi = have ? min(S.gstart[g + 1], ce) - cb : 0;
        const int nmax = __reduce_max_sync(0xffffffffu, hi - lo);
        float acc[7][3][3];
#pragma unroll
        for (int ch = 0; ch < 7; ++ch)
#pragma unroll
          for (int j = 0; j < 3; ++j) { acc[ch][j][0] = 0.f; acc[ch][j][1] = 0.f; acc[ch][j][2] = 0.f; }
        zpcs::sweep_plane(S.rec4, lo, hi, nmax, Lp, acc);
        {
          const int c6 = g / 6, zc = g - 6 * c6, cyp = c6 % 6;                    // g = (cx' * 6 + cy') * 6 + cz'
          const int lx = sl + pi, ly0 = cyp - 3 * rh;                             // arena node (cx' + i, cy' + j, cz' + k), region-local
          // one (j, k) at a time: inside a step every lane adds the same offset to its own (cell, plane), so the addresses differ;
          // ACROSS steps two lanes do meet (cell z with k = 1 and cell z + 1 with k = 0 are the same node), and nothing but the warp
          // barrier keeps the compiler from batching the loads of one step with the stores of another (compute-sanitizer racecheck
          // flagged exactly that in the first version, profiles/r02_sanitizer.md)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (have) {
                float *dstn = Pw + priv_idx(lx, ly0 + j, zc + k);
#pragma unroll
                for (int ch = 0; ch < 7; ++ch) dstn[ch * 160] += acc[ch][j][k];
              }
              __syncwarp();
            }
        }
      }
    }
    __syncthreads();  // records are overwritten by the next chunk / the private copies are read by the merge
  }

  // ---- (c) sum the private copies into eight [7][64] grid tiles (over the dead records), then add them to the grid: TMA bulk
  // reductions.  One float4 = four z-neighbours of one channel; a node is covered by up to three x-slabs and two y-halves.
  float4 *T4 = S.rec4;
  const float4 *P4 = reinterpret_cast<const float4 *>(S.priv);
  for (int o = tid; o < 8 * 112; o += PL_NT) {
    const int b = o / 112, r = o - 112 * b, ch = r >> 4, c4 = r & 15;
    const int X = ((b >> 2) << 2) | (c4 >> 2), Y = (((b >> 1) & 1) << 2) | (c4 & 3), zh = b & 1;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int lx = X - pl_xlo(a);
      if ((unsigned)lx < (unsigned)(pl_ncx(a) + 2)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ly = Y - 3 * h;
          if ((unsigned)ly < 5u) {
            const float4 v = P4[(2 * a + h) * (PL_PRIV / 4) + ch * 40 + lx * 10 + ly * 2 + ((zh + ly) & 1)];
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
          }
        }
      }
    }
    T4[o] = sum;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
  __syncthreads();
  bool remote = false;
  if (tid < 8) {
    const int id = S.tile_id[tid];
    if (id >= 0) {
      float *g = tiles + (size_t)id * 448;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g),
                   "r"(smem_u32(reinterpret_cast<float *>(T4) + tid * 448)), "r"(1792)
                   : "memory");
      remote = halo_send_tile(halo, id, reinterpret_cast<float *>(T4) + tid * 448);
    } else if (status) {  // an arena block the partition does not hold: flag it if any mass was headed there
      bool any = false;
      for (int c = 0; c < 64; ++c) any |= reinterpret_cast<float *>(T4)[tid * 448 + c] != 0.f;
      if (any) atomicOr(status, ZPC_BINS_STENCIL_BLOCK_MISSING);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }

  // ---- (d) far strays (moved more than one cell since the re-bin): per-particle scatter with REDs -----------------------------
  for (int t = n_fast + tid; t < np; t += PL_NT) {
    const size_t s = pslot((size_t)p0 + gorder[t]);
    float pos[3], vel[3], C[9], F[9];
    const float mass = pars[s + ZPC_PB_M * TS];
#pragma unroll
    for (int d = 0; d < 3; ++d) { pos[d] = pars[s + (ZPC_PB_X + d) * TS]; vel[d] = pars[s + (ZPC_PB_V + d) * TS]; }
#pragma unroll
    for (int d = 0; d < 9; ++d) { C[d] = pars[s + (ZPC_PB_C + d) * TS]; F[d] = pars[s + (ZPC_PB_F + d) * TS]; }
    if constexpr (MODEL == 1) zpcp::p2g_scatter_particle_vm(pos, vel, mass, C, F, zpcp::LegacyGrid{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}, tiles, 7, dx, dt, volume, mu, lam, yield_stress);
    else if constexpr (MODEL == 4) {
      zpcp::p2g_scatter_particle_eos(pos, vel, mass, C, scalar[(size_t)p0 + gorder[t]], zpcp::LegacyGrid{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}, tiles, 7, dx, dt, volume, pp.a, pp.b);
    } else if constexpr (MODEL >= 2) {
      float *lj = scalar + (size_t)p0 + gorder[t];
      float logJp = *lj, contrib[9];
      if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, pp.a, pp.b, pp.c, pp.flag != 0, logJp, F, contrib);
      else zpcm::stress_nacc(volume, mu, pp.a, pp.b, pp.c, pp.d, pp.flag != 0, logJp, F, contrib);
      *lj = logJp;
#pragma unroll
      for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
      zpcp::p2g_scatter_core(pos, vel, mass, C, contrib, zpcp::LegacyGrid{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}, tiles, 7, dx);
    } else zpcp::p2g_scatter_particle(pos, vel, mass, C, F, zpcp::LegacyGrid{tb, status, ZPC_BINS_STENCIL_BLOCK_MISSING}, tiles, 7, dx, dt, volume, mu, lam);
  }
  if (tid < 8) {
    // smem must outlive the bulk reads; a tile that also went to a peer waits for the writes themselves, so that the kernel's
    // completion (and the barrier the caller puts after it) orders them before the peer's grid update
    if (remote) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// One particle of the binned G2P: gather against the staged arena velocities sv (= G2PSmem::v), APIC C, advect pos.
template <bool FAST_DIV = false, class BG = BinGridLegacy>
__device__ __forceinline__ void g2p_arena_particle(const float *sv, int kx, int ky, int kz, const BG &bg,
                                                   const float *__restrict__ tiles, int nch, float dx, float dt, float D_inv,
                                                   float (&pos)[3], float (&vel)[3], float (&C)[9], int *status = nullptr,
                                                   float dx_inv = 0.f) {
  zpcm::Arena ar;
  zpcm::arena_init<FAST_DIV>(ar, dx, pos, dx_inv);   // the staged kernel divides through 1/dx and two FMAs (same bits, no slow-path branch)
  const int ax0 = ar.corner[0] - 4 * kx, ay0 = ar.corner[1] - 4 * ky, az0 = ar.corner[2] - 4 * kz;
  float G[9];  // G[r + 3e] = sum W v_r o_e
  if ((unsigned)ax0 < 6u && (unsigned)ay0 < 6u && (unsigned)az0 < 6u) {
    int fo[3], go[3], ho[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const int a = ax0 + o, b = ay0 + o, c = az0 + o;
      fo[o] = (a >> 2) * (4 * 192) + ((a & 3) << 4);
      go[o] = (b >> 2) * (2 * 192) + ((b & 3) << 2);
      ho[o] = (c >> 2) * 192 + (c & 3);
    }
    // separable contraction: z, then y, then x
    float px[3][3], pz[3][3], py[3][3];  // [i][r]
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
#pragma unroll
      for (int r = 0; r < 3; ++r) { px[ii][r] = 0.f; py[ii][r] = 0.f; pz[ii][r] = 0.f; }
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        const int base = fo[ii] + go[jj];
        float u[3], uz[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float v0 = sv[base + ho[0] + r * 64], v1 = sv[base + ho[1] + r * 64], v2 = sv[base + ho[2] + r * 64];
          const float t1 = ar.w[2][1] * v1, t2 = ar.w[2][2] * v2;
          u[r] = fmaf(ar.w[2][0], v0, t1 + t2);
          uz[r] = fmaf(2.f, t2, t1);
        }
        const float wyj = ar.w[1][jj];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float t = wyj * u[r];
          px[ii][r] += t;
          py[ii][r] = fmaf((float)jj, t, py[ii][r]);
          pz[ii][r] = fmaf(wyj, uz[r], pz[ii][r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float t0 = ar.w[0][0] * px[0][r], t1 = ar.w[0][1] * px[1][r], t2 = ar.w[0][2] * px[2][r];
      vel[r] = t0 + t1 + t2;
      G[r] = fmaf(2.f, t2, t1);
      G[r + 3] = ar.w[0][0] * py[0][r] + ar.w[0][1] * py[1][r] + ar.w[0][2] * py[2][r];
      G[r + 6] = ar.w[0][0] * pz[0][r] + ar.w[0][1] * pz[1][r] + ar.w[0][2] * pz[2][r];
    }
  } else {
    // the out-of-arena path is a call: it gets copies, so that the arena path keeps ar / vel / G in registers (passing them by
    // reference put 27 local-memory stores per particle on the common path)
    zpcm::Arena ar2 = ar;
    float vel2[3], G2[9];
    zpcp::g2p_gather_particle(ar2, bg.accessor(status), tiles, nch, vel2, G2);
#pragma unroll
    for (int d = 0; d < 3; ++d) vel[d] = vel2[d];
#pragma unroll
    for (int d = 0; d < 9; ++d) G[d] = G2[d];
  }
  // C[r + 3e] = D_inv * sum W v_r (o_e dx - local_e) = D_inv * (dx G_re - local_e v_r)
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r + 3 * e] = (dx * G[r + 3 * e] - ar.local[e] * vel[r]) * D_inv;
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] += vel[d] * dt;
}

#ifndef ZPC_G2P_NT
#define ZPC_G2P_NT 256
#endif
#ifndef ZPC_G2P_MINB
#define ZPC_G2P_MINB 4
#endif
constexpr int G2P_NT = ZPC_G2P_NT;
struct G2PSmem {
  float v[8][3][64];  // 6144 B: channels 1..3 of the eight arena tiles
  unsigned long long bar;
  int tile_id[8];
  int cnt[NGRP + 3];               // cell-order cache for the next P2G
  unsigned char grp_of[BIN_MAX];
};

// EOS = true: G2PTransfer with EquationOfStateConfig (G2P.hpp:69-73) — J (side array `scalar`, bin order) <- (1 + tr(C) dt) J, the F
// channels are neither read nor written
template <bool EOS = false>
__global__ void __launch_bounds__(G2P_NT, ZPC_G2P_MINB)
g2p_binned_kernel(float *__restrict__ pars, const int *__restrict__ binStart, const int *__restrict__ binKey,
                  const int *__restrict__ numBins, unsigned short *__restrict__ cellOrder, unsigned short *__restrict__ cellStart,
                  zpc_hashtable_view tb, const float *__restrict__ tiles, int nch, float dx, float dt, float *__restrict__ scalar,
                  int *__restrict__ status) {
  __shared__ __align__(128) G2PSmem S;
  const int bin = blockIdx.x;
  if (bin >= *numBins) return;
  const int tid = threadIdx.x;
  const int p0 = binStart[bin], np = min(binStart[bin + 1] - p0, BIN_MAX);
  const int kx = binKey[3 * bin], ky = binKey[3 * bin + 1], kz = binKey[3 * bin + 2];
  if (tid < 8) S.tile_id[tid] = zpcm::table_query(kx + (tid >> 2), ky + ((tid >> 1) & 1), kz + (tid & 1), tb.tableSize, tb.keys, tb.indices);
  for (int i = tid; i < NGRP + 3; i += G2P_NT) S.cnt[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(&S.bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    unsigned bytes = 0;
    for (int b = 0; b < 8; ++b) bytes += S.tile_id[b] >= 0 ? 768u : 0u;
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.bar)), "r"(bytes) : "memory");
    for (int b = 0; b < 8; ++b)
      if (S.tile_id[b] >= 0) {
        const float *g = tiles + ((size_t)S.tile_id[b] * nch + 1) * 64;
        asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&S.v[b][0][0])),
                     "l"(g), "r"(768), "r"(smem_u32(&S.bar))
                     : "memory");
      }
  }
  // blocks missing from the partition read as zero velocity
  for (int b = 0; b < 8; ++b)
    if (S.tile_id[b] < 0)
      for (int i = tid; i < 192; i += G2P_NT) (&S.v[b][0][0])[i] = 0.f;
  {  // wait for the TMA bytes (phase 0)
    unsigned done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                   : "=r"(done)
                   : "r"(smem_u32(&S.bar)), "r"(0)
                   : "memory");
    }
  }
  __syncthreads();
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  const float *sv = &S.v[0][0][0];
  for (int i = tid; i < np; i += G2P_NT) {
    const size_t s = pslot((size_t)p0 + i);
    float pos[3], Fo[9];
#pragma unroll
    for (int d = 0; d < 3; ++d) pos[d] = pars[s + (ZPC_PB_X + d) * TS];
    if constexpr (!EOS) {
#pragma unroll
      for (int d = 0; d < 9; ++d) Fo[d] = pars[s + (ZPC_PB_F + d) * TS];  // issued early: consumed after the contraction
    }
    float vel[3], C[9], tmp[9];
    g2p_arena_particle(sv, kx, ky, kz, BinGridLegacy{tb}, tiles, nch, dx, dt, D_inv, pos, vel, C, status);
    if (cellOrder) {  // group of the NEW home cell, exactly as the binned P2G computes it from the stored position
      const int cx = (int)floorf(pos[0] / dx - 0.5f) - 1 - 4 * kx, cy = (int)floorf(pos[1] / dx - 0.5f) - 1 - 4 * ky,
                cz = (int)floorf(pos[2] / dx - 0.5f) - 1 - 4 * kz;
      const int g = ((unsigned)(cx + 1) < 6u && (unsigned)(cy + 1) < 6u && (unsigned)(cz + 1) < 6u)
                        ? ((cx + 1) * 6 + (cy + 1)) * 6 + (cz + 1)
                        : GRP_FAR;
      S.grp_of[i] = (unsigned char)g;
      atomicAdd(&S.cnt[g], 1);
    }
    if constexpr (EOS) {
      scalar[(size_t)p0 + i] = (1 + (C[0] + C[4] + C[8]) * dt) * scalar[(size_t)p0 + i];
    } else {
#pragma unroll
      for (int d = 0; d < 9; ++d) tmp[d] = C[d] * dt + ((d & 3) ? 0.f : 1.f);
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
          pars[s + (ZPC_PB_F + 3 * c + r) * TS] = tmp[r] * Fo[3 * c] + tmp[3 + r] * Fo[3 * c + 1] + tmp[6 + r] * Fo[3 * c + 2];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) { pars[s + (ZPC_PB_X + d) * TS] = pos[d]; pars[s + (ZPC_PB_V + d) * TS] = vel[d]; }
#pragma unroll
    for (int d = 0; d < 9; ++d) pars[s + (ZPC_PB_C + d) * TS] = C[d];
  }
  if (cellOrder) {  // counting sort of the bin by new cell group -> global cache for the next P2G
    __syncthreads();
    const int w = tid >> 5, l = tid & 31;
    if (w == 0) {
      int c[7], sum = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { const int g = l * 7 + k; c[k] = g < NGRP ? S.cnt[g] : 0; sum += c[k]; }
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (l >= d) inc += t; }
      int run = inc - sum;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int g = l * 7 + k;
        if (g <= NGRP) {
          S.cnt[g] = run;
          cellStart[(size_t)bin * ZPCB200_CELL_GROUPS_PAD + g] = (unsigned short)run;
        }
        run += c[k];
      }
    }
    __syncthreads();
    for (int i = tid; i < np; i += G2P_NT) cellOrder[p0 + atomicAdd(&S.cnt[S.grp_of[i]], 1)] = (unsigned short)i;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Binned G2P, staged variant: the particle channels G2P reads (x: 3 x 128 B, F: 9 x 128 B per 32-particle tile, both
// contiguous inside a TileVector tile) are brought in with TMA bulk copies, eight tiles (256 particles) per stage, two
// stages.  One thread issues the copies of the first two stages before anything else happens, so a whole bin's reads
// are in flight while the arena blocks are looked up and staged — the memory-level parallelism no longer depends on
// how many loads a thread can hold in registers.  Threads map to (tile, lane) of the stage: every global store is a
// full 128-byte line.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// A transaction that never completes (it cannot, short of a wrong byte count) must neither hang the GPU nor kill the context: after
// 2^20 timed-out try_waits the thread raises ZPC_BINS_TMA_TIMEOUT in the bins' status word and goes on with whatever the stage holds.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity, int *status) {
  unsigned done = 0, spins = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1u << 20)) {
      if (status) atomicOr(status, ZPC_BINS_TMA_TIMEOUT);
      break;
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NT> struct G2PStagedSmem {
  static constexpr int G2P_ST = NT / 32;  // particle tiles per stage (one thread per particle)
  float v[8][3][64];               // 6144 B: channels 1..3 of the eight arena tiles
  float xs[2][G2P_ST][3][TS];      // 2 x 3072 B
  float fs[2][G2P_ST][9][TS];      // 2 x 9216 B
  unsigned long long bar_grid, bar_stage[2];
  int tile_id[8];
  int cnt[NGRP + 3];
  unsigned char grp_of[BIN_MAX];
};
static_assert(sizeof(G2PStagedSmem<256>) <= 48 * 1024, "static shared memory");

// NT threads per CTA: small CTAs interleave their prologue / wait / store phases better — measured at C3:
// 64 threads (15 CTAs/SM) 2.18 ms, 128 (8/SM) 2.30 ms, 256 (4/SM) 2.98 ms.
template <int NT, bool EOS = false, class BG = BinGridLegacy>
__global__ void __launch_bounds__(NT, 1024 / NT)
g2p_binned_staged_kernel(float *__restrict__ pars, const int *__restrict__ binStart, const int *__restrict__ binKey,
                         const int *__restrict__ numBins, unsigned short *__restrict__ cellOrder,
                         unsigned short *__restrict__ cellStart, BG bg, const float *__restrict__ tiles, int nch,
                         float dx, float dt, float *__restrict__ scalar, int *__restrict__ status) {
  constexpr int G2P_ST = NT / 32;
  __shared__ __align__(128) G2PStagedSmem<NT> S;
  const int bin = blockIdx.x;
  if (bin >= *numBins) return;
  const int tid = threadIdx.x;
  const int p0 = binStart[bin], np = min(binStart[bin + 1] - p0, BIN_MAX);
  const int t0 = p0 >> 5;                                            // first particle tile the bin overlaps
  const int ntiles = np > 0 ? ((p0 + np - 1) >> 5) - t0 + 1 : 0;     // <= 33
  const int nstages = (ntiles + G2P_ST - 1) / G2P_ST;
  if (tid == 0) {
    mbar_init(&S.bar_grid, 1);
    mbar_init(&S.bar_stage[0], 1);
    mbar_init(&S.bar_stage[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage c <- tiles [t0 + G2P_ST c, t0 + G2P_ST (c + 1)) ∩ the bin's tiles: x = channels 1..3, F = channels 16..24
  auto issue_stage = [&](int c) {
    const int b = c & 1, tb0 = c * G2P_ST, nt = min(G2P_ST, ntiles - tb0);
    mbar_expect_tx(&S.bar_stage[b], (unsigned)nt * (EOS ? 3u : 3u + 9u) * TS * 4u);
    for (int t = 0; t < nt; ++t) {
      const float *src = pars + (size_t)(t0 + tb0 + t) * NCH * TS;
      bulk_g2s(&S.xs[b][t][0][0], src + ZPC_PB_X * TS, 3 * TS * 4, &S.bar_stage[b]);
      if constexpr (!EOS) bulk_g2s(&S.fs[b][t][0][0], src + ZPC_PB_F * TS, 9 * TS * 4, &S.bar_stage[b]);
    }
  };
  if (tid == 0) {  // the same thread initialised the barriers: program order suffices
    if (nstages > 0) issue_stage(0);
    if (nstages > 1) issue_stage(1);
  }
  const int kx = binKey[3 * bin], ky = binKey[3 * bin + 1], kz = binKey[3 * bin + 2];
  if (tid < 8) S.tile_id[tid] = bg.query(kx + (tid >> 2), ky + ((tid >> 1) & 1), kz + (tid & 1));
  for (int i = tid; i < NGRP + 3; i += NT) S.cnt[i] = 0;
  __syncthreads();
  if constexpr (BG::SPARSE) {
    // SparseGrid: the three velocity channels of the eight octants, 8 x 3 x 16 rows of four z-cells, one 128-bit load each
    for (int q = tid; q < 8 * 48; q += NT) {
      const int b = q / 48, r = q - 48 * b, c = r >> 4, row = r & 15;
      const int id = S.tile_id[b];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // blocks missing from the partition read as zero velocity
      if (id >= 0) v = *reinterpret_cast<const float4 *>(tiles + BG::row_offset(id, nch, 1 + c, row >> 2, row & 3));
      *reinterpret_cast<float4 *>(&S.v[b][c][row * 4]) = v;
    }
  } else {
    if (tid == 0) {
      unsigned bytes = 0;
      for (int b = 0; b < 8; ++b) bytes += S.tile_id[b] >= 0 ? 768u : 0u;
      mbar_expect_tx(&S.bar_grid, bytes);
      for (int b = 0; b < 8; ++b)
        if (S.tile_id[b] >= 0) bulk_g2s(&S.v[b][0][0], tiles + ((size_t)S.tile_id[b] * nch + 1) * 64, 768, &S.bar_grid);
    }
    for (int b = 0; b < 8; ++b)  // blocks missing from the partition read as zero velocity
      if (S.tile_id[b] < 0)
        for (int i = tid; i < 192; i += NT) (&S.v[b][0][0])[i] = 0.f;
    mbar_wait(&S.bar_grid, 0, status);
  }
  __syncthreads();
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  const float *sv = &S.v[0][0][0];
  const int tl = tid >> 5, ln = tid & 31;
  for (int c = 0; c < nstages; ++c) {
    const int b = c & 1;
    mbar_wait(&S.bar_stage[b], (unsigned)(c >> 1) & 1u, status);
    const int gp = ((t0 + c * G2P_ST + tl) << 5) + ln;  // global particle slot of this thread
    const int i = gp - p0;                              // slot relative to the bin
    const bool mine = i >= 0 && i < np;
    float pos[3], Fo[9];
    if (mine) {
#pragma unroll
      for (int d = 0; d < 3; ++d) pos[d] = S.xs[b][tl][d][ln];
      if constexpr (!EOS) {
#pragma unroll
        for (int d = 0; d < 9; ++d) Fo[d] = S.fs[b][tl][d][ln];
      }
    }
    if (c + 2 < nstages) {  // bins with more than 512 particles: refill this stage once everybody has read it
      __syncthreads();
      if (tid == 0) issue_stage(c + 2);
    }
    if (mine) {
      const size_t s = pslot((size_t)gp);
      float vel[3], C[9], tmp[9];
      g2p_arena_particle<true>(sv, kx, ky, kz, bg, tiles, nch, dx, dt, D_inv, pos, vel, C, status, dx_inv);
      if (cellOrder) {  // group of the NEW home cell, exactly as the binned P2G computes it from the stored position
        const int cx = (int)floorf(zpcm::div_exact(pos[0], dx, dx_inv) - 0.5f) - 1 - 4 * kx,
                  cy = (int)floorf(zpcm::div_exact(pos[1], dx, dx_inv) - 0.5f) - 1 - 4 * ky,
                  cz = (int)floorf(zpcm::div_exact(pos[2], dx, dx_inv) - 0.5f) - 1 - 4 * kz;
        const int g = ((unsigned)(cx + 1) < 6u && (unsigned)(cy + 1) < 6u && (unsigned)(cz + 1) < 6u)
                          ? ((cx + 1) * 6 + (cy + 1)) * 6 + (cz + 1)
                          : GRP_FAR;
        S.grp_of[i] = (unsigned char)g;
        atomicAdd(&S.cnt[g], 1);
      }
      if constexpr (EOS) {
        scalar[gp] = (1 + (C[0] + C[4] + C[8]) * dt) * scalar[gp];
      } else {
#pragma unroll
        for (int d = 0; d < 9; ++d) tmp[d] = C[d] * dt + ((d & 3) ? 0.f : 1.f);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int r = 0; r < 3; ++r)
            pars[s + (ZPC_PB_F + 3 * cc + r) * TS] = tmp[r] * Fo[3 * cc] + tmp[3 + r] * Fo[3 * cc + 1] + tmp[6 + r] * Fo[3 * cc + 2];
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) { pars[s + (ZPC_PB_X + d) * TS] = pos[d]; pars[s + (ZPC_PB_V + d) * TS] = vel[d]; }
#pragma unroll
      for (int d = 0; d < 9; ++d) pars[s + (ZPC_PB_C + d) * TS] = C[d];
    }
  }
  if (cellOrder) {  // counting sort of the bin by new cell group -> global cache for the next P2G
    __syncthreads();
    const int w = tid >> 5, l = tid & 31;
    if (w == 0) {
      int cn[7], sum = 0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { const int g = l * 7 + k; cn[k] = g < NGRP ? S.cnt[g] : 0; sum += cn[k]; }
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (l >= d) inc += t; }
      int run = inc - sum;
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int g = l * 7 + k;
        if (g <= NGRP) {
          S.cnt[g] = run;
          cellStart[(size_t)bin * ZPCB200_CELL_GROUPS_PAD + g] = (unsigned short)run;
        }
        run += cn[k];
      }
    }
    __syncthreads();
    for (int i = tid; i < np; i += NT) cellOrder[p0 + atomicAdd(&S.cnt[S.grp_of[i]], 1)] = (unsigned short)i;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// binning
// ----------------------------------------------------------------------------------------------------------------
// key = (block rank << 6) | cell id of the home cell ; val = particle index
template <bool AOSOA, class BG>
__global__ void bin_keys_kernel(const float *__restrict__ X, size_t n, float dxinv, BG bg, unsigned *keys,
                                int *vals, int *err, int cap) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[3];
  if (AOSOA) {
    const size_t s = pslot(i);
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = X[s + (ZPC_PB_X + d) * TS];
  } else {
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = X[3 * i + d];
  }
  const int c0 = zpcm::sparsity_coord(x[0], dxinv), c1 = zpcm::sparsity_coord(x[1], dxinv), c2 = zpcm::sparsity_coord(x[2], dxinv);
  int b = bg.query(c0 >> 2, c1 >> 2, c2 >> 2);
  if (b < 0) { if (err) atomicOr(err, ZPC_BINS_HOME_BLOCK_MISSING); b = 0; }
  if (b >= cap) { if (err) atomicOr(err, ZPC_BINS_BLOCK_CAPACITY); b = cap - 1; }   // keeps start[] / end[] in bounds and the sort's ebit honest
  keys[i] = ((unsigned)b << 6) | (unsigned)(((c0 & 3) << 4) | ((c1 & 3) << 2) | (c2 & 3));
  vals[i] = (int)i;
}
__global__ void bin_bounds_kernel(const unsigned *__restrict__ keys, size_t n, int *start, int *end) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned b = keys[i] >> 6;
  if (i == 0 || (keys[i - 1] >> 6) != b) start[b] = (int)i;
  if (i == n - 1 || (keys[i + 1] >> 6) != b) end[b] = (int)i + 1;
}
template <class BG>
__global__ void bin_count_kernel(const int *start, const int *end, BG bg, int cap, int *nbins, int *err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int cnt = bg.count();
  if (b == 0 && err && cnt > cap) atomicOr(err, ZPC_BINS_BLOCK_CAPACITY);
  if (b >= cap) return;
  nbins[b] = b < cnt ? (end[b] - start[b] + BIN_MAX - 1) / BIN_MAX : 0;
}
template <class BG>
__global__ void bin_fill_kernel(const int *start, const int *end, const int *nbins, const int *binoff, BG bg, int cap,
                                int n, int *binStart, int *binKey, int *numBins, int binCapacity,
                                int *err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = min(bg.count(), cap);
  if (b == 0) {
    const int total = nb > 0 ? binoff[nb - 1] + nbins[nb - 1] : 0;
    if (total > binCapacity) { if (err) atomicOr(err, ZPC_BINS_BIN_CAPACITY); *numBins = 0; }
    else { *numBins = total; binStart[total] = n; }
  }
  if (b >= nb) return;
  const int k = nbins[b], o = binoff[b];
  if (o + k > binCapacity) return;
  int kx = 0, ky = 0, kz = 0;
  if (k > 0) bg.key_of(b, kx, ky, kz);
  for (int j = 0; j < k; ++j) {
    binStart[o + j] = start[b] + j * BIN_MAX;
    binKey[3 * (o + j)] = kx;
    binKey[3 * (o + j) + 1] = ky;
    binKey[3 * (o + j) + 2] = kz;
  }
}
// dst slot i <- src particle perm[i]
template <bool SRC_AOSOA>
__global__ void bin_gather_kernel(zpc_particles_view A, const float *__restrict__ srcT, const int *__restrict__ perm, size_t n,
                                  float *__restrict__ dstT) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t j = (size_t)perm[i], d = pslot(i);
  if (SRC_AOSOA) {
    const size_t s = pslot(j);
#pragma unroll
    for (int c = 0; c < NCH; ++c) dstT[d + c * TS] = srcT[s + c * TS];
  } else {
    dstT[d + ZPC_PB_M * TS] = A.M[j];
#pragma unroll
    for (int c = 0; c < 3; ++c) { dstT[d + (ZPC_PB_X + c) * TS] = A.X[3 * j + c]; dstT[d + (ZPC_PB_V + c) * TS] = A.V[3 * j + c]; }
#pragma unroll
    for (int c = 0; c < 9; ++c) { dstT[d + (ZPC_PB_C + c) * TS] = A.C[9 * j + c]; dstT[d + (ZPC_PB_F + c) * TS] = A.F[9 * j + c]; }
  }
}
__global__ void unbin_kernel(const float *__restrict__ srcT, size_t n, zpc_particles_view A) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t s = pslot(i);
  if (A.M) A.M[i] = srcT[s + ZPC_PB_M * TS];
#pragma unroll
  for (int c = 0; c < 3; ++c) { A.X[3 * i + c] = srcT[s + (ZPC_PB_X + c) * TS]; A.V[3 * i + c] = srcT[s + (ZPC_PB_V + c) * TS]; }
#pragma unroll
  for (int c = 0; c < 9; ++c) { A.C[9 * i + c] = srcT[s + (ZPC_PB_C + c) * TS]; A.F[9 * i + c] = srcT[s + (ZPC_PB_F + c) * TS]; }
}

__global__ void gather_f32_kernel(const float *__restrict__ src, const int *__restrict__ idx, float *__restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

int bit_length(unsigned v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

// shared pipeline of bin_particles / rebin_particles
template <bool SRC_AOSOA, class BG = BinGridLegacy>
int bin_pipeline(void *temp, size_t *temp_bytes, zpc_particles_view A, const float *srcT, size_t n, BG tb, float dx,
                 zpc_bins_view dst, int *order_out, cudaStream_t s) {
  if (!temp_bytes) return ZPCB200_E_BADARG;
  if (dst.pars.numChannels != NCH || dst.binCapacity <= 0) return ZPCB200_E_BADARG;
  if (n > (size_t)INT_MAX) return ZPCB200_E_UNSUPPORTED;
  const int cap = dst.binCapacity;                          // also the bound on the number of blocks
  const int ebit = 6 + bit_length((unsigned)(cap - 1));
  if (ebit > 32) return ZPCB200_E_UNSUPPORTED;
  size_t sort_bytes = 0, scan_bytes = 0;
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = zpcb200_radix_sort_pair_u32(nullptr, &sort_bytes, none, none, none, none, n, 0, ebit, nullptr);
  if (rc) return rc;
  rc = zpcb200_exclusive_scan_sum_i32(nullptr, &scan_bytes, none, none, (size_t)cap, nullptr);
  if (rc) return rc;
  const size_t o_keys = 256, o_vals = zpc_align_up(o_keys + 4 * n, 256), o_skeys = zpc_align_up(o_vals + 4 * n, 256),
               o_svals = zpc_align_up(o_skeys + 4 * n, 256), o_start = zpc_align_up(o_svals + 4 * n, 256),
               o_end = o_start + zpc_align_up(4 * (size_t)cap, 256), o_nb = o_end + zpc_align_up(4 * (size_t)cap, 256),
               o_off = o_nb + zpc_align_up(4 * (size_t)cap, 256), o_scan = o_off + zpc_align_up(4 * (size_t)cap, 256),
               o_sort = zpc_align_up(o_scan + scan_bytes, 256), need = o_sort + sort_bytes;
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  char *t = (char *)temp;
  int *err = dst.status;   // may be NULL
  unsigned *keys = (unsigned *)(t + o_keys), *skeys = (unsigned *)(t + o_skeys);
  int *vals = (int *)(t + o_vals), *svals = order_out ? order_out : (int *)(t + o_svals);
  int *start = (int *)(t + o_start), *end = (int *)(t + o_end), *nbins = (int *)(t + o_nb), *binoff = (int *)(t + o_off);
  ZPC_CUDA(cudaMemsetAsync(t, 0, 256, s));
  ZPC_CUDA(cudaMemsetAsync(start, 0, o_off - o_start, s));
  if (dst.cellOrderValid) ZPC_CUDA(cudaMemsetAsync(dst.cellOrderValid, 0, sizeof(int), s));  // new slots: cache is stale
  const unsigned gp = (unsigned)((n + 255) / 256), gb = (unsigned)((cap + 255) / 256);
  if (n) {
    bin_keys_kernel<SRC_AOSOA, BG><<<gp, 256, 0, s>>>(SRC_AOSOA ? srcT : A.X, n, 1.0f / dx, tb, keys, vals, err, cap);
    ZPC_CHECK_LAUNCH();
    zpc_port pk = {keys, 0, 0, 0, 1}, pv = {vals, 0, 0, 0, 1}, psk = {skeys, 0, 0, 0, 1}, psv = {svals, 0, 0, 0, 1};
    size_t sb = sort_bytes;
    rc = zpcb200_radix_sort_pair_u32(t + o_sort, &sb, pk, pv, psk, psv, n, 0, ebit, s);
    if (rc) return rc;
    bin_bounds_kernel<<<gp, 256, 0, s>>>(skeys, n, start, end);
    ZPC_CHECK_LAUNCH();
  }
  bin_count_kernel<BG><<<gb, 256, 0, s>>>(start, end, tb, cap, nbins, err);
  ZPC_CHECK_LAUNCH();
  {
    zpc_port pi = {nbins, 0, 0, 0, 1}, po = {binoff, 0, 0, 0, 1};
    size_t sb = scan_bytes;
    rc = zpcb200_exclusive_scan_sum_i32(t + o_scan, &sb, pi, po, (size_t)cap, s);
    if (rc) return rc;
  }
  bin_fill_kernel<BG><<<gb, 256, 0, s>>>(start, end, nbins, binoff, tb, cap, (int)n, dst.binStart, dst.binKey,
                                          dst.numBins, dst.binCapacity, err);
  ZPC_CHECK_LAUNCH();
  if (n) {
    bin_gather_kernel<SRC_AOSOA><<<gp, 256, 0, s>>>(A, srcT, svals, n, dst.pars.base);
    ZPC_CHECK_LAUNCH();
  }
  return ZPCB200_OK;
}

// Kernel variants (see zpcb200_set_tuning): defaults from the environment, once.
struct Tuning {
  int p2g_sweep;   // 4 = three cells x nine node columns per warp (default: converged Jacobi sweeps skipped, next chunk's records computed
                   // while the other warps finish their sweeps); 5 = the same on packed fp32 (FFMA2); 3 = one cell x 27 nodes, the
                   // reference's four Jacobi sweeps always; 6 = plane sweep kernel (atomic-free private regions)
  int plane_prefetch;  // plane sweep: distance (in bins) of the L2 prefetch of a later CTA's particle lines, 0 = off (env ZPCB200_PLANE_PREFETCH)
  int g2p_staged;  // 0 = plain loads, 256-thread CTAs; 1 (= 64) | 64 | 128 | 256 = particle channels staged with TMA bulk copies, that many threads per CTA
};
Tuning &tuning() {
  static Tuning t = [] {
    Tuning d = {4, 448, 1};
    if (const char *e = getenv("ZPCB200_PLANE_PREFETCH")) d.plane_prefetch = atoi(e);
    if (const char *e = getenv("ZPCB200_P2G_SWEEP")) d.p2g_sweep = (e[0] >= '3' && e[0] <= '6') ? e[0] - '0' : 4;
    if (const char *e = getenv("ZPCB200_G2P_STAGED")) d.g2p_staged = atoi(e);
    return d;
  }();
  return t;
}

}  // namespace

// shared launch of the binned P2G: MODEL 0 fixed-corotated, 1 von Mises
template <int MODEL>
static int p2g_binned_launch(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt, float volume, float E, float nu,
                             float yield_stress, zpc_stream_t stream, float *scalar = nullptr, zpcm::PlasticPrm pp = {},
                             zpc_halo_view halo = {}) {
  if (g.numChannels != 7 || bins.pars.numChannels != NCH || !bins.binStart || !bins.binKey || !bins.numBins)
    return ZPCB200_E_BADARG;
  static std::atomic<bool> attr_set{false};  // idempotent: two threads racing here both set the same attribute
  if (!attr_set.load(std::memory_order_acquire)) {
    ZPC_CUDA(cudaFuncSetAttribute(p2g_binned_kernel<3, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2GSmem)));
    ZPC_CUDA(cudaFuncSetAttribute(p2g_binned_kernel<4, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2GSmem)));
    ZPC_CUDA(cudaFuncSetAttribute(p2g_binned_kernel<5, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2GSmem)));
    ZPC_CUDA(cudaFuncSetAttribute(p2g_plane_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2GPlaneSmem)));
    attr_set.store(true, std::memory_order_release);
  }
  const int variant = tuning().p2g_sweep;
  float mu, lam;
  zpcm::lame_host(E, nu, mu, lam);
  const bool cache = bins.cellOrder && bins.cellStart && bins.cellOrderValid;
  if (variant == 6) {
    p2g_plane_kernel<MODEL><<<bins.binCapacity, PL_NT, sizeof(P2GPlaneSmem), (cudaStream_t)stream>>>(
        bins.pars.base, bins.binStart, bins.binKey, bins.numBins, cache ? bins.cellOrder : nullptr, bins.cellStart, bins.cellOrderValid, tb,
        g.tiles, g.dx, dt, volume, mu, lam, yield_stress, scalar, pp, tuning().plane_prefetch, bins.status, halo);
    ZPC_CHECK_LAUNCH();
    return ZPCB200_OK;
  }
  auto kern = variant == 3   ? p2g_binned_kernel<3, MODEL>
              : variant == 5 ? p2g_binned_kernel<5, MODEL>
                             : p2g_binned_kernel<4, MODEL>;
  kern<<<bins.binCapacity, P2G_NT, sizeof(P2GSmem), (cudaStream_t)stream>>>(
      bins.pars.base, bins.binStart, bins.binKey, bins.numBins, cache ? bins.cellOrder : nullptr, bins.cellStart,
      bins.cellOrderValid, BinGridLegacy{tb}, g.tiles, g.dx, dt, volume, mu, lam, variant == 3 ? 0 : 1, yield_stress, scalar, pp, bins.status, halo, 7);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

template <bool EOS>
static int g2p_binned_launch(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt, float *scalar, zpc_stream_t stream) {
  if (g.numChannels < 4 || bins.pars.numChannels != NCH || !bins.binStart || !bins.binKey || !bins.numBins)
    return ZPCB200_E_BADARG;
  const bool cache = bins.cellOrder && bins.cellStart && bins.cellOrderValid;
  const int staged = tuning().g2p_staged;
  auto kern = staged == 128 ? g2p_binned_staged_kernel<128, EOS> : staged == 256 ? g2p_binned_staged_kernel<256, EOS> : g2p_binned_staged_kernel<64, EOS>;
  const int nt = staged == 128 ? 128 : staged == 256 ? 256 : staged ? 64 : G2P_NT;
  if (staged)
    kern<<<bins.binCapacity, nt, 0, (cudaStream_t)stream>>>(bins.pars.base, bins.binStart, bins.binKey, bins.numBins,
                                                                cache ? bins.cellOrder : nullptr, bins.cellStart, BinGridLegacy{tb}, g.tiles,
                                                                g.numChannels, g.dx, dt, scalar, bins.status);
  else
    g2p_binned_kernel<EOS><<<bins.binCapacity, nt, 0, (cudaStream_t)stream>>>(bins.pars.base, bins.binStart, bins.binKey, bins.numBins,
                                                                                  cache ? bins.cellOrder : nullptr, bins.cellStart, tb, g.tiles,
                                                                                  g.numChannels, g.dx, dt, scalar, bins.status);
  ZPC_CHECK_LAUNCH();
  if (cache) ZPC_CUDA(cudaMemsetAsync(bins.cellOrderValid, 1, sizeof(int), (cudaStream_t)stream));  // non-zero = valid
  return ZPCB200_OK;
}

extern "C" {

int zpcb200_set_tuning(int p2g_sweep, int g2p_staged) {
  if (((p2g_sweep < 3 || p2g_sweep > 6) && p2g_sweep != -1) || (g2p_staged != -1 && g2p_staged != 0 && g2p_staged != 1 && g2p_staged != 64 && g2p_staged != 128 && g2p_staged != 256))
    return ZPCB200_E_BADARG;
  if (p2g_sweep != -1) tuning().p2g_sweep = p2g_sweep;
  if (g2p_staged != -1) tuning().g2p_staged = g2p_staged;
  return ZPCB200_OK;
}
int zpcb200_get_tuning(int *p2g_sweep, int *g2p_staged) {
  if (p2g_sweep) *p2g_sweep = tuning().p2g_sweep;
  if (g2p_staged) *g2p_staged = tuning().g2p_staged;
  return ZPCB200_OK;
}

int zpcb200_bin_particles(void *temp, size_t *temp_bytes, zpc_particles_view pars, zpc_hashtable_view table, float dx,
                          zpc_bins_view bins, int *order_out, zpc_stream_t stream) {
  if (temp && pars.count && (!pars.X || !pars.V || !pars.M || !pars.C || !pars.F)) return ZPCB200_E_BADARG;
  if (temp && bins.pars.size < pars.count) return ZPCB200_E_BADARG;
  return bin_pipeline<false>(temp, temp_bytes, pars, nullptr, pars.count, BinGridLegacy{table}, dx, bins, order_out, (cudaStream_t)stream);
}
int zpcb200_rebin_particles(void *temp, size_t *temp_bytes, zpc_bins_view src, zpc_hashtable_view table, float dx,
                            zpc_bins_view dst, zpc_stream_t stream) {
  zpc_particles_view none = {};
  if (temp && (src.pars.base == dst.pars.base || dst.pars.size < src.pars.size)) return ZPCB200_E_BADARG;
  return bin_pipeline<true>(temp, temp_bytes, none, src.pars.base, src.pars.size, BinGridLegacy{table}, dx, dst, nullptr, (cudaStream_t)stream);
}
int zpcb200_gather_f32(const float *src, const int *idx, float *dst, size_t n, zpc_stream_t stream) {
  if (n && (!src || !idx || !dst || src == dst)) return ZPCB200_E_BADARG;
  if (!n) return ZPCB200_OK;
  gather_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
int zpcb200_unbin_particles(zpc_bins_view bins, zpc_particles_view pars, zpc_stream_t stream) {
  if (!pars.X || !pars.V || !pars.C || !pars.F || pars.count > bins.pars.size) return ZPCB200_E_BADARG;
  if (!pars.count) return ZPCB200_OK;
  unbin_kernel<<<(unsigned)((pars.count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bins.pars.base, pars.count, pars);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

int zpcb200_p2g_apic_fcr_binned(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_fixed_corotated model,
                                zpc_stream_t stream) {
  return p2g_binned_launch<0>(bins, tb, g, dt, model.volume, model.E, model.nu, 0.f, stream);
}
int zpcb200_p2g_apic_fcr_binned_halo(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_fixed_corotated model,
                                     zpc_halo_view halo, zpc_stream_t stream) {
  if (halo.peer && (!halo.pos || !halo.recv || halo.world < 1 || halo.world > ZPCB200_HALO_MAX_PEERS || (unsigned)halo.rank >= (unsigned)halo.world ||
                    halo.seg <= 0 || (halo.half != 0 && halo.half != 1)))
    return ZPCB200_E_BADARG;
  return p2g_binned_launch<0>(bins, tb, g, dt, model.volume, model.E, model.nu, 0.f, stream, nullptr, zpcm::PlasticPrm{}, halo);
}
int zpcb200_p2g_apic_drucker_prager_binned(zpc_bins_view bins, float *logJp, zpc_hashtable_view tb, zpc_grids_view g, float dt,
                                           zpc_drucker_prager model, zpc_stream_t stream) {
  if (!logJp) return ZPCB200_E_BADARG;
  return p2g_binned_launch<2>(bins, tb, g, dt, model.volume, model.E, model.nu, 0.f, stream, logJp,
                              zpcm::PlasticPrm{model.cohesion, model.beta, model.yieldSurface, 0.f, model.volumeCorrection});
}
int zpcb200_p2g_apic_nacc_binned(zpc_bins_view bins, float *logJp, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_nacc model,
                                 zpc_stream_t stream) {
  if (!logJp || model.dim != 3) return ZPCB200_E_BADARG;
  return p2g_binned_launch<3>(bins, tb, g, dt, model.volume, model.E, model.nu, 0.f, stream, logJp,
                              zpcm::PlasticPrm{zpcm::nacc_bulk_host(model.E, model.nu), model.xi, model.beta,
                                               zpcm::nacc_msqr_host(model.fa, model.dim), model.hardeningOn});
}
/* zpcb200_rebin_particles that also returns the permutation it applied: dst slot i <- src slot order_out[i] (for per-particle
 * side arrays such as logJp, which the caller permutes with it) */
int zpcb200_rebin_particles_ordered(void *temp, size_t *temp_bytes, zpc_bins_view src, zpc_hashtable_view table, float dx,
                                    zpc_bins_view dst, int *order_out, zpc_stream_t stream) {
  zpc_particles_view none = {};
  if (temp && (src.pars.base == dst.pars.base || dst.pars.size < src.pars.size)) return ZPCB200_E_BADARG;
  return bin_pipeline<true>(temp, temp_bytes, none, src.pars.base, src.pars.size, BinGridLegacy{table}, dx, dst, order_out, (cudaStream_t)stream);
}
int zpcb200_p2g_apic_vonmises_binned(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt,
                                     zpc_vonmises_fixed_corotated model, zpc_stream_t stream) {
  return p2g_binned_launch<1>(bins, tb, g, dt, model.volume, model.E, model.nu, model.yieldStress, stream);
}

int zpcb200_g2p_apic_binned(zpc_bins_view bins, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  return g2p_binned_launch<false>(bins, tb, g, dt, nullptr, stream);
}
/* EquationOfStateConfig on the binned layout: J = one float per particle in bin order (like logJp above) */
int zpcb200_p2g_apic_eos_binned(zpc_bins_view bins, const float *J, zpc_hashtable_view tb, zpc_grids_view g, float dt,
                                zpc_equation_of_state model, zpc_stream_t stream) {
  if (!J) return ZPCB200_E_BADARG;
  return p2g_binned_launch<4>(bins, tb, g, dt, model.volume, 0.f, 0.f, 0.f, stream, const_cast<float *>(J),
                              zpcm::PlasticPrm{model.bulk, model.viscosity, 0.f, 0.f, 0});
}
int zpcb200_g2p_apic_eos_binned(zpc_bins_view bins, float *J, zpc_hashtable_view tb, zpc_grids_view g, float dt, zpc_stream_t stream) {
  if (!J) return ZPCB200_E_BADARG;
  return g2p_binned_launch<true>(bins, tb, g, dt, J, stream);
}
}

// ---- block-binned fast path on SparseGrid<3,f32,8> (round 2; geometry/SparseGrid.hpp:16-188, 275-309) ------------------------------
// Same bins, same kernels: a bin is one octant (4^3 cells) of a side-8 block, BinGridSparse maps octants to (block number, offset).
namespace {
bool sgb_uniform_dx(const zpc_sparsegrid_view &sg, float &dx) {
  const float *M = sg.transform;
  dx = M[0];
  if (!(dx > 0.f) || M[5] != dx || M[10] != dx || M[15] != 1.f) return false;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (i != j && M[4 * i + j] != 0.f) return false;
  return true;
}
bool sgb_table_ok(const zpc_bht_view &t) { return t.keys && t.indices && t.status && t.activeKeys && t.cnt && t.numBuckets * 16u == t.tableSize; }
}  // namespace

extern "C" {
int zpcb200_sg_bin_particles(void *temp, size_t *temp_bytes, zpc_particles_view pars, zpc_sparsegrid_view sg, zpc_bins_view bins,
                             int *order_out, zpc_stream_t stream) {
  float dx;
  if (!sgb_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (temp && (!sgb_table_ok(sg.table) || (pars.count && (!pars.X || !pars.V || !pars.M || !pars.C || !pars.F)))) return ZPCB200_E_BADARG;
  if (temp && bins.pars.size < pars.count) return ZPCB200_E_BADARG;
  return bin_pipeline<false>(temp, temp_bytes, pars, nullptr, pars.count, BinGridSparse{sg.table}, dx, bins, order_out, (cudaStream_t)stream);
}
int zpcb200_sg_rebin_particles(void *temp, size_t *temp_bytes, zpc_bins_view src, zpc_sparsegrid_view sg, zpc_bins_view dst, int *order_out,
                               zpc_stream_t stream) {
  float dx;
  if (!sgb_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (temp && (!sgb_table_ok(sg.table) || !src.pars.base || !dst.pars.base || src.pars.base == dst.pars.base || dst.pars.size < src.pars.size)) return ZPCB200_E_BADARG;
  zpc_particles_view none = {};
  return bin_pipeline<true>(temp, temp_bytes, none, src.pars.base, src.pars.size, BinGridSparse{sg.table}, dx, dst, order_out, (cudaStream_t)stream);
}
}  // extern "C"
template <int MODEL>
static int sg_p2g_binned_launch(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, float volume, float E, float nu, float yield_stress,
                                float *scalar, zpcm::PlasticPrm pp, zpc_stream_t stream) {
  float dx;
  if (!sgb_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (sg.numChannels < 7 || !sg.grid || !sgb_table_ok(sg.table) || bins.pars.numChannels != NCH || !bins.binStart || !bins.binKey || !bins.numBins)
    return ZPCB200_E_BADARG;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.load(std::memory_order_acquire)) {
    ZPC_CUDA(cudaFuncSetAttribute(p2g_binned_kernel<4, MODEL, BinGridSparse>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2GSmem)));
    attr_set.store(true, std::memory_order_release);
  }
  float mu = 0.f, lam = 0.f;
  if (MODEL != 4) zpcm::lame_host(E, nu, mu, lam);
  const bool cache = bins.cellOrder && bins.cellStart && bins.cellOrderValid;
  p2g_binned_kernel<4, MODEL, BinGridSparse><<<bins.binCapacity, P2G_NT, sizeof(P2GSmem), (cudaStream_t)stream>>>(
      bins.pars.base, bins.binStart, bins.binKey, bins.numBins, cache ? bins.cellOrder : nullptr, bins.cellStart, bins.cellOrderValid,
      BinGridSparse{sg.table}, sg.grid, dx, dt, volume, mu, lam, 1, yield_stress, scalar, pp, bins.status, zpc_halo_view{}, sg.numChannels);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
extern "C" {
int zpcb200_sg_p2g_apic_fcr_binned(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, zpc_fixed_corotated model, zpc_stream_t stream) {
  return sg_p2g_binned_launch<0>(bins, sg, dt, model.volume, model.E, model.nu, 0.f, nullptr, zpcm::PlasticPrm{}, stream);
}
/* the other four constitutive models on the binned SparseGrid path: model_kind = ZPC_MODEL_*, model -> the matching struct (host memory),
 * scalar = logJp (Drucker-Prager, NACC; read and written) or J (equation of state; read) per particle in BIN order, NULL otherwise */
int zpcb200_sg_p2g_apic_model_binned(zpc_bins_view bins, float *scalar, zpc_sparsegrid_view sg, float dt, int model_kind, const void *model,
                                     zpc_stream_t stream) {
  if (!model) return ZPCB200_E_BADARG;
  switch (model_kind) {
    case ZPC_MODEL_FIXED_COROTATED: {
      const auto &m = *(const zpc_fixed_corotated *)model;
      return sg_p2g_binned_launch<0>(bins, sg, dt, m.volume, m.E, m.nu, 0.f, nullptr, zpcm::PlasticPrm{}, stream);
    }
    case ZPC_MODEL_VONMISES: {
      const auto &m = *(const zpc_vonmises_fixed_corotated *)model;
      return sg_p2g_binned_launch<1>(bins, sg, dt, m.volume, m.E, m.nu, m.yieldStress, nullptr, zpcm::PlasticPrm{}, stream);
    }
    case ZPC_MODEL_DRUCKER_PRAGER: {
      const auto &m = *(const zpc_drucker_prager *)model;
      if (!scalar) return ZPCB200_E_BADARG;
      return sg_p2g_binned_launch<2>(bins, sg, dt, m.volume, m.E, m.nu, 0.f, scalar,
                                     zpcm::PlasticPrm{m.cohesion, m.beta, m.yieldSurface, 0.f, m.volumeCorrection}, stream);
    }
    case ZPC_MODEL_NACC: {
      const auto &m = *(const zpc_nacc *)model;
      if (!scalar || m.dim != 3) return ZPCB200_E_BADARG;
      return sg_p2g_binned_launch<3>(bins, sg, dt, m.volume, m.E, m.nu, 0.f, scalar,
                                     zpcm::PlasticPrm{zpcm::nacc_bulk_host(m.E, m.nu), m.xi, m.beta, zpcm::nacc_msqr_host(m.fa, m.dim), m.hardeningOn}, stream);
    }
    case ZPC_MODEL_EOS: {
      const auto &m = *(const zpc_equation_of_state *)model;
      if (!scalar) return ZPCB200_E_BADARG;
      return sg_p2g_binned_launch<4>(bins, sg, dt, m.volume, 0.f, 0.f, 0.f, scalar, zpcm::PlasticPrm{m.bulk, m.viscosity, 0.f, 0.f, 0}, stream);
    }
  }
  return ZPCB200_E_BADARG;
}
}  // extern "C"
template <bool EOS>
static int sg_g2p_binned_launch(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, float *scalar, zpc_stream_t stream) {
  float dx;
  if (!sgb_uniform_dx(sg, dx)) return ZPCB200_E_UNSUPPORTED;
  if (sg.numChannels < 4 || !sg.grid || !sgb_table_ok(sg.table) || bins.pars.numChannels != NCH || !bins.binStart || !bins.binKey || !bins.numBins)
    return ZPCB200_E_BADARG;
  const bool cache = bins.cellOrder && bins.cellStart && bins.cellOrderValid;
  g2p_binned_staged_kernel<64, EOS, BinGridSparse><<<bins.binCapacity, 64, 0, (cudaStream_t)stream>>>(
      bins.pars.base, bins.binStart, bins.binKey, bins.numBins, cache ? bins.cellOrder : nullptr, bins.cellStart, BinGridSparse{sg.table}, sg.grid,
      sg.numChannels, dx, dt, scalar, bins.status);
  ZPC_CHECK_LAUNCH();
  if (cache) ZPC_CUDA(cudaMemsetAsync(bins.cellOrderValid, 1, sizeof(int), (cudaStream_t)stream));  // non-zero = valid
  return ZPCB200_OK;
}
extern "C" {
int zpcb200_sg_g2p_apic_binned(zpc_bins_view bins, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream) {
  return sg_g2p_binned_launch<false>(bins, sg, dt, nullptr, stream);
}
/* EquationOfStateConfig: J (one float per particle in bin order) is advanced instead of F (G2P.hpp:69-73) */
int zpcb200_sg_g2p_apic_eos_binned(zpc_bins_view bins, float *J, zpc_sparsegrid_view sg, float dt, zpc_stream_t stream) {
  if (!J) return ZPCB200_E_BADARG;
  return sg_g2p_binned_launch<true>(bins, sg, dt, J, stream);
}
}  // extern "C"

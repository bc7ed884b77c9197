// Kernels shared by the two grid flavours of the drop-in (any-order AoS) path: Grids<f32,3,4> + HashTable (mpm.cu)
// and SparseGrid<3,f32,8> + bht (sparsegrid.cu).  GA = zpcp::LegacyGrid | zpcp::SparseGrid8 (mpm_particle.cuh).
#pragma once
#include "common.cuh"
#include "mpm_math.cuh"
#include "mpm_particle.cuh"

namespace {

// ---- grid -----------------------------------------------------------------------------------------------
// cells = cells per block (64 for Grids<f32,3,4>, 512 for SparseGrid<3,f32,8>)
__global__ void __launch_bounds__(256) clean_grid_kernel(float4 *tiles, const int *cnt, int nch, size_t cap_blocks, int cells) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  const size_t n4 = nb * (size_t)nch * (size_t)(cells / 4);  // float4 per tile
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) tiles[i] = z;
}

// one thread per (block, cell); CS = cells per block = channel stride (64: four blocks per 256-thread CTA)
template <int CS>
__global__ void __launch_bounds__(256) grid_update_kernel(float *tiles, const int *cnt, int nch, size_t cap_blocks, float dt,
                                                          float ex, float ey, float ez, int mode, float *max_vel_sqr) {
  size_t nb = (size_t)*cnt;
  if (nb > cap_blocks) nb = cap_blocks;
  float mx = 0.f;
  for (size_t gc = (size_t)blockIdx.x * 256 + threadIdx.x; gc < nb * CS; gc += (size_t)gridDim.x * 256) {
    const size_t b = gc / CS;
    const int cell = (int)(gc % CS);
    float *t = tiles + b * (size_t)nch * CS;
    float mass = t[cell];
    if (mass != 0.f) {
      float mvx = t[CS + cell], mvy = t[2 * CS + cell], mvz = t[3 * CS + cell];
      if (mode == 1) { mvx += t[4 * CS + cell]; mvy += t[5 * CS + cell]; mvz += t[6 * CS + cell]; }
      mass = 1.f / mass;
      const float vx = mvx * mass + ex * dt, vy = mvy * mass + ey * dt, vz = mvz * mass + ez * dt;
      t[CS + cell] = vx; t[2 * CS + cell] = vy; t[3 * CS + cell] = vz;
      mx = fmaxf(mx, vx * vx + vy * vy + vz * vz);
    } else if (mode == 1) {
      // explicit mode folds rhs into mv for every cell (oracle: mv += rhs before the mass test)
      t[CS + cell] += t[4 * CS + cell]; t[2 * CS + cell] += t[5 * CS + cell]; t[3 * CS + cell] += t[6 * CS + cell];
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

// ---- P2G / G2P on AoS particles, any order -----------------------------------------------------------------
template <class GA>
__global__ void __launch_bounds__(128) p2g_aos_kernel(zpc_particles_view P, GA tb, float *tiles, int nch,
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

template <class GA>
__global__ void __launch_bounds__(128) p2g_aos_vm_kernel(zpc_particles_view P, GA tb, float *tiles, int nch, float dx, float dt,
                                                         float volume, float mu, float lam, float yield_stress) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  float pos[3], vel[3], C[9], F[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pos[d] = P.X[3 * p + d]; vel[d] = P.V[3 * p + d]; }
#pragma unroll
  for (int d = 0; d < 9; ++d) { C[d] = P.C[9 * p + d]; F[d] = P.F[9 * p + d]; }
  zpcp::p2g_scatter_particle_vm(pos, vel, P.M[p], C, F, tb, tiles, nch, dx, dt, volume, mu, lam, yield_stress);
}

template <class GA>
__global__ void __launch_bounds__(128) p2g_aos_eos_kernel(zpc_particles_view P, GA tb, float *tiles, int nch,
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

// DruckerPragerConfig (MODEL 2) / NACCConfig (MODEL 3): P2G.hpp:92-102 — logJp is read, updated by the return mapping and
// written back; the projected F stays in registers.  prm: model 2 {cohesion, beta, yieldSurface, -}, flag = volumeCorrection;
// model 3 {bulk, xi, beta, Msqr}, flag = hardeningOn.
using PlasticParams = zpcm::PlasticPrm;
template <int MODEL, class GA>
__global__ void __launch_bounds__(128) p2g_aos_plastic_kernel(zpc_particles_view P, GA tb, float *tiles, int nch, float dx, float dt,
                                                              float volume, float mu, float lam, PlasticParams prm) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  float pos[3], vel[3], C[9], F[9], contrib[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) { pos[d] = P.X[3 * p + d]; vel[d] = P.V[3 * p + d]; }
#pragma unroll
  for (int d = 0; d < 9; ++d) { C[d] = P.C[9 * p + d]; F[d] = P.F[9 * p + d]; }
  float logJp = P.logJp[p];
  if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, prm.a, prm.b, prm.c, prm.flag != 0, logJp, F, contrib);
  else zpcm::stress_nacc(volume, mu, prm.a, prm.b, prm.c, prm.d, prm.flag != 0, logJp, F, contrib);
  P.logJp[p] = logJp;
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
#pragma unroll
  for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
  zpcp::p2g_scatter_core(pos, vel, P.M[p], C, contrib, tb, tiles, nch, dx);
}

// G2PTransfer<apic> for particle p (G2P.hpp:43-84), F or J variant — one function for the kernel and for tests/hostmath
template <bool EOS, class GA>
ZPC_HD void g2p_aos_particle(const zpc_particles_view &P, size_t p, GA tb, const float *tiles, int nch, float dx, float dt) {
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float pos[3], vel[3] = {0.f, 0.f, 0.f}, C[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] = P.X[3 * p + d];
#pragma unroll
  for (int d = 0; d < 9; ++d) C[d] = 0.f;
  zpcm::Arena ar;
  zpcm::arena_init(ar, dx, pos);
  long long boff[8];
  constexpr int S = GA::S, M = (1 << S) - 1, CS = 1 << (3 * S);
  zpcp::resolve_blocks(ar.corner, tb, nch, boff);
  const int lx0 = ar.corner[0] & M, ly0 = ar.corner[1] & M, lz0 = ar.corner[2] & M;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long off = boff[((lx >> S) << 2) | ((ly >> S) << 1) | (lz >> S)];
        if (off < 0) continue;
        const float *t = tiles + off + zpcp::cell_offset<GA>(lx, ly, lz);
        const float xixp[3] = {(float)i * dx - ar.local[0], (float)j * dx - ar.local[1], (float)k * dx - ar.local[2]};
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
        const float vi[3] = {zpcm::grid_load(t + CS), zpcm::grid_load(t + 2 * CS), zpcm::grid_load(t + 3 * CS)};
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

template <bool EOS, class GA>
__global__ void __launch_bounds__(128) g2p_aos_kernel(zpc_particles_view P, GA tb, const float *tiles,
                                                      int nch, float dx, float dt) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.count) return;
  g2p_aos_particle<EOS>(P, p, tb, tiles, nch, dx, dt);
}

// G2P2GTransfer (simulation/transfer/G2P2G.hpp:49-141) for particle p: C gathered from the DOF vector gridv (3 floats per node,
// node = block * cells + cell), trial F = (I + dt C) F (kept in registers), stress of the trial state, W * (contrib * D_inv) * xixp
// scattered into the DOF vector gridr.  MODEL 0 fixed-corotated, 1 von Mises, 2 Drucker-Prager, 3 NACC (logJp read, not written
// back), 4 equation of state (J).  prm as in PlasticParams; model 1: a = yield stress; model 4: a = bulk, b = viscosity.
template <int MODEL, class GA>
ZPC_HD void g2p2g_particle(const zpc_particles_view &P, size_t p, GA tb, const float *gridv, float *gridr, float dx, float dt,
                           float volume, float mu, float lam, const PlasticParams &prm) {
  constexpr int S = GA::S, M = (1 << S) - 1, CS = 1 << (3 * S);
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  float pos[3], C[9], contrib[9];
#pragma unroll
  for (int d = 0; d < 3; ++d) pos[d] = P.X[3 * p + d];
#pragma unroll
  for (int d = 0; d < 9; ++d) C[d] = 0.f;
  zpcm::Arena ar;
  zpcm::arena_init(ar, dx, pos);
  long long boff[8];
  zpcp::resolve_blocks(ar.corner, tb, 1, boff);                 // nch = 1: offsets in nodes (block * cells)
  const int lx0 = ar.corner[0] & M, ly0 = ar.corner[1] & M, lz0 = ar.corner[2] & M;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long off = boff[((lx >> S) << 2) | ((ly >> S) << 1) | (lz >> S)];
        if (off < 0) continue;
        const float *t = gridv + 3 * (off + zpcp::cell_offset<GA>(lx, ly, lz));
        const float xixp[3] = {(float)i * dx - ar.local[0], (float)j * dx - ar.local[1], (float)k * dx - ar.local[2]};
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
        const float vi[3] = {zpcm::grid_load(t), zpcm::grid_load(t + 1), zpcm::grid_load(t + 2)};
#pragma unroll
        for (int d = 0; d < 9; ++d) C[d] += W * vi[d % 3] * xixp[d / 3] * D_inv;
      }
  if constexpr (MODEL == 4) {
    float J = P.J[p];
    J = (1 + (C[0] + C[4] + C[8]) * dt) * J;
    zpcm::eos_contrib(C, J, volume, prm.a, prm.b, contrib);
  } else {
    float Fo[9], tmp[9], F[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) { Fo[d] = P.F[9 * p + d]; tmp[d] = C[d] * dt + ((d & 3) ? 0.f : 1.f); }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) F[3 * c + r] = tmp[r] * Fo[3 * c] + tmp[3 + r] * Fo[3 * c + 1] + tmp[6 + r] * Fo[3 * c + 2];
    if constexpr (MODEL == 0) zpcm::stress_fcr(volume, mu, lam, F, contrib);
    else if constexpr (MODEL == 1) zpcm::stress_vonmises(volume, mu, lam, prm.a, F, contrib);
    else {
      float logJp = P.logJp[p];
      if constexpr (MODEL == 2) zpcm::stress_sand(volume, mu, lam, prm.a, prm.b, prm.c, prm.flag != 0, logJp, F, contrib);
      else zpcm::stress_nacc(volume, mu, prm.a, prm.b, prm.c, prm.d, prm.flag != 0, logJp, F, contrib);
    }
  }
#pragma unroll
  for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * D_inv;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lx = lx0 + i, ly = ly0 + j, lz = lz0 + k;
        const long long off = boff[((lx >> S) << 2) | ((ly >> S) << 1) | (lz >> S)];
        if (off < 0) continue;
        float *r = gridr + 3 * (off + zpcp::cell_offset<GA>(lx, ly, lz));
        const float x0 = (float)i * dx - ar.local[0], x1 = (float)j * dx - ar.local[1], x2 = (float)k * dx - ar.local[2];
        const float W = ar.w[0][i] * ar.w[1][j] * ar.w[2][k];
#pragma unroll
        for (int d = 0; d < 3; ++d) zpcm::grid_add(r + d, W * (contrib[d] * x0 + contrib[3 + d] * x1 + contrib[6 + d] * x2));
      }
}
template <int MODEL, class GA>
__global__ void __launch_bounds__(128) g2p2g_aos_kernel(zpc_particles_view P, GA tb, const float *gridv, float *gridr, float dx, float dt,
                                                        float volume, float mu, float lam, PlasticParams prm) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P.count) g2p2g_particle<MODEL>(P, p, tb, gridv, gridr, dx, dt, volume, mu, lam, prm);
}

}  // namespace

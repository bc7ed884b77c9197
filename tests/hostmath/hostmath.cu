// TEST INFRASTRUCTURE — compiles the product's per-particle device math (zpc_b200/csrc/mpm_math.cuh, __host__ __device__)
// for the CPU so that tests without a GPU can compare it with the oracle: a transcription error in a stress model or in
// the stencil shows up here.  Not linked into libzpcb200.so; nothing in zpc_b200/ uses it.
#include "../../zpc_b200/csrc/mpm_math.cuh"
#include "../../zpc_b200/csrc/lbvh_core.cuh"
#include "../../zpc_b200/csrc/mpm_kernels.cuh"
#include "../../zpc_b200/csrc/p2g_sweep.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

extern "C" {
void hm_svd3(const float *F, float *U, float *S, float *V) {
  float f[9], u[9], s[3], v[9];
  for (int d = 0; d < 9; ++d) f[d] = F[d];
  zpcm::svd3(f, u, s, v);
  for (int d = 0; d < 9; ++d) { U[d] = u[d]; V[d] = v[d]; }
  for (int d = 0; d < 3; ++d) S[d] = s[d];
}
void hm_lame(float E, float nu, float *mu, float *lam) { zpcm::lame_host(E, nu, *mu, *lam); }
void hm_nacc_consts(float E, float nu, float fa, int dim, float *bulk, float *msqr) {
  *bulk = zpcm::nacc_bulk_host(E, nu);
  *msqr = zpcm::nacc_msqr_host(fa, dim);
}
// model: 0 fixed-corotated, 1 von Mises {yield}, 2 Drucker-Prager {cohesion, beta, yieldSurface, volumeCorrection},
// 3 NACC {bulk, xi, beta, Msqr, hardeningOn}; n matrices, logJp in/out for 2 and 3
void hm_stress(int model, int n, float volume, float mu, float lam, const float *prm, float *logJp, const float *F, float *PF) {
  for (int p = 0; p < n; ++p) {
    float f[9], pf[9];
    for (int d = 0; d < 9; ++d) f[d] = F[9 * p + d];
    if (model == 0) zpcm::stress_fcr(volume, mu, lam, f, pf);
    else if (model == 10) zpcm::stress_fcr_lean(volume, mu, lam, f, pf);   // the binned fast path's arithmetic
    else if (model == 11) zpcm::stress_fcr_lean<true>(volume, mu, lam, f, pf);   // ... with the converged-sweep skip (sweep variant 8)
    else if (model == 1) zpcm::stress_vonmises(volume, mu, lam, prm[0], f, pf);
    else if (model == 2) zpcm::stress_sand(volume, mu, lam, prm[0], prm[1], prm[2], prm[3] != 0.f, logJp[p], f, pf);
    else zpcm::stress_nacc(volume, mu, prm[0], prm[1], prm[2], prm[3], prm[4] != 0.f, logJp[p], f, pf);
    for (int d = 0; d < 9; ++d) PF[9 * p + d] = pf[d];
  }
}
// AnalyticLevelSet<Cuboid>: signed distance and finite-difference normal at n points
void hm_cuboid(int n, const float *x, const float *mn, const float *mx, float *sdf, float *normal) {
  for (int p = 0; p < n; ++p) {
    sdf[p] = zpcm::cuboid_sdf(x[3 * p], x[3 * p + 1], x[3 * p + 2], mn, mx);
    zpcm::cuboid_normal(x[3 * p], x[3 * p + 1], x[3 * p + 2], mn, mx, normal[3 * p], normal[3 * p + 1], normal[3 * p + 2]);
  }
}
// zpcm::collide at n nodes: positions p[n][3], velocities v[n][3] in/out
void hm_collide(int n, zpc_collider col, const float *p, float *v) {
  for (int i = 0; i < n; ++i) zpcm::collide(col, p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2]);
}
// LocalArena: corner[3], local[3], w[9] per position
void hm_arena(int n, float dx, const float *x, int *corner, float *local, float *w) {
  for (int p = 0; p < n; ++p) {
    zpcm::Arena a;
    float pos[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    zpcm::arena_init(a, dx, pos);
    for (int d = 0; d < 3; ++d) {
      corner[3 * p + d] = a.corner[d];
      local[3 * p + d] = a.local[d];
      for (int k = 0; k < 3; ++k) w[9 * p + 3 * d + k] = a.w[d][k];
    }
  }
}
// the same through the branch-free exact division of the binned G2P (arena_init<true>: 1/dx and two FMAs)
void hm_arena_fastdiv(int n, float dx, const float *x, int *corner, float *local, float *w) {
  const float dx_inv = 1.0f / dx;
  for (int p = 0; p < n; ++p) {
    zpcm::Arena a;
    float pos[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    zpcm::arena_init<true>(a, dx, pos, dx_inv);
    for (int d = 0; d < 3; ++d) {
      corner[3 * p + d] = a.corner[d];
      local[3 * p + d] = a.local[d];
      for (int k = 0; k < 3; ++k) w[9 * p + 3 * d + k] = a.w[d][k];
    }
  }
}
// the per-node functions of the LBvh build (what lbvh.cu's kernels call), run index by index with a host stable sort and scan in
// place of the device primitives; box = the padded whole box (6 floats).  n > 2.
void hm_lbvh_build(int n, const float *prims, const float *box, int *auxIndices, int *parents, int *levels, int *leafInds) {
  const int numTrunk = n - 1;
  std::vector<unsigned> codes(n), smcs(n);
  std::vector<int> ids(n), pInds(n), tPars(numTrunk), tRs(numTrunk), tDst(numTrunk), lPars(n), lLcas(n), lDepths(n + 1), lOffsets(n + 1);
  for (int i = 0; i < n; ++i) { codes[i] = zpcb::morton_of(prims, box, i); ids[i] = i; }
  std::iota(pInds.begin(), pInds.end(), 0);
  std::stable_sort(pInds.begin(), pInds.end(), [&](int a, int b) { return codes[a] < codes[b]; });
  for (int i = 0; i < n; ++i) smcs[i] = codes[pInds[i]];
  for (int i = 0; i < n; ++i) lDepths[i] = 1;
  lDepths[n] = 0;
  for (int idx = numTrunk - 1; idx >= 0; --idx) zpcb::topo_node(idx, smcs.data(), numTrunk, tPars.data(), tRs.data(), lPars.data(), lDepths.data());
  int run = 0;
  for (int i = 0; i <= n; ++i) { lOffsets[i] = run; run += lDepths[i]; }
  for (int idx = n - 1; idx >= 0; --idx)
    zpcb::supp_topo_leaf(idx, n, lOffsets.data(), lPars.data(), tPars.data(), pInds.data(), tDst.data(), lLcas.data(), levels, auxIndices, leafInds);
  for (int idx = n - 1; idx >= 0; --idx)
    zpcb::reorder_node(idx, n, lOffsets.data(), lPars.data(), lLcas.data(), tPars.data(), tRs.data(), tDst.data(), auxIndices, parents);
}
// zpcb::iter_neighbors on the host: ids of the primitives overlapping bv, in visiting order; returns their number
int hm_lbvh_iter_neighbors(int n, const float *bvs, const int *auxIndices, const int *levels, const float *bv, int *out, int cap) {
  int c = 0;
  zpcb::iter_neighbors(n, bvs, auxIndices, levels, bv, [&](int prim) { if (c < cap) out[c] = prim; ++c; });
  return c;
}
// The AoS P2G / G2P of the product, particle by particle on the host: the same zpcp::p2g_scatter_* / g2p_aos_particle the kernels
// call (mpm_particle.cuh, mpm_kernels.cuh), against the legacy hash table.  model: 0 fixed-corotated, 1 von Mises {yield},
// 2 Drucker-Prager {cohesion, beta, yieldSurface, volumeCorrection}, 3 NACC {bulk, xi, beta, Msqr, hardeningOn},
// 4 equation of state {bulk, viscosity}.
void hm_p2g_aos(int model, zpc_particles_view P, zpc_hashtable_view tb, float *tiles, float dx, float dt, float volume, float mu,
                float lam, const float *prm) {
  const zpcp::LegacyGrid g{tb};
  const float dx_inv = 1.0f / dx, D_inv = 4.f * dx_inv * dx_inv;
  for (size_t p = 0; p < P.count; ++p) {
    float pos[3], vel[3], C[9], F[9], contrib[9];
    for (int d = 0; d < 3; ++d) { pos[d] = P.X[3 * p + d]; vel[d] = P.V[3 * p + d]; }
    for (int d = 0; d < 9; ++d) { C[d] = P.C[9 * p + d]; F[d] = P.F ? P.F[9 * p + d] : 0.f; }
    if (model == 0) zpcp::p2g_scatter_particle(pos, vel, P.M[p], C, F, g, tiles, 7, dx, dt, volume, mu, lam);
    else if (model == 1) zpcp::p2g_scatter_particle_vm(pos, vel, P.M[p], C, F, g, tiles, 7, dx, dt, volume, mu, lam, prm[0]);
    else if (model == 4) zpcp::p2g_scatter_particle_eos(pos, vel, P.M[p], C, P.J[p], g, tiles, 7, dx, dt, volume, prm[0], prm[1]);
    else {  // the body of p2g_aos_plastic_kernel
      float logJp = P.logJp[p];
      if (model == 2) zpcm::stress_sand(volume, mu, lam, prm[0], prm[1], prm[2], prm[3] != 0.f, logJp, F, contrib);
      else zpcm::stress_nacc(volume, mu, prm[0], prm[1], prm[2], prm[3], prm[4] != 0.f, logJp, F, contrib);
      P.logJp[p] = logJp;
      for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;
      zpcp::p2g_scatter_core(pos, vel, P.M[p], C, contrib, g, tiles, 7, dx);
    }
  }
}
void hm_g2p_aos(int eos, zpc_particles_view P, zpc_hashtable_view tb, const float *tiles, float dx, float dt) {
  const zpcp::LegacyGrid g{tb};
  for (size_t p = 0; p < P.count; ++p) {
    if (eos) g2p_aos_particle<true>(P, p, g, tiles, 7, dx, dt);
    else g2p_aos_particle<false>(P, p, g, tiles, 7, dx, dt);
  }
}
// g2p2g_particle (the body of g2p2g_aos_kernel) particle by particle; model / prm as hm_p2g_aos
void hm_g2p2g(int model, zpc_particles_view P, zpc_hashtable_view tb, const float *gridv, float *gridr, float dx, float dt, float volume,
              float mu, float lam, const float *prm) {
  const zpcp::LegacyGrid g{tb};
  PlasticParams pp{prm[0], prm[1], prm[2], prm[3], prm[4] != 0.f};
  if (model == 2) pp.flag = prm[3] != 0.f;
  for (size_t p = 0; p < P.count; ++p) {
    if (model == 0) g2p2g_particle<0>(P, p, g, gridv, gridr, dx, dt, volume, mu, lam, pp);
    else if (model == 1) g2p2g_particle<1>(P, p, g, gridv, gridr, dx, dt, volume, mu, lam, pp);
    else if (model == 2) g2p2g_particle<2>(P, p, g, gridv, gridr, dx, dt, volume, mu, lam, pp);
    else if (model == 3) g2p2g_particle<3>(P, p, g, gridv, gridr, dx, dt, volume, mu, lam, pp);
    else g2p2g_particle<4>(P, p, g, gridv, gridr, dx, dt, volume, mu, lam, pp);
  }
}
// the table lookups the kernels use (mpm_math.cuh), on host copies of the tables
void hm_table_query(int n, const int *keys3, int table_size, const int *tkeys, const int *tindices, int *out) {
  for (int i = 0; i < n; ++i) out[i] = zpcm::table_query(keys3[3 * i], keys3[3 * i + 1], keys3[3 * i + 2], table_size, tkeys, tindices);
}
void hm_bht_query(int n, const int *keys3, zpc_bht_view tb, int *out) {
  for (int i = 0; i < n; ++i) out[i] = zpcm::bht_query(keys3[3 * i], keys3[3 * i + 1], keys3[3 * i + 2], tb);
}
// The three sweep variants of the binned P2G (p2g_sweep.cuh) over n particles of ONE cell: records written by write_record<VAR> from
// (d0, mass, A, a, B, Kd), then swept by every lane role.  out3[27][7]: sweep_cell, lane = node (ox*9 + oy*3 + oz);
// out4 / out5 [9][7][3]: sweep_cells3 / sweep_cells3_packed, lane = column (ox*3 + oy), last index = oz.
static void hm_coef(int o, float &a, float &b, float &c) {   // the polynomial table of p2g_binned_kernel
  a = o == 1 ? -1.0f : 0.5f; b = o == 0 ? -1.5f : (o == 1 ? 2.0f : -0.5f); c = o == 0 ? 1.125f : (o == 1 ? -0.25f : 0.125f);
}
void hm_p2g_sweeps(int n, const float *d0, const float *mass, const float *A, const float *a, const float *B, const float *Kd, float *out3,
                   float *out4, float *out5) {
  std::vector<float4> r3((size_t)7 * n + 8), r4((size_t)zpcs::rec_at<4>(n) + 8), r5((size_t)zpcs::rec_at<5>(n) + 8);
  for (int i = 0; i < n; ++i) {
    float d[3], Av[3], av[3], Bv[9], Kv[9];
    for (int k = 0; k < 3; ++k) { d[k] = d0[3 * i + k]; Av[k] = A[3 * i + k]; av[k] = a[3 * i + k]; }
    for (int k = 0; k < 9; ++k) { Bv[k] = B[9 * i + k]; Kv[k] = Kd[9 * i + k]; }
    zpcs::write_record<3>(r3.data() + zpcs::rec_at<3>(i), d, mass[i], Av, av, Bv, Kv);
    zpcs::write_record<4>(r4.data() + zpcs::rec_at<4>(i), d, mass[i], Av, av, Bv, Kv);
    zpcs::write_record<5>(r5.data() + zpcs::rec_at<5>(i), d, mass[i], Av, av, Bv, Kv);
  }
  for (int lane = 0; lane < 27; ++lane) {
    const int ox = lane / 9, oy = (lane / 3) % 3, oz = lane % 3;
    zpcs::LaneCoef L;
    hm_coef(ox, L.ax, L.bx, L.cx); hm_coef(oy, L.ay, L.by, L.cy); hm_coef(oz, L.az, L.bz, L.cz);
    L.fx = (float)ox; L.fy = (float)oy; L.fz = (float)oz;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    zpcs::sweep_cell(r3.data(), 0, n, L, acc);
    for (int ch = 0; ch < 7; ++ch) out3[lane * 7 + ch] = acc[ch];
  }
  for (int col = 0; col < 9; ++col) {
    const int ox = col / 3, oy = col % 3;
    zpcs::ColCoef L;
    hm_coef(ox, L.ax, L.bx, L.cx); hm_coef(oy, L.ay, L.by, L.cy);
    L.fx = (float)ox; L.fy = (float)oy;
    float acc[7][3] = {};
    zpcs::sweep_cells3(r4.data(), 0, n, n, L, acc);
    float accm[3] = {0.f, 0.f, 0.f};
    float2 accp[3][3];
    for (int q = 0; q < 3; ++q) for (int k = 0; k < 3; ++k) accp[q][k] = make_float2(0.f, 0.f);
    zpcs::sweep_cells3_packed(r5.data(), 0, n, n, L, accm, accp);
    for (int k = 0; k < 3; ++k) {
      for (int ch = 0; ch < 7; ++ch) out4[(col * 7 + ch) * 3 + k] = acc[ch][k];
      out5[(col * 7 + 0) * 3 + k] = accm[k];
      for (int q = 0; q < 3; ++q) { out5[(col * 7 + 1 + 2 * q) * 3 + k] = accp[q][k].x; out5[(col * 7 + 2 + 2 * q) * 3 + k] = accp[q][k].y; }
    }
  }
}
}

// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI driver around the UNMODIFIED reference (zenustech/zpc) headers, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libzpcref.so (git-ignored).  It composes the
// explicit APIC substep exactly as SURVEY.md §3.1 lists it, on the reference's own host policies
// (seq_exec / omp_exec), and exposes the reference's serial/OpenMP sort / scan / reduce.
// Used only by tests/ (to pin oracle/*.c) and by bench.py's reference arm.  No reference source is
// copied: every functor below is the reference's, instantiated here.
#include <cstdint>
#include <cstring>
#include <vector>

#include "zensim/container/HashTable.hpp"
#include "zensim/container/Vector.hpp"
#include "zensim/execution/ExecutionPolicy.hpp"
#include "zensim/geometry/AnalyticLevelSet.h"
#include "zensim/geometry/Collider.h"
#include "zensim/geometry/SparseLevelSet.hpp"
#include "zensim/geometry/Structure.hpp"
#include "zensim/geometry/Structurefree.hpp"
#include "zensim/omp/execution/ExecutionPolicy.hpp"
#include "zensim/physics/ConstitutiveModel_Vol_dP.hpp"
#include "zensim/simulation/grid/GridOp.hpp"        // patched copy (template disambiguator) on -I path
#include "zensim/simulation/sparsity/SparsityOp.hpp"
#include "zensim/simulation/transfer/G2P.hpp"       // patched copy on -I path
#include "zensim/simulation/transfer/P2G.hpp"

using namespace zs;

namespace {
  struct RefMpm {
    int n;
    float dx;
    int nthreads;  // 0 → seq_exec(), else omp_exec().threads(nthreads)
    Particles<f32, 3> pars;
    HashTable<i32, 3, int> table;
    Grids<f32, 3, 4> grids;
    Vector<float> maxVel;
    int nblocks{0};
    RefMpm(int n_, float dx_, int nthreads_, int expectedBlocks)
        : n{n_},
          dx{dx_},
          nthreads{nthreads_},
          pars{(size_t)n_},
          table{(size_t)expectedBlocks},
          grids{{{"m", 1}, {"v", 3}, {"rhs", 3}}, dx_, (size_t)expectedBlocks},
          maxVel{1} {
      pars.addAttr("m", attrib_e::scalar);
      pars.addAttr("v", attrib_e::vector);
      pars.addAttr("F", attrib_e::matrix);
      pars.addAttr("C", attrib_e::matrix);
      pars.addAttr("J", attrib_e::scalar);
      pars.addAttr("logJp", attrib_e::scalar);
    }
  };

  template <typename F> void with_policy(int nthreads, F &&f) {
    if (nthreads <= 0) {
      auto pol = seq_exec();
      f(pol, exec_seq);
    } else {
      auto pol = omp_exec().threads(nthreads);
      f(pol, exec_omp);
    }
  }
}  // namespace

extern "C" {

void *zpcref_mpm_create(int n, float dx, int nthreads, int expectedBlocks) {
  return new RefMpm(n, dx, nthreads, expectedBlocks);
}
void zpcref_mpm_destroy(void *h) { delete (RefMpm *)h; }

void zpcref_mpm_set_particles(void *h, const float *x, const float *v, const float *m,
                              const float *C, const float *F) {
  auto &s = *(RefMpm *)h;
  std::memcpy(s.pars.attrVector("x").data(), x, sizeof(float) * 3 * s.n);
  std::memcpy(s.pars.attrVector("v").data(), v, sizeof(float) * 3 * s.n);
  std::memcpy(s.pars.attrScalar("m").data(), m, sizeof(float) * s.n);
  std::memcpy(s.pars.attrMatrix("C").data(), C, sizeof(float) * 9 * s.n);
  std::memcpy(s.pars.attrMatrix("F").data(), F, sizeof(float) * 9 * s.n);
}
void zpcref_mpm_get_particles(void *h, float *x, float *v, float *C, float *F) {
  auto &s = *(RefMpm *)h;
  std::memcpy(x, s.pars.attrVector("x").data(), sizeof(float) * 3 * s.n);
  std::memcpy(v, s.pars.attrVector("v").data(), sizeof(float) * 3 * s.n);
  std::memcpy(C, s.pars.attrMatrix("C").data(), sizeof(float) * 9 * s.n);
  std::memcpy(F, s.pars.attrMatrix("F").data(), sizeof(float) * 9 * s.n);
}

/// SURVEY §3.1 lines 1-4: CleanSparsity, ComputeSparsity, EnlargeSparsity{0,2}; returns table.size()
int zpcref_mpm_partition(void *h) {
  auto &s = *(RefMpm *)h;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.table._tableSize), CleanSparsity{tag, s.table});
    pol(range(s.n), ComputeSparsity{tag, s.dx, 4, s.table, s.pars.attrVector("x")});
    const int cnt = s.table.size();
    pol(range(cnt),
        EnlargeSparsity{tag, s.table, vec<int, 3>{0, 0, 0}, vec<int, 3>{2, 2, 2}});
  });
  s.nblocks = s.table.size();
  if ((size_t)s.nblocks > s.grids.grid(collocated_c).numBlocks())
    s.grids.grid(collocated_c).resize(s.nblocks);
  return s.nblocks;
}
void zpcref_mpm_get_keys(void *h, int *keys) {
  auto &s = *(RefMpm *)h;
  std::memcpy(keys, s.table._activeKeys.data(), sizeof(int) * 3 * s.nblocks);
}
int zpcref_mpm_table_size(void *h) { return ((RefMpm *)h)->table._tableSize; }
/// raw table arrays: keys[tableSize*3], indices[tableSize]
void zpcref_mpm_get_table(void *h, int *keys, int *indices) {
  auto &s = *(RefMpm *)h;
  std::memcpy(keys, s.table._table.keys.data(), sizeof(int) * 3 * s.table._tableSize);
  std::memcpy(indices, s.table._table.indices.data(), sizeof(int) * s.table._tableSize);
}

void zpcref_mpm_clean_grid(void *h) {
  auto &s = *(RefMpm *)h;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(Collapse{(size_t)s.nblocks, (size_t)64}, CleanGridBlocks{tag, s.grids});
  });
}
void zpcref_mpm_p2g(void *h, float dt, float E, float nu, float volume) {
  auto &s = *(RefMpm *)h;
  FixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.volume = volume;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        P2GTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
  });
}
void zpcref_mpm_p2g_vonmises(void *h, float dt, float E, float nu, float yieldStress, float volume) {
  auto &s = *(RefMpm *)h;
  VonMisesFixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.yieldStress = yieldStress;
  model.volume = volume;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        P2GTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
  });
}
/// mode 0: ComputeGridBlockVelocity as shipped (v = mv/m + g dt; rhs ignored, GridOp.hpp:71-110).
/// mode 1: explicit update v = (mv + rhs)/m + g dt, composed as "mv += rhs" then the functor.
void zpcref_mpm_grid_update(void *h, float dt, float gravity, int mode) {
  auto &s = *(RefMpm *)h;
  s.maxVel.setVal(0.f);
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    if (mode == 1) {
      auto gv = proxy<decltype(tag)::value>(s.grids);
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, [gv](int b, int c) mutable {
        auto block = gv[b];
        for (int d = 0; d != 3; ++d) block(1 + d, c) += block(4 + d, c);
      });
    }
    pol(Collapse{(size_t)s.nblocks, (size_t)64},
        ComputeGridBlockVelocity{tag, wrapv<transfer_scheme_e::apic>{}, s.grids, dt, gravity,
                                 s.maxVel.data()});
  });
}
float zpcref_mpm_get_maxvel(void *h) { return ((RefMpm *)h)->maxVel.getVal(); }
void zpcref_mpm_g2p(void *h, float dt) {
  auto &s = *(RefMpm *)h;
  FixedCorotatedConfig model{};
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        G2PTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.grids, s.table, s.pars});
  });
}
/// EquationOfStateConfig variants (P2G.hpp:66-87, G2P.hpp:69-73)
void zpcref_mpm_set_J(void *h, const float *J) {
  auto &s = *(RefMpm *)h;
  std::memcpy(s.pars.attrScalar("J").data(), J, sizeof(float) * s.n);
}
void zpcref_mpm_get_J(void *h, float *J) {
  auto &s = *(RefMpm *)h;
  std::memcpy(J, s.pars.attrScalar("J").data(), sizeof(float) * s.n);
}
void zpcref_mpm_p2g_eos(void *h, float dt, float bulk, float gamma, float viscosity, float volume) {
  auto &s = *(RefMpm *)h;
  EquationOfStateConfig model{};
  model.bulk = bulk;
  model.gamma = gamma;
  model.viscosity = viscosity;
  model.volume = volume;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        P2GTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
  });
}
/// DruckerPragerConfig / NACCConfig (P2G.hpp:92-102): per-particle logJp, read and written back by P2G
void zpcref_mpm_set_logJp(void *h, const float *logJp) {
  auto &s = *(RefMpm *)h;
  std::memcpy(s.pars.attrScalar("logJp").data(), logJp, sizeof(float) * s.n);
}
void zpcref_mpm_get_logJp(void *h, float *logJp) {
  auto &s = *(RefMpm *)h;
  std::memcpy(logJp, s.pars.attrScalar("logJp").data(), sizeof(float) * s.n);
}
void zpcref_mpm_p2g_sand(void *h, float dt, float E, float nu, float cohesion, float beta, float yieldSurface,
                         int volumeCorrection, float volume) {
  auto &s = *(RefMpm *)h;
  DruckerPragerConfig model{};
  model.E = E;
  model.nu = nu;
  model.cohesion = cohesion;
  model.beta = beta;
  model.yieldSurface = yieldSurface;
  model.volumeCorrection = volumeCorrection != 0;
  model.volume = volume;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        P2GTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
  });
}
void zpcref_mpm_p2g_nacc(void *h, float dt, float E, float nu, float fa, float xi, float beta, int hardeningOn,
                         float volume) {
  auto &s = *(RefMpm *)h;
  NACCConfig model{};
  model.E = E;
  model.nu = nu;
  model.fa = fa;
  model.xi = xi;
  model.beta = beta;
  model.hardeningOn = hardeningOn != 0;
  model.volume = volume;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        P2GTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
  });
}
void zpcref_mpm_g2p_eos(void *h, float dt) {
  auto &s = *(RefMpm *)h;
  EquationOfStateConfig model{};
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(range(s.n),
        G2PTransfer{tag, wrapv<transfer_scheme_e::apic>{}, dt, model, s.grids, s.table, s.pars});
  });
}
/// GridMomentumToVelocity (GridOp.hpp:184-214) on the collocated grid: v = mv / m, max |v|^2; no gravity, rhs untouched
void zpcref_mpm_momentum_to_velocity(void *h) {
  auto &s = *(RefMpm *)h;
  s.maxVel.setVal(0.f);
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(Collapse{(size_t)s.nblocks, (size_t)64},
        GridMomentumToVelocity{tag, s.grids.grid(collocated_c), 0, 1, s.maxVel.data()});
  });
}
/// GridAngularMomentum (GridOp.hpp:216-262): out6 = sum x cross mv (3), sum mv (3), accumulated in double
void zpcref_mpm_angular_momentum(void *h, double *out6) {
  auto &s = *(RefMpm *)h;
  Vector<double> sum{6};
  for (int i = 0; i != 6; ++i) sum.setVal(0., i);
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    pol(Collapse{(size_t)s.nblocks, (size_t)64},
        GridAngularMomentum{tag, s.table, s.grids.grid(collocated_c), 0, 1, sum.data()});
  });
  for (int i = 0; i != 6; ++i) out6[i] = sum.getVal(i);
}
/// ApplyBoundaryConditionOnGridBlocks with a static analytic collider (GridOp.hpp:112-164)
void zpcref_mpm_apply_boundary(void *h, int geom, int type, const float *p0, const float *p1) {
  auto &s = *(RefMpm *)h;
  const auto ct = type == 0 ? collider_e::Sticky : (type == 1 ? collider_e::Slip : collider_e::Separate);
  using TV = vec<float, 3>;
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    if (geom == 0) {
      Collider col{AnalyticLevelSet<analytic_geometry_e::Plane, float, 3>{TV{p0[0], p0[1], p0[2]}, TV{p1[0], p1[1], p1[2]}}, ct};
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    } else if (geom == 2) {  // Cuboid{min, max} (AnalyticLevelSet.h:55-126)
      Collider col{AnalyticLevelSet<analytic_geometry_e::Cuboid, float, 3>{TV{p0[0], p0[1], p0[2]}, TV{p1[0], p1[1], p1[2]}}, ct};
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    } else {
      Collider col{AnalyticLevelSet<analytic_geometry_e::Sphere, float, 3>{TV{p0[0], p0[1], p0[2]}, p1[0]}, ct};
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    }
  });
}
/// the same with a rigid motion: motion = b[3], dbdt[3], R[9] row-major, omega[3], s, dsdt (Collider.h:16-24, 136-143)
void zpcref_mpm_apply_boundary_moving(void *h, int geom, int type, const float *p0, const float *p1, const float *m) {
  auto &s = *(RefMpm *)h;
  const auto ct = type == 0 ? collider_e::Sticky : (type == 1 ? collider_e::Slip : collider_e::Separate);
  using TV = vec<float, 3>;
  auto setup = [&](auto &col) {
    col.setTranslation(TV{m[0], m[1], m[2]}, TV{m[3], m[4], m[5]});
    vec<float, 3, 3> R{};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R(i, j) = m[6 + 3 * i + j];
    AngularVelocity<float, 3> om{};
    om.omega = TV{m[15], m[16], m[17]};
    col.setRotation(Rotation<float, 3>{R}, om);
    col.s = m[18];
    col.dsdt = m[19];
  };
  with_policy(s.nthreads, [&](auto &pol, auto tag) {
    if (geom == 0) {
      Collider col{AnalyticLevelSet<analytic_geometry_e::Plane, float, 3>{TV{p0[0], p0[1], p0[2]}, TV{p1[0], p1[1], p1[2]}}, ct};
      setup(col);
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    } else if (geom == 2) {
      Collider col{AnalyticLevelSet<analytic_geometry_e::Cuboid, float, 3>{TV{p0[0], p0[1], p0[2]}, TV{p1[0], p1[1], p1[2]}}, ct};
      setup(col);
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    } else {
      Collider col{AnalyticLevelSet<analytic_geometry_e::Sphere, float, 3>{TV{p0[0], p0[1], p0[2]}, p1[0]}, ct};
      setup(col);
      pol(Collapse{(size_t)s.nblocks, (size_t)64}, ApplyBoundaryConditionOnGridBlocks{tag, col, s.table, s.grids});
    }
  });
}
/// grid tiles in table numbering: out[nblocks][7][64]
void zpcref_mpm_get_grid(void *h, float *out) {
  auto &s = *(RefMpm *)h;
  std::memcpy(out, s.grids.grid(collocated_c).blocks.data(), sizeof(float) * 7 * 64 * s.nblocks);
}

/// ---- primitives (reference serial / OpenMP policies) ----
#define ZPCREF_SORT_PAIR(SUFFIX, KT)                                                               \
  void zpcref_radix_sort_pair_##SUFFIX(int nthreads, const KT *kin, const int *vin, KT *kout,     \
                                       int *vout, size_t n, int sbit, int ebit) {                 \
    KT *ki = const_cast<KT *>(kin);                                                                \
    int *vi = const_cast<int *>(vin);                                                              \
    with_policy(nthreads, [&](auto &pol, auto) {                                                   \
      radix_sort_pair(pol, ki, vi, kout, vout, (std::ptrdiff_t)n, sbit, ebit);                     \
    });                                                                                            \
  }                                                                                                \
  void zpcref_radix_sort_##SUFFIX(int nthreads, const KT *kin, KT *kout, size_t n, int sbit,       \
                                  int ebit) {                                                      \
    const KT *ke = kin + n;                                                                        \
    with_policy(nthreads, [&](auto &pol, auto) { radix_sort(pol, kin, ke, kout, sbit, ebit); });   \
  }
ZPCREF_SORT_PAIR(u32, uint32_t)
ZPCREF_SORT_PAIR(i32, int32_t)
ZPCREF_SORT_PAIR(u64, uint64_t)

#define ZPCREF_SCAN_REDUCE(SUFFIX, T)                                                             \
  void zpcref_exclusive_scan_sum_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads,                                                                          \
                [&](auto &pol, auto) { exclusive_scan(pol, in, e, out, (T)0, plus<T>{}); });  \
  }                                                                                                \
  void zpcref_inclusive_scan_sum_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads,                                                                          \
                [&](auto &pol, auto) { inclusive_scan(pol, in, e, out, plus<T>{}); });        \
  }                                                                                                \
  void zpcref_reduce_sum_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads,                                                                          \
                [&](auto &pol, auto) { reduce(pol, in, e, out, (T)0, plus<T>{}); });          \
  }                                                                                                \
  void zpcref_reduce_prod_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads,                                                                          \
                [&](auto &pol, auto) { reduce(pol, in, e, out, (T)1, multiplies<T>{}); });    \
  }                                                                                                \
  void zpcref_reduce_min_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads, [&](auto &pol, auto) {                                                   \
      reduce(pol, in, e, out, detail::deduce_numeric_max<T>(), getmin<T>{});                  \
    });                                                                                            \
  }                                                                                                \
  void zpcref_reduce_max_##SUFFIX(int nthreads, const T *in, T *out, size_t n) {\
    const T *e = in + n;\
    with_policy(nthreads, [&](auto &pol, auto) {                                                   \
      reduce(pol, in, e, out, detail::deduce_numeric_lowest<T>(), getmax<T>{});               \
    });                                                                                            \
  }
ZPCREF_SCAN_REDUCE(i32, int32_t)
ZPCREF_SCAN_REDUCE(f32, float)
ZPCREF_SCAN_REDUCE(u32, uint32_t)
ZPCREF_SCAN_REDUCE(i64, int64_t)
ZPCREF_SCAN_REDUCE(f64, double)

#define ZPCREF_MERGE_SORT_PAIR(SUFFIX, KT)                                                          \
  void zpcref_merge_sort_pair_##SUFFIX(int nthreads, KT *keys, int *vals, size_t n) {               \
    with_policy(nthreads, [&](auto &pol, auto) { merge_sort_pair(pol, keys, vals, (std::ptrdiff_t)n); }); \
  }
ZPCREF_MERGE_SORT_PAIR(i32, int32_t)
ZPCREF_MERGE_SORT_PAIR(f32, float)
ZPCREF_MERGE_SORT_PAIR(f64, double)

/// ---- per-particle math ----
void zpcref_svd3(const float *F, float *U, float *S, float *V) {  // column-major 9-vectors
  math::svd_3d(F[0], F[3], F[6], F[1], F[4], F[7], F[2], F[5], F[8], U[0], U[3], U[6], U[1], U[4],
               U[7], U[2], U[5], U[8], S[0], S[1], S[2], V[0], V[3], V[6], V[1], V[4], V[7], V[2],
               V[5], V[8]);
}
void zpcref_stress_fixedcorotated(float volume, float E, float nu, const float *F, float *PF) {
  const auto [mu, lambda] = lame_parameters(E, nu);
  vec<float, 9> f{}, pf{};
  for (int d = 0; d != 9; ++d) f[d] = F[d];
  compute_stress_fixedcorotated(volume, mu, lambda, f, pf);
  for (int d = 0; d != 9; ++d) PF[d] = pf[d];
}
void zpcref_stress_vonmises(float volume, float E, float nu, float yieldStress, const float *F, float *PF) {
  const auto [mu, lambda] = lame_parameters(E, nu);
  vec<float, 9> f{}, pf{};
  for (int d = 0; d != 9; ++d) f[d] = F[d];
  compute_stress_vonmisesfixedcorotated(volume, mu, lambda, yieldStress, f, pf);
  for (int d = 0; d != 9; ++d) PF[d] = pf[d];
}
void zpcref_stress_sand(float volume, float E, float nu, float cohesion, float beta, float yieldSurface,
                        int volumeCorrection, float *logJp, const float *F, float *PF) {
  const auto [mu, lambda] = lame_parameters(E, nu);
  vec<float, 9> f{}, pf{};
  for (int d = 0; d != 9; ++d) f[d] = F[d];
  compute_stress_sand(volume, mu, lambda, cohesion, beta, yieldSurface, volumeCorrection != 0, *logJp, f, pf);
  for (int d = 0; d != 9; ++d) PF[d] = pf[d];
}
void zpcref_stress_nacc(float volume, float E, float nu, float fa, float xi, float beta, int hardeningOn, float *logJp,
                        const float *F, float *PF) {
  const auto [mu, lambda] = lame_parameters(E, nu);
  NACCConfig model{};
  model.E = E;
  model.nu = nu;
  model.fa = fa;
  vec<float, 9> f{}, pf{};
  for (int d = 0; d != 9; ++d) f[d] = F[d];
  compute_stress_nacc(volume, mu, lambda, model.bulk(), xi, beta, model.Msqr(), hardeningOn != 0, *logJp, f, pf);
  for (int d = 0; d != 9; ++d) PF[d] = pf[d];
}
float zpcref_math_sqrt(float x) { return math::sqrt(x); }
void zpcref_nacc_consts(float E, float nu, float fa, float *bulk, float *msqr) {
  NACCConfig model{};
  model.E = E;
  model.nu = nu;
  model.fa = fa;
  *bulk = model.bulk();
  *msqr = model.Msqr();
}
int zpcref_max_threads() { return (int)std::thread::hardware_concurrency(); }
}

// ---------------------------------------------------------------------------------------------------------
// bht<i32,3,int,16> and SparseGrid<3,f32,8> (container/Bht.hpp, geometry/SparseGrid.hpp) — the reference's own
// containers and host views, used to pin oracle/sparse_oracle.c and to check tables built by the CUDA path.
// ---------------------------------------------------------------------------------------------------------
#include "zensim/container/Bht.hpp"
#include "zensim/geometry/SparseGrid.hpp"

namespace {
  using RefBht = bht<int, 3, int, 16>;
  using RefSg = SparseGrid<3, f32, 8>;
  using IKey = vec<int, 3>;
}  // namespace

extern "C" {

void *zpcref_bht_create(int expected) { return new RefBht{(size_t)expected, memsrc_e::host, -1}; }
void zpcref_bht_destroy(void *h) { delete (RefBht *)h; }
// info[0] = tableSize, info[1] = numBuckets, info[2] = cnt ; hf = {hf0.x, hf0.y, hf1.x, hf1.y, hf2.x, hf2.y}
void zpcref_bht_info(void *h, int *info, unsigned *hf) {
  auto &t = *(RefBht *)h;
  info[0] = (int)t._tableSize;
  info[1] = (int)(t._tableSize / RefBht::bucket_size);
  info[2] = (int)t.size();
  hf[0] = t._hf0._hashx; hf[1] = t._hf0._hashy; hf[2] = t._hf1._hashx; hf[3] = t._hf1._hashy;
  hf[4] = t._hf2._hashx; hf[5] = t._hf2._hashy;
}
// sequential host insert (BHTView::insert, Bht.hpp:609-664); out[i] = returned index (-1 when already present)
void zpcref_bht_insert(void *h, const int *keys, int n, int *out) {
  auto tv = proxy<execspace_e::host>(*(RefBht *)h);
  for (int i = 0; i < n; ++i) out[i] = tv.insert(IKey{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]});
}
void zpcref_bht_query(void *h, const int *keys, int n, int *out) {
  auto tv = proxy<execspace_e::host>(*(const RefBht *)h);
  for (int i = 0; i < n; ++i) out[i] = tv.query(IKey{keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]});
}
// raw table arrays: keys16 = tableSize x 4 ints (storage_key_type, 16-byte slots), indices, status, activeKeys (cnt x 3)
void zpcref_bht_get(void *h, int *keys16, int *indices, int *status, int *active_keys) {
  auto &t = *(RefBht *)h;
  static_assert(sizeof(typename RefBht::storage_key_type) == 16, "padded key slots");
  std::memcpy(keys16, t._table.keys.data(), (size_t)t._tableSize * 16);
  std::memcpy(indices, t._table.indices.data(), (size_t)t._tableSize * sizeof(int));
  std::memcpy(status, t._table.status.data(), (size_t)t._tableSize * sizeof(int));
  std::memcpy(active_keys, t._activeKeys.data(), (size_t)t.size() * 3 * sizeof(int));
}
// overwrite the container's arrays with externally built ones (e.g. downloaded from the GPU build): afterwards the
// UNMODIFIED BHTView::query runs on them (zpcref_bht_query)
void zpcref_bht_load(void *h, const int *keys16, const int *indices, const int *active_keys, int cnt) {
  auto &t = *(RefBht *)h;
  std::memcpy(t._table.keys.data(), keys16, (size_t)t._tableSize * 16);
  std::memcpy(t._table.indices.data(), indices, (size_t)t._tableSize * sizeof(int));
  std::memcpy(t._activeKeys.data(), active_keys, (size_t)cnt * 3 * sizeof(int));
  t._cnt.setVal(cnt);
}

void *zpcref_sg_create(int nblocks, int nch) { return new RefSg{(size_t)nch, (size_t)nblocks, memsrc_e::host, -1}; }
void zpcref_sg_destroy(void *h) { delete (RefSg *)h; }
void *zpcref_sg_table(void *h) { return &((RefSg *)h)->_table; }
void zpcref_sg_scale(void *h, float s) { ((RefSg *)h)->scale(s); }
void zpcref_sg_translate(void *h, const float *t) { ((RefSg *)h)->translate(vec<float, 3>{t[0], t[1], t[2]}); }
void zpcref_sg_get_transform(void *h, float *m16) {
  auto m = ((RefSg *)h)->getIndexToWorldTransformation();
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) m16[4 * i + j] = m(i, j);
}
void zpcref_sg_set_background(void *h, float b) { ((RefSg *)h)->_background = b; }
// grid = TileVector<f32,512>: tile b = [nch][512] floats
void zpcref_sg_load_grid(void *h, const float *data, int nblocks) {
  auto &g = *(RefSg *)h;
  std::memcpy(g._grid.data(), data, (size_t)nblocks * g.numChannels() * 512 * sizeof(float));
}
// SparseGridView::valueOr(false_c, chn, indexCoord, default) (SparseGrid.hpp:345-351) at n integer coordinates
void zpcref_sg_value_or(void *h, int chn, const int *coords, int n, float dflt, float *out) {
  auto sv = proxy<execspace_e::host>(*(const RefSg *)h);
  for (int i = 0; i < n; ++i) out[i] = sv.valueOr(false_c, chn, IKey{coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]}, dflt);
}
// iCoord(bno, cno) / wCoord(bno, cno) (SparseGrid.hpp:407-416) for n (block, cell) pairs
void zpcref_sg_coords(void *h, const int *bno, const int *cno, int n, int *icoord, float *wcoord) {
  auto sv = proxy<execspace_e::host>(*(const RefSg *)h);
  for (int i = 0; i < n; ++i) {
    auto ic = sv.iCoord((size_t)bno[i], cno[i]);
    auto wc = sv.wCoord((size_t)bno[i], cno[i]);
    for (int d = 0; d < 3; ++d) { icoord[3 * i + d] = ic[d]; wcoord[3 * i + d] = wc[d]; }
  }
}
void zpcref_sg_world_to_index(void *h, const float *x, int n, float *X) {
  auto sv = proxy<execspace_e::host>(*(const RefSg *)h);
  for (int i = 0; i < n; ++i) {
    auto r = sv.worldToIndex(vec<float, 3>{x[3 * i], x[3 * i + 1], x[3 * i + 2]});
    for (int d = 0; d < 3; ++d) X[3 * i + d] = r[d];
  }
}
}

// TileVector::reorderTiles (container/TileVector.hpp:641-691) on the SparseGrid's grid and bht::reorder
// (container/Bht.hpp:343-400), both with the reference's seq policy; scatter != 0 -> wrapv<true>
extern "C" {
void zpcref_sg_reorder_tiles(void *h, const int *map, int ntiles, int scatter, float *grid_out) {
  auto &g = *(RefSg *)h;
  Vector<int> m{(size_t)g._grid.numTiles(), memsrc_e::host, -1};
  for (size_t i = 0; i < m.size(); ++i) m[i] = i < (size_t)ntiles ? map[i] : (int)i;
  auto pol = seq_exec();
  if (scatter) g._grid.reorderTiles(pol, m, wrapv<true>{});
  else g._grid.reorderTiles(pol, m, wrapv<false>{});
  std::memcpy(grid_out, g._grid.data(), (size_t)ntiles * g.numChannels() * 512 * sizeof(float));
}
void zpcref_bht_reorder(void *h, const int *map, int scatter) {
  auto &t = *(RefBht *)h;
  Vector<int> m{(size_t)t.size(), memsrc_e::host, -1};
  for (size_t i = 0; i < m.size(); ++i) m[i] = map[i];
  auto pol = seq_exec();
  if (scatter) t.reorder(pol, m, wrapv<true>{});
  else t.reorder(pol, m, wrapv<false>{});
}
}

// ---------------------------------------------------------------------------------------------------------
// LBvh<3, int, f32> (container/Bvh.hpp): the reference's own build / refit on its host policies — pins oracle/bvh_oracle.c
// ---------------------------------------------------------------------------------------------------------
#include "zensim/container/Bvh.hpp"

extern "C" {
/// bvs: n boxes {min[3], max[3]}; outputs sized 2n-1 (n if n <= 2) / n as LBvh lays them out.  Returns numNodes.
int zpcref_lbvh_build(int nthreads, int n, const float *bvs, float *orderedBvs, int *auxIndices, int *parents, int *levels,
                      int *leafInds, int refit) {
  using Bvh = LBvh<3, int, float>;
  using Box = typename Bvh::Box;
  static_assert(sizeof(Box) == 6 * sizeof(float), "AABBBox<3,f32> is six floats");
  Vector<Box> prims{(size_t)n};
  std::memcpy((void *)prims.data(), bvs, sizeof(float) * 6 * n);
  Bvh bvh{};
  with_policy(nthreads, [&](auto &pol, auto) {
    if (refit)
      bvh.build(pol, prims, true_c);
    else
      bvh.build(pol, prims, false_c);
  });
  const int nn = (int)bvh.getNumNodes();
  std::memcpy(orderedBvs, (const void *)bvh.orderedBvs.data(), sizeof(float) * 6 * nn);
  std::memcpy(auxIndices, bvh.auxIndices.data(), sizeof(int) * bvh.auxIndices.size());
  if (n > 2) {
    std::memcpy(parents, bvh.parents.data(), sizeof(int) * nn);
    std::memcpy(levels, bvh.levels.data(), sizeof(int) * nn);
  }
  std::memcpy(leafInds, bvh.leafInds.data(), sizeof(int) * n);
  return nn;
}
/// build on the first set of boxes, refit with the second (same count), read orderedBvs back
void zpcref_lbvh_build_then_refit(int nthreads, int n, const float *bvs0, const float *bvs1, float *orderedBvs) {
  using Bvh = LBvh<3, int, float>;
  using Box = typename Bvh::Box;
  Vector<Box> a{(size_t)n}, b{(size_t)n};
  std::memcpy((void *)a.data(), bvs0, sizeof(float) * 6 * n);
  std::memcpy((void *)b.data(), bvs1, sizeof(float) * 6 * n);
  Bvh bvh{};
  with_policy(nthreads, [&](auto &pol, auto) {
    bvh.build(pol, a, true_c);
    bvh.refit(pol, b);
  });
  std::memcpy(orderedBvs, (const void *)bvh.orderedBvs.data(), sizeof(float) * 6 * bvh.getNumNodes());
}
}

// ---------------------------------------------------------------------------------------------------------
// index_buckets_for_particles (simulation/particle/Query.tpp:9-58): the reference's functors in the reference's order on
// its host policies.  (Query.tpp itself is only instantiated for GeneralParticles inside the un-built simulator library;
// the body below is that function's sequence, functor for functor.)
// ---------------------------------------------------------------------------------------------------------
#include "zensim/container/IndexBuckets.hpp"

extern "C" {
int zpcref_index_buckets(int nthreads, int n, const float *x, float dx, float displacement, int *activeKeys, int *counts,
                         int *offsets, int *indices) {
  using ib_t = IndexBuckets<3, i32, int>;
  using vector_t = typename ib_t::vector_t;
  Vector<vec<float, 3>> pos{(size_t)n};
  std::memcpy((void *)pos.data(), x, sizeof(float) * 3 * n);
  ib_t ib{};
  ib._dx = dx;
  auto &table = ib._table;
  table = RM_CVREF_T(table){pos.get_allocator(), (size_t)n};
  int numCells = 0;
  with_policy(nthreads, [&](auto &pol, auto tag) {
    pol(range(table._tableSize), CleanSparsity{tag, table});
    pol(range(n), ComputeSparsity{tag, dx, 1, table, pos, 0, displacement});
    numCells = table.size() + 1;
    ib._counts = vector_t{pos.get_allocator(), (size_t)numCells};
    std::memset(ib._counts.data(), 0, sizeof(int) * numCells);
    auto tmp = ib._counts;
    pol(range(n), SpatiallyCount{tag, dx, table, pos, ib._counts, 1, 0, displacement});
    ib._offsets = vector_t{pos.get_allocator(), (size_t)numCells};
    exclusive_scan(pol, ib._counts.begin(), ib._counts.end(), ib._offsets.begin());
    ib._indices = vector_t{pos.get_allocator(), (size_t)n};
    pol(range(n), SpatiallyDistribute{tag, dx, table, pos, tmp, ib._offsets, ib._indices, 1, 0, displacement});
  });
  std::memcpy(activeKeys, table._activeKeys.data(), sizeof(int) * 3 * (numCells - 1));
  std::memcpy(counts, ib._counts.data(), sizeof(int) * numCells);
  std::memcpy(offsets, ib._offsets.data(), sizeof(int) * numCells);
  std::memcpy(indices, ib._indices.data(), sizeof(int) * n);
  return numCells - 1;
}
}

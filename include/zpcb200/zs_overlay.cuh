// zs_overlay.cuh — the binding INTEGRATION.md describes, as code that compiles against the UNMODIFIED reference headers
// (include it after the zensim headers of the path; link libzpcb200.so).
//
//   zs::B200ExecutionPolicy / zs::b200_exec()   a CudaExecutionPolicy whose reduce / exclusive_scan / inclusive_scan /
//       radix_sort / radix_sort_pair members go to libzpcb200 when the iterators are raw pointers or zs::Vector iterators and
//       the operator is plus (reduce: plus / multiplies / getmin / getmax) on {i32, u32, i64, f32, f64} — every other call, and every
//       policy(range, functor) launch, stays the reference's (members are hidden, not removed).  Generic code that is templated
//       on the policy (zs::radix_sort_pair(pol, ...), LBvh::build(pol, ...), ...) needs no change: pass b200_exec().
//   zs::b200::partition_for_particles / clean_grid_blocks / p2g / compute_grid_block_velocity / g2p
//       the MPM functor launches of SURVEY §3.1 on the reference's own containers (Particles, HashTable, Grids):
//       pol(range(n), P2GTransfer{cuda_c, wrapv<apic>{}, dt, model, pars, table, grids})  becomes  b200::p2g(pol, dt, model, pars, table, grids).
//   zs::b200::BinnedParticles + bin_particles / rebin_particles / p2g / g2p   the block-binned fast path on TileVector<f32, 32> (the type of
//       Particles::particleBins) with zs::Vector metadata.
//
// Temporary storage comes from the policy's stream-ordered pool (streamMemAlloc / streamMemFree, like
// cuda/execution/ExecutionPolicy.cuh:806-815); errors surface like the reference's (checkCuApiError is private to Cuda, so a
// std::runtime_error carries the code); the policy's sync(true) default is honoured.
#pragma once
#include <atomic>
#include <stdexcept>
#include <string>

#include "zensim/container/HashTable.hpp"
#include "zensim/container/TileVector.hpp"
#include "zensim/container/Vector.hpp"
#include "zensim/cuda/execution/ExecutionPolicy.cuh"
#include "zensim/geometry/AnalyticLevelSet.h"
#include "zensim/geometry/Collider.h"
#include "zensim/geometry/SparseGrid.hpp"
#include "zensim/geometry/Structure.hpp"
#include "zensim/geometry/Structurefree.hpp"
#include "zensim/physics/ConstitutiveModel.hpp"
#include "zensim/py_interop/GenericIterator.hpp"
#include "zpcb200.h"

namespace zs {

  namespace b200_detail {
    /// iterator -> zpc_port: raw pointers, zs::Vector iterators (contiguous) and the reference's aosoa_iterator<T, 1> — TileVector
    /// channels as its own C layer passes them (py_interop/GenericIterator.hpp:16-98; the port struct is the same five fields)
    template <class It> struct raw_iter {
      using I = remove_cvref_t<It>;
      using value_type = remove_cv_t<typename std::iterator_traits<I>::value_type>;
      static constexpr bool is_vector_iter
          = is_same_v<I, decltype(zs::begin(declval<Vector<value_type> &>()))> || is_same_v<I, decltype(zs::begin(declval<const Vector<value_type> &>()))>;
      static constexpr bool is_aosoa = std::is_base_of_v<aosoa_iterator<value_type, 1>, I> || std::is_base_of_v<aosoa_iterator<const value_type, 1>, I>;
      static constexpr bool ok = std::is_pointer_v<I> || is_vector_iter || is_aosoa;
      static zpc_port get(I it) {
        if constexpr (std::is_pointer_v<I>) return zpc_port{(void *)it, 0, 0, 0, 1};
        else if constexpr (is_aosoa) return zpc_port{(void *)it.base, it.idx, it.numTileBits, it.tileMask, it.numChns};
        else return zpc_port{(void *)it.operator->(), 0, 0, 0, 1};
      }
    };
    template <class T> constexpr int kind_of() {  // index into the per-type entry tables below, -1 = unsupported
      if constexpr (is_same_v<T, int>) return 0;
      else if constexpr (is_same_v<T, unsigned>) return 1;
      else if constexpr (is_same_v<T, long long> || is_same_v<T, long>) return sizeof(T) == 8 ? 2 : -1;
      else if constexpr (is_same_v<T, float>) return 3;
      else if constexpr (is_same_v<T, double>) return 4;
      else return -1;
    }
    using prim2_t = int (*)(void *, size_t *, zpc_port, zpc_port, size_t, zpc_stream_t);
    inline prim2_t reduce_entry(int op, int kind) {
      static const prim2_t t[4][5] = {
          {zpcb200_reduce_sum_i32, zpcb200_reduce_sum_u32, zpcb200_reduce_sum_i64, zpcb200_reduce_sum_f32, zpcb200_reduce_sum_f64},
          {zpcb200_reduce_min_i32, zpcb200_reduce_min_u32, zpcb200_reduce_min_i64, zpcb200_reduce_min_f32, zpcb200_reduce_min_f64},
          {zpcb200_reduce_max_i32, zpcb200_reduce_max_u32, zpcb200_reduce_max_i64, zpcb200_reduce_max_f32, zpcb200_reduce_max_f64},
          {zpcb200_reduce_prod_i32, zpcb200_reduce_prod_u32, zpcb200_reduce_prod_i64, zpcb200_reduce_prod_f32, zpcb200_reduce_prod_f64}};
      return t[op][kind];
    }
    inline prim2_t scan_entry(bool inclusive, int kind) {
      static const prim2_t t[2][5] = {{zpcb200_exclusive_scan_sum_i32, zpcb200_exclusive_scan_sum_u32, zpcb200_exclusive_scan_sum_i64,
                                       zpcb200_exclusive_scan_sum_f32, zpcb200_exclusive_scan_sum_f64},
                                      {zpcb200_inclusive_scan_sum_i32, zpcb200_inclusive_scan_sum_u32, zpcb200_inclusive_scan_sum_i64,
                                       zpcb200_inclusive_scan_sum_f32, zpcb200_inclusive_scan_sum_f64}};
      return t[inclusive ? 1 : 0][kind];
    }
    template <class Op, class T> constexpr int reduce_op() {  // 0 plus, 1 min, 2 max, 3 multiplies, -1 other
      using O = remove_cvref_t<Op>;
      if constexpr (is_same_v<O, plus<T>> || is_same_v<O, plus<void>>) return 0;
      else if constexpr (is_same_v<O, getmin<T>> || is_same_v<O, getmin<void>>) return 1;
      else if constexpr (is_same_v<O, getmax<T>> || is_same_v<O, getmax<void>>) return 2;
      else if constexpr (is_same_v<O, multiplies<T>> || is_same_v<O, multiplies<void>>) return 3;
      else return -1;
    }
    inline zpc_port port_of(zpc_port p) { return p; }
  }  // namespace b200_detail

  struct B200ExecutionPolicy : CudaExecutionPolicy {
    using Base = CudaExecutionPolicy;
    using Base::operator();
    B200ExecutionPolicy() = default;
    B200ExecutionPolicy(const CudaExecutionPolicy &p) : CudaExecutionPolicy{p} {}

    void *b200Stream() const {
      auto &ctx = Cuda::context(getProcid());
      ctx.setContext();
      return ctx.streamSpare(getStreamid());
    }
    void b200Done(int rc, const char *what, const source_location &loc) const {
      if (rc) throw std::runtime_error(std::string("[") + what + "] failed with code " + std::to_string(rc));
      if (this->shouldSync()) Cuda::context(getProcid()).syncStreamSpare(getStreamid(), loc);
    }
    /// temp == nullptr size query, allocation from the stream-ordered pool, the call, release on the same stream
    /// The reference releases the stream-ordered pool at every synchronisation (release threshold 0): each primitive call then pays a
    /// fresh cuMemCreate / map for its scratch — measured 2.5 ms per call on a B200 whatever the size (r02_prims_vs_refcuda).  The
    /// overlay keeps the device's default pool warm instead (once per device; memory stays reserved for later scratch requests).
    static void b200KeepPoolWarm(int dev) {
      static std::atomic<unsigned long long> done{0};
      const unsigned long long bit = 1ull << (dev & 63);
      if (done.load(std::memory_order_acquire) & bit) return;
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      done.fetch_or(bit, std::memory_order_release);
    }
    template <class Fn, class... Args> void b200TwoPhase(const char *what, const source_location &loc, Fn fn, Args... args) const {
      auto &ctx = Cuda::context(getProcid());
      b200KeepPoolWarm(getProcid());
      void *stream = b200Stream();
      size_t bytes = 0;
      int rc = fn(nullptr, &bytes, args..., nullptr);
      if (rc) b200Done(rc, what, loc);
      void *tmp = ctx.streamMemAlloc(bytes ? bytes : 256, stream, loc);
      rc = fn(tmp, &bytes, args..., stream);
      ctx.streamMemFree(tmp, stream, loc);
      b200Done(rc, what, loc);
    }

    template <class InputIt, class OutputIt, class BinaryOp = plus<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>>
    void reduce(InputIt &&first, InputIt &&last, OutputIt &&d_first,
                typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type init
                = deduce_identity<BinaryOp, typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>(),
                BinaryOp &&binary_op = {}, const source_location &loc = source_location::current()) const {
      using T = remove_cv_t<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>;
      constexpr int kind = b200_detail::kind_of<T>(), op = b200_detail::reduce_op<BinaryOp, T>();
      if constexpr (kind >= 0 && op >= 0 && b200_detail::raw_iter<InputIt>::ok && b200_detail::raw_iter<OutputIt>::ok) {
        // the library reduces from the operator's identity, which is what every in-tree caller passes (deduce_identity)
        if (init == deduce_identity<remove_cvref_t<BinaryOp>, T>()) {
          const auto n = (size_t)(last - first);
          b200TwoPhase("zpcb200_reduce", loc, b200_detail::reduce_entry(op, kind), b200_detail::port_of(b200_detail::raw_iter<InputIt>::get(first)),
                       b200_detail::port_of(b200_detail::raw_iter<OutputIt>::get(d_first)), n);
          return;
        }
      }
      Base::reduce(FWD(first), FWD(last), FWD(d_first), init, FWD(binary_op), loc);
    }
    template <class InputIt, class OutputIt, class BinaryOperation = plus<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>>
    void exclusive_scan(InputIt &&first, InputIt &&last, OutputIt &&d_first,
                        typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type init
                        = deduce_identity<BinaryOperation, typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>(),
                        BinaryOperation &&binary_op = {}, const source_location &loc = source_location::current()) const {
      using T = remove_cv_t<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>;
      constexpr int kind = b200_detail::kind_of<T>();
      if constexpr (kind >= 0 && b200_detail::reduce_op<BinaryOperation, T>() == 0 && b200_detail::raw_iter<InputIt>::ok
                    && b200_detail::raw_iter<OutputIt>::ok) {
        if (init == T{}) {
          b200TwoPhase("zpcb200_exclusive_scan_sum", loc, b200_detail::scan_entry(false, kind),
                       b200_detail::port_of(b200_detail::raw_iter<InputIt>::get(first)),
                       b200_detail::port_of(b200_detail::raw_iter<OutputIt>::get(d_first)), (size_t)(last - first));
          return;
        }
      }
      Base::exclusive_scan(FWD(first), FWD(last), FWD(d_first), init, FWD(binary_op), loc);
    }
    template <class InputIt, class OutputIt, class BinaryOperation = plus<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>>
    void inclusive_scan(InputIt &&first, InputIt &&last, OutputIt &&d_first, BinaryOperation &&binary_op = {},
                        const source_location &loc = source_location::current()) const {
      using T = remove_cv_t<typename std::iterator_traits<remove_cvref_t<InputIt>>::value_type>;
      constexpr int kind = b200_detail::kind_of<T>();
      if constexpr (kind >= 0 && b200_detail::reduce_op<BinaryOperation, T>() == 0 && b200_detail::raw_iter<InputIt>::ok
                    && b200_detail::raw_iter<OutputIt>::ok) {
        b200TwoPhase("zpcb200_inclusive_scan_sum", loc, b200_detail::scan_entry(true, kind),
                     b200_detail::port_of(b200_detail::raw_iter<InputIt>::get(first)),
                     b200_detail::port_of(b200_detail::raw_iter<OutputIt>::get(d_first)), (size_t)(last - first));
      } else
        Base::inclusive_scan(FWD(first), FWD(last), FWD(d_first), FWD(binary_op), loc);
    }
    template <class KeyIter, class ValueIter,
              typename Tn = typename std::iterator_traits<remove_reference_t<KeyIter>>::difference_type>
    enable_if_type<is_ra_iter_v<remove_reference_t<KeyIter>> && is_ra_iter_v<remove_reference_t<ValueIter>>> radix_sort_pair(
        KeyIter &&keysIn, ValueIter &&valsIn, KeyIter &&keysOut, ValueIter &&valsOut, Tn count = 0, int sbit = 0,
        int ebit = sizeof(typename std::iterator_traits<remove_reference_t<KeyIter>>::value_type) * 8,
        const source_location &loc = source_location::current()) const {
      using K = remove_cv_t<typename std::iterator_traits<remove_cvref_t<KeyIter>>::value_type>;
      using V = remove_cv_t<typename std::iterator_traits<remove_cvref_t<ValueIter>>::value_type>;
      constexpr bool keyOk = is_same_v<K, unsigned> || is_same_v<K, int> || (std::is_unsigned_v<K> && sizeof(K) == 8);
      if constexpr (keyOk && is_same_v<V, int> && b200_detail::raw_iter<KeyIter>::ok && b200_detail::raw_iter<ValueIter>::ok) {
        auto fn = is_same_v<K, unsigned> ? zpcb200_radix_sort_pair_u32 : (is_same_v<K, int> ? zpcb200_radix_sort_pair_i32 : zpcb200_radix_sort_pair_u64);
        b200TwoPhase("zpcb200_radix_sort_pair", loc, fn, b200_detail::port_of(b200_detail::raw_iter<KeyIter>::get(keysIn)),
                     b200_detail::port_of(b200_detail::raw_iter<ValueIter>::get(valsIn)),
                     b200_detail::port_of(b200_detail::raw_iter<KeyIter>::get(keysOut)),
                     b200_detail::port_of(b200_detail::raw_iter<ValueIter>::get(valsOut)), (size_t)count, sbit, ebit);
      } else
        Base::radix_sort_pair(FWD(keysIn), FWD(valsIn), FWD(keysOut), FWD(valsOut), count, sbit, ebit, loc);
    }
  };
  inline B200ExecutionPolicy b200_exec() noexcept { return B200ExecutionPolicy{}; }

  /// the MPM functor launches on the reference's own containers (device memory)
  namespace b200 {
    inline zpc_particles_view view(Particles<f32, 3> &pars) {
      auto addr = [&](const char *name) { return pars.hasAttr(name) ? (float *)pars.getAttrAddress(name) : (float *)nullptr; };
      return zpc_particles_view{addr("m"), addr("x"), addr("v"), nullptr, addr("J"), addr("F"), addr("C"), addr("logJp"), pars.size()};
    }
    inline zpc_hashtable_view view(HashTable<i32, 3, int> &t) {
      return zpc_hashtable_view{(int *)t._table.keys.data(), t._table.indices.data(), t._table.status.data(), (int *)t._activeKeys.data(),
                                (int)t._tableSize, t._cnt.data()};
    }
    inline zpc_grids_view view(Grids<f32, 3, 4> &g) {
      auto &blocks = g.grid(collocated_c).blocks;
      return zpc_grids_view{(float *)blocks.data(), blocks.size() / 64, 7, g._dx};
    }
    /// partition_for_particles: CleanSparsity + ComputeSparsity + EnlargeSparsity{0, 2} (simulation/sparsity/SparsityOp.hpp:41-112)
    inline void partition_for_particles(const B200ExecutionPolicy &pol, HashTable<i32, 3, int> &table, Particles<f32, 3> &pars, float dx,
                                        const source_location &loc = source_location::current()) {
      zpc_port x{(void *)pars.getAttrAddress("x"), 0, 0, 0, 3};
      pol.b200TwoPhase("zpcb200_partition_build", loc, zpcb200_partition_build, x, (size_t)pars.size(), dx, view(table), 0, 2, (int *)nullptr);
    }
    inline void clean_grid_blocks(const B200ExecutionPolicy &pol, HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids,
                                  const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_clean_grid(view(grids), table._cnt.data(), pol.b200Stream()), "zpcb200_clean_grid", loc);
    }
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const FixedCorotatedConfig &model, Particles<f32, 3> &pars,
                    HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
      pol.b200Done(zpcb200_p2g_apic_fcr(view(pars), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_fcr", loc);
    }
    /// the other constitutive models of P2GTransfer (P2G.hpp:66-102): same call, the model type selects the entry
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const VonMisesFixedCorotatedConfig &model, Particles<f32, 3> &pars,
                    HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_vonmises_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu, model.yieldStress};
      pol.b200Done(zpcb200_p2g_apic_vonmises(view(pars), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_vonmises", loc);
    }
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const EquationOfStateConfig &model, Particles<f32, 3> &pars,
                    HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_equation_of_state m{model.rho, model.volume, model.dim, model.bulk, model.gamma, model.viscosity};
      pol.b200Done(zpcb200_p2g_apic_eos(view(pars), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_eos", loc);
    }
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const DruckerPragerConfig &model, Particles<f32, 3> &pars,
                    HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_drucker_prager m{model.rho, model.volume, model.dim, model.E, model.nu, model.logJp0, model.fa, model.cohesion, model.beta,
                           model.volumeCorrection ? 1 : 0, model.yieldSurface};
      pol.b200Done(zpcb200_p2g_apic_drucker_prager(view(pars), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_drucker_prager", loc);
    }
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const NACCConfig &model, Particles<f32, 3> &pars, HashTable<i32, 3, int> &table,
                    Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_nacc m{model.rho, model.volume, model.dim, model.E, model.nu, model.logJp0, model.fa, model.xi, model.beta, model.hardeningOn ? 1 : 0};
      pol.b200Done(zpcb200_p2g_apic_nacc(view(pars), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_nacc", loc);
    }
    /// G2PTransfer with EquationOfStateConfig (G2P.hpp:69-73): J instead of F
    inline void g2p(const B200ExecutionPolicy &pol, float dt, const EquationOfStateConfig &, Grids<f32, 3, 4> &grids, HashTable<i32, 3, int> &table,
                    Particles<f32, 3> &pars, const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_g2p_apic_eos(view(pars), view(table), view(grids), dt, pol.b200Stream()), "zpcb200_g2p_apic_eos", loc);
    }
    /// Collider<AnalyticLevelSet<Plane | Sphere | Cuboid>> with its rigid motion -> zpc_collider (geometry/Collider.h:10-24, 136-143)
    template <analytic_geometry_e geom> zpc_collider view(const Collider<AnalyticLevelSet<geom, f32, 3>> &col) {
      static_assert(geom == analytic_geometry_e::Plane || geom == analytic_geometry_e::Sphere || geom == analytic_geometry_e::Cuboid,
                    "plane, sphere and cuboid colliders are built");
      zpc_collider c{};
      c.type = col.type == collider_e::Sticky ? ZPC_COLLIDER_STICKY : (col.type == collider_e::Slip ? ZPC_COLLIDER_SLIP : ZPC_COLLIDER_SEPARATE);
      for (int d = 0; d < 3; ++d) {
        if constexpr (geom == analytic_geometry_e::Plane) { c.geometry = ZPC_GEOM_PLANE; c.origin[d] = col.levelset._origin[d]; c.normal[d] = col.levelset._normal[d]; }
        else if constexpr (geom == analytic_geometry_e::Sphere) { c.geometry = ZPC_GEOM_SPHERE; c.origin[d] = col.levelset._center[d]; c.normal[d] = d == 0 ? col.levelset._radius : 0.f; }
        else { c.geometry = ZPC_GEOM_CUBOID; c.origin[d] = col.levelset._min[d]; c.normal[d] = col.levelset._max[d]; }
        c.b[d] = col.b[d];
        c.dbdt[d] = col.dbdt[d];
        c.omega[d] = col.omega.omega[d];
        for (int e = 0; e < 3; ++e) c.R[3 * d + e] = col.R(d, e);
      }
      c.s = col.s;
      c.dsdt = col.dsdt;
      return c;
    }
    /// pol(Collapse{nb, 64}, ApplyBoundaryConditionOnGridBlocks{cuda_c, collider, table, grids}) (GridOp.hpp:112-164)
    template <analytic_geometry_e geom>
    void apply_boundary_condition(const B200ExecutionPolicy &pol, const Collider<AnalyticLevelSet<geom, f32, 3>> &col, HashTable<i32, 3, int> &table,
                                  Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_apply_boundary(view(grids), view(table), view(col), pol.b200Stream()), "zpcb200_apply_boundary", loc);
    }
    /// ComputeGridBlockVelocity{cuda_c, wrapv<apic>{}, grids, dt, gravity, maxVel}; mode 1 adds rhs first (explicit update)
    inline void compute_grid_block_velocity(const B200ExecutionPolicy &pol, Grids<f32, 3, 4> &grids, HashTable<i32, 3, int> &table, float dt,
                                            float gravity, float *maxVel, int mode = 0, const source_location &loc = source_location::current()) {
      const float extf[3] = {0.f, gravity, 0.f};
      pol.b200Done(zpcb200_grid_update(view(grids), table._cnt.data(), dt, extf, mode, maxVel, pol.b200Stream()), "zpcb200_grid_update", loc);
    }
    /// GridMomentumToVelocity{cuda_c, grids.grid(collocated_c), mChn, mvChn, maxVel} (GridOp.hpp:184-214)
    inline void grid_momentum_to_velocity(const B200ExecutionPolicy &pol, Grids<f32, 3, 4> &grids, HashTable<i32, 3, int> &table, float *maxVel,
                                          int mChn = 0, int mvChn = 1, const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_grid_momentum_to_velocity(view(grids), table._cnt.data(), mChn, mvChn, maxVel, pol.b200Stream()),
                   "zpcb200_grid_momentum_to_velocity", loc);
    }
    /// GridAngularMomentum{cuda_c, table, grids.grid(collocated_c), mChn, mvChn, sum} (GridOp.hpp:216-262); sum: six doubles, added to
    inline void grid_angular_momentum(const B200ExecutionPolicy &pol, HashTable<i32, 3, int> &table, Grids<f32, 3, 4> &grids, double *sum,
                                      int mChn = 0, int mvChn = 1, const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_grid_angular_momentum(view(grids), view(table), mChn, mvChn, sum, pol.b200Stream()),
                   "zpcb200_grid_angular_momentum", loc);
    }
    inline void g2p(const B200ExecutionPolicy &pol, float dt, Grids<f32, 3, 4> &grids, HashTable<i32, 3, int> &table, Particles<f32, 3> &pars,
                    const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_g2p_apic(view(pars), view(table), view(grids), dt, pol.b200Stream()), "zpcb200_g2p_apic", loc);
    }

    /// ---- the SparseGrid<3, f32, 8> seam (geometry/SparseGrid.hpp:16-188): bht keyed by block origins + TileVector<f32, 512> ----
    inline zpc_sparsegrid_view view(SparseGrid<3, f32, 8> &sg) {
      auto &t = sg._table;
      zpc_sparsegrid_view v{};
      v.table = zpc_bht_view{(int *)t._table.keys.data(), t._table.indices.data(), t._table.status.data(), (int *)t._activeKeys.data(),
                             (uint32_t)t._tableSize, (uint32_t)(t._tableSize / 16), t._cnt.data(), t._buildSuccess.data(),
                             {t._hf0._hashx, t._hf0._hashy, t._hf1._hashx, t._hf1._hashy, t._hf2._hashx, t._hf2._hashy}};
      v.grid = (float *)sg._grid.data();
      v.numBlocks = sg._grid.numTiles();
      v.numChannels = (int)sg.numChannels();
      const auto m = sg.getIndexToWorldTransformation();
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) v.transform[4 * i + j] = m(i, j);
      v.background = sg._background;
      return v;
    }
    inline void sg_partition_for_particles(const B200ExecutionPolicy &pol, SparseGrid<3, f32, 8> &sg, Particles<f32, 3> &pars,
                                           const source_location &loc = source_location::current()) {
      zpc_port x{(void *)pars.getAttrAddress("x"), 0, 0, 0, 3};
      pol.b200TwoPhase("zpcb200_sg_partition_build", loc, zpcb200_sg_partition_build, x, (size_t)pars.size(), view(sg), 0, 2, (int *)nullptr);
    }
    inline void sg_clean(const B200ExecutionPolicy &pol, SparseGrid<3, f32, 8> &sg, const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_sg_clean(view(sg), pol.b200Stream()), "zpcb200_sg_clean", loc);
    }
    inline void sg_p2g(const B200ExecutionPolicy &pol, float dt, const FixedCorotatedConfig &model, Particles<f32, 3> &pars, SparseGrid<3, f32, 8> &sg,
                       const source_location &loc = source_location::current()) {
      zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
      pol.b200Done(zpcb200_sg_p2g_apic_fcr(view(pars), view(sg), dt, m, pol.b200Stream()), "zpcb200_sg_p2g_apic_fcr", loc);
    }
    inline void sg_grid_update(const B200ExecutionPolicy &pol, SparseGrid<3, f32, 8> &sg, float dt, float gravity, float *maxVel, int mode = 0,
                               const source_location &loc = source_location::current()) {
      const float extf[3] = {0.f, gravity, 0.f};
      pol.b200Done(zpcb200_sg_grid_update(view(sg), dt, extf, mode, maxVel, pol.b200Stream()), "zpcb200_sg_grid_update", loc);
    }
    inline void sg_g2p(const B200ExecutionPolicy &pol, float dt, SparseGrid<3, f32, 8> &sg, Particles<f32, 3> &pars,
                       const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_sg_g2p_apic(view(pars), view(sg), dt, pol.b200Stream()), "zpcb200_sg_g2p_apic", loc);
    }

    /// ---- the block-binned fast path on the reference's containers ------------------------------------------------------------
    /// Particles binned by home block live in TileVector<f32, 32> — the type of Particles::particleBins (geometry/Structurefree.hpp:220) —
    /// with the 25 channels {m, x, v, C, F}; bin metadata and the cell-order cache are zs::Vectors.  Two particle buffers: re-binning
    /// ping-pongs between them.  Usage (one substep): partition_for_particles(pol, table, bins) [every rebinEvery substeps, followed by
    /// rebin], clean_grid_blocks, p2g, compute_grid_block_velocity, g2p.
    struct BinnedParticles {
      using TV = TileVector<f32, 32>;
      TV cur, alt;
      Vector<int> binStart, binKey, numBins, cellOrderValid, order;
      Vector<unsigned short> cellOrder, cellStart;
      size_t n;
      int binCapacity;
      BinnedParticles(size_t n_, int binCapacity_, ProcID dev = 0)
          : cur{{{"m", 1}, {"x", 3}, {"v", 3}, {"C", 9}, {"F", 9}}, n_, memsrc_e::device, dev},
            alt{{{"m", 1}, {"x", 3}, {"v", 3}, {"C", 9}, {"F", 9}}, n_, memsrc_e::device, dev},
            binStart{(size_t)binCapacity_ + 1, memsrc_e::device, dev},
            binKey{(size_t)binCapacity_ * 3, memsrc_e::device, dev},
            numBins{1, memsrc_e::device, dev},
            cellOrderValid{1, memsrc_e::device, dev},
            order{n_ ? n_ : 1, memsrc_e::device, dev},
            cellOrder{n_ ? n_ : 1, memsrc_e::device, dev},
            cellStart{(size_t)binCapacity_ * ZPCB200_CELL_GROUPS_PAD, memsrc_e::device, dev},
            n{n_},
            binCapacity{binCapacity_} {
        static_assert(ZPC_PB_M == 0 && ZPC_PB_X == 1 && ZPC_PB_V == 4 && ZPC_PB_C == 7 && ZPC_PB_F == 16 && ZPC_PB_NCH == 25, "channel order");
        cellOrderValid.setVal(0);
      }
      zpc_bins_view view(TV &tv) {
        return zpc_bins_view{zpc_tilevector_view{(float *)tv.data(), n, (int)tv.numChannels()}, binStart.data(), binKey.data(), numBins.data(),
                             binCapacity, cellOrder.data(), cellStart.data(), cellOrderValid.data()};
      }
      zpc_bins_view view() { return view(cur); }
      /// positions as an iterator port over channel "x" of the current buffer (for the partition build)
      zpc_port xPort() { return zpc_port{(void *)((float *)cur.data() + ZPC_PB_X * 32), 0, 5, 31, ZPC_PB_NCH}; }
    };
    /// AoS Particles -> bins (stable sort by home block and cell, AoSoA gather); order[i] = source particle of slot i
    inline void bin_particles(const B200ExecutionPolicy &pol, Particles<f32, 3> &pars, HashTable<i32, 3, int> &table, float dx, BinnedParticles &bins,
                              const source_location &loc = source_location::current()) {
      pol.b200TwoPhase("zpcb200_bin_particles", loc, zpcb200_bin_particles, view(pars), view(table), dx, bins.view(), bins.order.data());
    }
    /// re-bin after the particles have moved (and the partition has been rebuilt): cur -> alt, swap
    inline void rebin_particles(const B200ExecutionPolicy &pol, HashTable<i32, 3, int> &table, float dx, BinnedParticles &bins,
                                const source_location &loc = source_location::current()) {
      pol.b200TwoPhase("zpcb200_rebin_particles", loc, zpcb200_rebin_particles, bins.view(bins.cur), view(table), dx, bins.view(bins.alt));
      std::swap(bins.cur, bins.alt);
    }
    inline void unbin_particles(const B200ExecutionPolicy &pol, BinnedParticles &bins, Particles<f32, 3> &pars,
                                const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_unbin_particles(bins.view(), view(pars), pol.b200Stream()), "zpcb200_unbin_particles", loc);
    }
    inline void partition_for_particles(const B200ExecutionPolicy &pol, HashTable<i32, 3, int> &table, BinnedParticles &bins, float dx,
                                        int enlargeLo = 0, int enlargeHi = 2, const source_location &loc = source_location::current()) {
      pol.b200TwoPhase("zpcb200_partition_build", loc, zpcb200_partition_build, bins.xPort(), bins.n, dx, view(table), enlargeLo, enlargeHi, (int *)nullptr);
    }
    inline void p2g(const B200ExecutionPolicy &pol, float dt, const FixedCorotatedConfig &model, BinnedParticles &bins, HashTable<i32, 3, int> &table,
                    Grids<f32, 3, 4> &grids, const source_location &loc = source_location::current()) {
      zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
      pol.b200Done(zpcb200_p2g_apic_fcr_binned(bins.view(), view(table), view(grids), dt, m, pol.b200Stream()), "zpcb200_p2g_apic_fcr_binned", loc);
    }
    inline void g2p(const B200ExecutionPolicy &pol, float dt, Grids<f32, 3, 4> &grids, HashTable<i32, 3, int> &table, BinnedParticles &bins,
                    const source_location &loc = source_location::current()) {
      pol.b200Done(zpcb200_g2p_apic_binned(bins.view(), view(table), view(grids), dt, pol.b200Stream()), "zpcb200_g2p_apic_binned", loc);
    }
  }  // namespace b200
}  // namespace zs

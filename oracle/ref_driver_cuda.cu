// TEST / BASELINE INFRASTRUCTURE — not product code.
//
// The reference's OWN CUDA path of the MPM substep ("pick the CUDA reference as primary for GPU parity"):
// the unmodified reference headers and its CUDA backend sources, compiled in place from /root/reference by oracle/Makefile
// (target `refcuda`) for sm_100 into oracle/_ref/libzpcref_cuda.so.  Containers live on memsrc_e::device, functors run on
// cuda_exec() — CleanSparsity / ComputeSparsity / EnlargeSparsity{0,2}, CleanGridBlocks, P2GTransfer<apic, FixedCorotated>,
// ComputeGridBlockVelocity (preceded by the "mv += rhs" lambda for the explicit update, like oracle/ref_driver.cpp),
// G2PTransfer — exactly the composed step of SURVEY §3.1.  Used by tests (parity of libzpcb200.so against the reference's
// device arithmetic) and by bench.py --impl reference-cuda (informational GPU-vs-GPU baseline).  Needs a GPU to run: the
// build container can only compile it.
#include <chrono>
#include <cstring>
#include <vector>

#include "zensim/container/HashTable.hpp"
#include "zensim/container/Vector.hpp"
#include "zensim/cuda/execution/ExecutionPolicy.cuh"
#include "zensim/execution/ExecutionPolicy.hpp"
#include "zensim/geometry/AnalyticLevelSet.h"
#include "zensim/geometry/Collider.h"
#include "zensim/geometry/SparseLevelSet.hpp"
#include "zensim/geometry/Structure.hpp"
#include "zensim/geometry/Structurefree.hpp"
#include "zensim/physics/ConstitutiveModel_Vol_dP.hpp"
#include "zensim/simulation/grid/GridOp.hpp"
#include "zensim/simulation/sparsity/SparsityOp.hpp"
#include "zensim/simulation/transfer/G2P.hpp"
#include "zensim/simulation/transfer/G2P2G.hpp"
#include "zensim/simulation/transfer/P2G.hpp"

#include "zensim/container/Bvh.hpp"

#include "zpcb200/zs_overlay.cuh"  // the binding of INTEGRATION.md, compiled against the headers above

using namespace zs;

namespace {
  struct RefMpmCuda {
    int n;
    float dx;
    Particles<f32, 3> pars;
    HashTable<i32, 3, int> table;
    Grids<f32, 3, 4> grids;
    Vector<float> maxVel;
    int nblocks{0};
    RefMpmCuda(int n_, float dx_, int expectedBlocks)
        : n{n_},
          dx{dx_},
          pars{(size_t)n_, memsrc_e::device, 0},
          table{(size_t)expectedBlocks, memsrc_e::device, 0},
          grids{{{"m", 1}, {"v", 3}, {"rhs", 3}}, dx_, (size_t)expectedBlocks, memsrc_e::device, 0},
          maxVel{1, memsrc_e::device, 0} {
      pars.addAttr("m", attrib_e::scalar);
      pars.addAttr("v", attrib_e::vector);
      pars.addAttr("F", attrib_e::matrix);
      pars.addAttr("C", attrib_e::matrix);
    }
  };
  void h2d(void *dst, const void *src, size_t bytes) { cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); }
  void d2h(void *dst, const void *src, size_t bytes) { cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); }
}  // namespace

/// ---- G2P2GTransfer (simulation/transfer/G2P2G.hpp:49-141): the reference's own functor, instantiated with a minimal grid-dof view ----
/// The functor is generic in GridDofView and only asks it for scalar_value_type, get(node, vector_c) and ref(i); the reference's
/// own DofView (types/View.h) does not compile under gcc 13, so this driver hands it three floats per node in a plain device
/// array — the arithmetic that runs is the reference's, line for line (round 1 could only check a restatement).
namespace {
  struct PlainDofView {
    using scalar_value_type = float;
    float *p;
    template <class Tag> __device__ vec<float, 3> get(size_t node, Tag) const { return vec<float, 3>{p[3 * node], p[3 * node + 1], p[3 * node + 2]}; }
    __device__ float &ref(size_t i) { return p[i]; }
  };
  template <class Model> void run_g2p2g(RefMpmCuda &s, float dt, const Model &model, float *gridv_dev, float *gridr_dev) {
    auto pol = cuda_exec().device(0);
    pol(range(s.n), G2P2GTransfer{exec_cuda, wrapv<transfer_scheme_e::apic>{}, dt, model, proxy<execspace_e::cuda>(s.grids.grid(collocated_c)),
                                  PlainDofView{gridv_dev}, PlainDofView{gridr_dev}, s.table, s.pars});
  }
}  // namespace

extern "C" {
void *zpcrefcuda_mpm_create(int n, float dx, int expectedBlocks) { return new RefMpmCuda{n, dx, expectedBlocks}; }
void zpcrefcuda_mpm_destroy(void *h) { delete (RefMpmCuda *)h; }
void zpcrefcuda_mpm_set_particles(void *h, const float *x, const float *v, const float *m, const float *C, const float *F) {
  auto &s = *(RefMpmCuda *)h;
  h2d(s.pars.attrVector("x").data(), x, sizeof(float) * 3 * s.n);
  h2d(s.pars.attrVector("v").data(), v, sizeof(float) * 3 * s.n);
  h2d(s.pars.attrScalar("m").data(), m, sizeof(float) * s.n);
  h2d(s.pars.attrMatrix("C").data(), C, sizeof(float) * 9 * s.n);
  h2d(s.pars.attrMatrix("F").data(), F, sizeof(float) * 9 * s.n);
}
void zpcrefcuda_mpm_get_particles(void *h, float *x, float *v, float *C, float *F) {
  auto &s = *(RefMpmCuda *)h;
  d2h(x, s.pars.attrVector("x").data(), sizeof(float) * 3 * s.n);
  d2h(v, s.pars.attrVector("v").data(), sizeof(float) * 3 * s.n);
  d2h(C, s.pars.attrMatrix("C").data(), sizeof(float) * 9 * s.n);
  d2h(F, s.pars.attrMatrix("F").data(), sizeof(float) * 9 * s.n);
}
/// CleanSparsity, ComputeSparsity, EnlargeSparsity{0,2}; returns table.size() (a device-to-host read, like the reference's app)
int zpcrefcuda_mpm_partition(void *h) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = cuda_exec().device(0);
  constexpr auto tag = exec_cuda;
  pol(range(s.table._tableSize), CleanSparsity{tag, s.table});
  pol(range(s.n), ComputeSparsity{tag, s.dx, 4, s.table, s.pars.attrVector("x")});
  const int cnt = s.table.size();
  pol(range(cnt), EnlargeSparsity{tag, s.table, vec<int, 3>{0, 0, 0}, vec<int, 3>{2, 2, 2}});
  s.nblocks = s.table.size();
  return s.nblocks;
}
void zpcrefcuda_mpm_get_keys(void *h, int *keys) {
  auto &s = *(RefMpmCuda *)h;
  d2h(keys, s.table._activeKeys.data(), sizeof(int) * 3 * s.nblocks);
}
void zpcrefcuda_mpm_clean_grid(void *h) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = cuda_exec().device(0);
  pol(Collapse{(size_t)s.nblocks, (size_t)64}, CleanGridBlocks{exec_cuda, s.grids});
}
void zpcrefcuda_mpm_p2g(void *h, float dt, float E, float nu, float volume) {
  auto &s = *(RefMpmCuda *)h;
  FixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.volume = volume;
  auto pol = cuda_exec().device(0);
  pol(range(s.n), P2GTransfer{exec_cuda, wrapv<transfer_scheme_e::apic>{}, dt, model, s.pars, s.table, s.grids});
}
/// mode 0: ComputeGridBlockVelocity as shipped; mode 1: "mv += rhs" first (explicit update).  Returns max |v|^2.
float zpcrefcuda_mpm_grid_update(void *h, float dt, float gravity, int mode) {
  auto &s = *(RefMpmCuda *)h;
  s.maxVel.setVal(0.f);
  auto pol = cuda_exec().device(0);
  if (mode == 1) {
    auto gv = proxy<execspace_e::cuda>(s.grids);
    pol(Collapse{(size_t)s.nblocks, (size_t)64}, [gv] __device__(int b, int c) mutable {
      auto block = gv[b];
      for (int d = 0; d != 3; ++d) block(1 + d, c) += block(4 + d, c);
    });
  }
  pol(Collapse{(size_t)s.nblocks, (size_t)64},
      ComputeGridBlockVelocity{exec_cuda, wrapv<transfer_scheme_e::apic>{}, s.grids, dt, gravity, s.maxVel.data()});
  return s.maxVel.getVal();
}
/// the reference's GridAngularMomentum / GridMomentumToVelocity (GridOp.hpp:184-262) on cuda_exec(): out6 first, then the velocities
float zpcrefcuda_mpm_grid_momentum(void *h, double *out6) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = cuda_exec().device(0);
  Vector<double> sum{6, memsrc_e::device, 0};
  cudaMemset(sum.data(), 0, sizeof(double) * 6);
  pol(Collapse{(size_t)s.nblocks, (size_t)64}, GridAngularMomentum{exec_cuda, s.table, s.grids.grid(collocated_c), 0, 1, sum.data()});
  d2h(out6, sum.data(), sizeof(double) * 6);
  s.maxVel.setVal(0.f);
  pol(Collapse{(size_t)s.nblocks, (size_t)64}, GridMomentumToVelocity{exec_cuda, s.grids.grid(collocated_c), 0, 1, s.maxVel.data()});
  return s.maxVel.getVal();
}
/// the same two through the overlay (libzpcb200's kernels on the reference's containers)
float zpcrefcuda_overlay_grid_momentum(void *h, double *out6) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = b200_exec();
  Vector<double> sum{6, memsrc_e::device, 0};
  cudaMemset(sum.data(), 0, sizeof(double) * 6);
  b200::grid_angular_momentum(pol, s.table, s.grids, sum.data());
  d2h(out6, sum.data(), sizeof(double) * 6);
  s.maxVel.setVal(0.f);
  b200::grid_momentum_to_velocity(pol, s.grids, s.table, s.maxVel.data());
  return s.maxVel.getVal();
}
void zpcrefcuda_mpm_g2p(void *h, float dt) {
  auto &s = *(RefMpmCuda *)h;
  FixedCorotatedConfig model{};
  auto pol = cuda_exec().device(0);
  pol(range(s.n), G2PTransfer{exec_cuda, wrapv<transfer_scheme_e::apic>{}, dt, model, s.grids, s.table, s.pars});
}
/// model_kind 0 fixed-corotated {E, nu}, 1 von Mises {E, nu, yieldStress}; gridv / gridr: host arrays [nblocks * 64 * 3] in the
/// partition's block numbering (zpcrefcuda_mpm_get_keys); gridr is zeroed first and returned.
int zpcrefcuda_mpm_g2p2g(void *h, int model_kind, const float *prm, float volume, float dt, const float *gridv, float *gridr) {
  auto &s = *(RefMpmCuda *)h;
  const size_t ndof = (size_t)s.nblocks * 64 * 3;
  Vector<float> gv{ndof, memsrc_e::device, 0}, gr{ndof, memsrc_e::device, 0};
  h2d(gv.data(), gridv, sizeof(float) * ndof);
  cudaMemset(gr.data(), 0, sizeof(float) * ndof);
  if (model_kind == 0) {
    FixedCorotatedConfig m{};
    m.E = prm[0]; m.nu = prm[1]; m.volume = volume;
    run_g2p2g(s, dt, m, gv.data(), gr.data());
  } else if (model_kind == 1) {
    VonMisesFixedCorotatedConfig m{};
    m.E = prm[0]; m.nu = prm[1]; m.yieldStress = prm[2]; m.volume = volume;
    run_g2p2g(s, dt, m, gv.data(), gr.data());
  } else
    return -1;
  cudaDeviceSynchronize();
  d2h(gridr, gr.data(), sizeof(float) * ndof);
  return 0;
}
void zpcrefcuda_mpm_get_grid(void *h, float *out) {
  auto &s = *(RefMpmCuda *)h;
  d2h(out, s.grids.grid(collocated_c).blocks.data(), sizeof(float) * 7 * 64 * s.nblocks);
}
void zpcrefcuda_sync() { cudaDeviceSynchronize(); }

/// ---- primitives through the reference's CudaExecutionPolicy (CUB underneath, ExecutionPolicy.cuh:552-866), device pointers ----
void zpcrefcuda_radix_sort_pair_u32(const unsigned *kin, const int *vin, unsigned *kout, int *vout, size_t n) {
  auto pol = cuda_exec().device(0);
  unsigned *ki = const_cast<unsigned *>(kin);
  int *vi = const_cast<int *>(vin);
  radix_sort_pair(pol, ki, vi, kout, vout, (std::ptrdiff_t)n, 0, 32);   // all lvalues: KeyIter is deduced from both key arguments
}
void zpcrefcuda_exclusive_scan_i32(const int *in, int *out, size_t n) {
  auto pol = cuda_exec().device(0);
  int *first = const_cast<int *>(in), *last = first + n;
  exclusive_scan(pol, first, last, out);
}
void zpcrefcuda_reduce_sum_i32(const int *in, int *out, size_t n) {
  auto pol = cuda_exec().device(0);
  int *first = const_cast<int *>(in), *last = first + n;
  reduce(pol, first, last, out, 0);
}

/// ---- the same composed substep and primitives through the overlay (include/zpcb200/zs_overlay.cuh): the reference's containers,
/// this repository's kernels.  What a zpc application runs after replacing cuda_exec() by b200_exec() and the five functor launches.
int zpcrefcuda_overlay_partition(void *h) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = b200_exec();
  b200::partition_for_particles(pol, s.table, s.pars, s.dx);
  s.nblocks = s.table.size();
  return s.nblocks;
}
void zpcrefcuda_overlay_clean_grid(void *h) {
  auto &s = *(RefMpmCuda *)h;
  b200::clean_grid_blocks(b200_exec(), s.table, s.grids);
}
void zpcrefcuda_overlay_p2g(void *h, float dt, float E, float nu, float volume) {
  auto &s = *(RefMpmCuda *)h;
  FixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.volume = volume;
  b200::p2g(b200_exec(), dt, model, s.pars, s.table, s.grids);
}
float zpcrefcuda_overlay_grid_update(void *h, float dt, float gravity, int mode) {
  auto &s = *(RefMpmCuda *)h;
  s.maxVel.setVal(0.f);
  b200::compute_grid_block_velocity(b200_exec(), s.grids, s.table, dt, gravity, s.maxVel.data(), mode);
  return s.maxVel.getVal();
}
void zpcrefcuda_overlay_g2p(void *h, float dt) {
  auto &s = *(RefMpmCuda *)h;
  b200::g2p(b200_exec(), dt, s.grids, s.table, s.pars);
}
/// compile-time coverage of the remaining overlay launches (other constitutive models, analytic colliders); run by no test yet
void zpcrefcuda_overlay_models_and_colliders(void *h, float dt) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = b200_exec();
  using TV = vec<float, 3>;
  b200::p2g(pol, dt, VonMisesFixedCorotatedConfig{}, s.pars, s.table, s.grids);
  b200::p2g(pol, dt, EquationOfStateConfig{}, s.pars, s.table, s.grids);
  b200::p2g(pol, dt, DruckerPragerConfig{}, s.pars, s.table, s.grids);
  b200::p2g(pol, dt, NACCConfig{}, s.pars, s.table, s.grids);
  b200::g2p(pol, dt, EquationOfStateConfig{}, s.grids, s.table, s.pars);
  Collider plane{AnalyticLevelSet<analytic_geometry_e::Plane, float, 3>{TV{0.f, 0.1f, 0.f}, TV{0.f, 1.f, 0.f}}, collider_e::Separate};
  Collider sphere{AnalyticLevelSet<analytic_geometry_e::Sphere, float, 3>{TV{0.3f, 0.3f, 0.3f}, 0.1f}, collider_e::Slip};
  Collider box{AnalyticLevelSet<analytic_geometry_e::Cuboid, float, 3>{TV{0.2f, 0.2f, 0.2f}, TV{0.3f, 0.3f, 0.4f}}, collider_e::Sticky};
  b200::apply_boundary_condition(pol, plane, s.table, s.grids);
  b200::apply_boundary_condition(pol, sphere, s.table, s.grids);
  b200::apply_boundary_condition(pol, box, s.table, s.grids);
}
/// generic code templated on the policy, unchanged: zs::radix_sort_pair / exclusive_scan / reduce with b200_exec() — host arrays
/// in, host arrays out (zs::Vector on the device in between, so the Vector-iterator path of the overlay is the one exercised)
void zpcrefcuda_overlay_prims(const unsigned *keys, const int *vals, unsigned *keysOut, int *valsOut, int *scanOut, int *sumOut, int *maxOut,
                              size_t n) {
  auto pol = b200_exec();
  Vector<unsigned> k{n, memsrc_e::device, 0}, ko{n, memsrc_e::device, 0};
  Vector<int> v{n, memsrc_e::device, 0}, vo{n, memsrc_e::device, 0}, sc{n, memsrc_e::device, 0}, red{2, memsrc_e::device, 0};
  h2d(k.data(), keys, sizeof(unsigned) * n);
  h2d(v.data(), vals, sizeof(int) * n);
  radix_sort_pair(pol, k.begin(), v.begin(), ko.begin(), vo.begin(), (std::ptrdiff_t)n);
  exclusive_scan(pol, v.begin(), v.end(), sc.begin());
  reduce(pol, v.begin(), v.end(), red.begin(), 0);
  reduce(pol, v.begin(), v.end(), red.begin() + 1, detail::deduce_numeric_lowest<int>(), getmax<int>{});
  d2h(keysOut, ko.data(), sizeof(unsigned) * n);
  d2h(valsOut, vo.data(), sizeof(int) * n);
  d2h(scanOut, sc.data(), sizeof(int) * n);
  d2h(sumOut, red.data(), sizeof(int));
  d2h(maxOut, red.data() + 1, sizeof(int));
}

/// The reference's own LBvh<3,int,f32>::build (container/Bvh.hpp:835-1000), unchanged, on the device — with cuda_exec() (use_b200 = 0)
/// or with b200_exec() (use_b200 = 1): the template is generic in the policy, so its radix_sort_pair and exclusive_scan then run in
/// libzpcb200 while its own functors keep launching through the inherited operator().  Host arrays in / out; returns numNodes.
int zpcrefcuda_lbvh_build(int use_b200, int n, const float *bvs, float *orderedBvs, int *auxIndices, int *parents, int *levels, int *leafInds) {
  using Bvh = LBvh<3, int, float>;
  using Box = typename Bvh::Box;
  Vector<Box> prims{(size_t)n, memsrc_e::device, 0};
  h2d((void *)prims.data(), bvs, sizeof(float) * 6 * n);
  Bvh bvh{};
  if (use_b200) {
    auto pol = b200_exec();
    bvh.build(pol, prims, true_c);
  } else {
    auto pol = cuda_exec().device(0);
    bvh.build(pol, prims, true_c);
  }
  const int nn = (int)bvh.getNumNodes();
  d2h(orderedBvs, (const void *)bvh.orderedBvs.data(), sizeof(float) * 6 * nn);
  d2h(auxIndices, bvh.auxIndices.data(), sizeof(int) * bvh.auxIndices.size());
  if (n > 2) {
    d2h(parents, bvh.parents.data(), sizeof(int) * nn);
    d2h(levels, bvh.levels.data(), sizeof(int) * nn);
  }
  d2h(leafInds, bvh.leafInds.data(), sizeof(int) * n);
  return nn;
}

/// TileVector channels through the overlay, passed the way the reference's own C layer passes them (LegacyIterator<aosoa_iterator<T, 1>>,
/// py_interop/GenericIterator.hpp): data = an AoSoA buffer of `ntiles` tiles x nch channels x 32 lanes (host, int); channel chn of the first
/// n elements is scanned into channel chn of `out` (same layout) and reduced (sum) into *sum.
void zpcrefcuda_overlay_aosoa(const int *data, int ntiles, int nch, int chn, size_t n, int *out, int *sum) {
  const size_t total = (size_t)ntiles * nch * 32;
  Vector<int> buf{total, memsrc_e::device, 0}, obuf{total, memsrc_e::device, 0}, red{1, memsrc_e::device, 0};
  h2d(buf.data(), data, sizeof(int) * total);
  cudaMemset(obuf.data(), 0, sizeof(int) * total);
  using It = LegacyIterator<aosoa_iterator<int, 1>>;
  static_assert(b200_detail::raw_iter<It>::is_aosoa && b200_detail::raw_iter<It>::ok, "the overlay takes the reference's aosoa iterators to the library");
  static_assert(b200_detail::raw_iter<decltype(zs::begin(declval<Vector<int> &>()))>::is_vector_iter && !b200_detail::raw_iter<int *>::is_aosoa, "iterator classification");
  It first{aosoa_iterator<int, 1>{wrapv<layout_e::aosoa>{}, buf.data(), 0u, 32u, (u32)chn, (u32)nch}};
  It last{aosoa_iterator<int, 1>{wrapv<layout_e::aosoa>{}, buf.data(), (u32)n, 32u, (u32)chn, (u32)nch}};
  It ofirst{aosoa_iterator<int, 1>{wrapv<layout_e::aosoa>{}, obuf.data(), 0u, 32u, (u32)chn, (u32)nch}};
  It rfirst{aosoa_iterator<int, 1>{wrapv<layout_e::aos>{}, red.data(), 0u}};
  auto pol = b200_exec();
  exclusive_scan(pol, first, last, ofirst);
  reduce(pol, first, last, rfirst, 0);
  d2h(out, obuf.data(), sizeof(int) * total);
  d2h(sum, red.data(), sizeof(int));
}

/// BASELINE config C5 "GB/s vs reference CudaExecutionPolicy": the reference's radix_sort_pair (u32 keys, i32 values, 32 bits) /
/// exclusive_scan (i32) / reduce (i32 sum) through cuda_exec() with its defaults (sync(true): every call ends with a stream
/// synchronisation, ExecutionPolicy.cuh:824), on device-resident zs::Vectors of n elements; wall clock per call, mean of `iters`
/// after one warm-up each.  use_b200 = 1 times the same generic calls with b200_exec() instead.  ms[0..2] = sort, scan, reduce.
void zpcrefcuda_prims_bench(int use_b200, size_t n, int iters, double *ms) {
  Vector<unsigned> k{n, memsrc_e::device, 0}, ko{n, memsrc_e::device, 0};
  Vector<int> v{n, memsrc_e::device, 0}, vo{n, memsrc_e::device, 0}, red{1, memsrc_e::device, 0};
  {
    std::vector<unsigned> hk(n);
    std::vector<int> hv(n);
    unsigned x = 12345u;
    for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; hk[i] = x; hv[i] = (int)(i & 1); }
    h2d(k.data(), hk.data(), sizeof(unsigned) * n);
    h2d(v.data(), hv.data(), sizeof(int) * n);
  }
  auto run = [&](auto &pol) {
    auto timeit = [&](auto &&f) {
      f();
      cudaDeviceSynchronize();
      const auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < iters; ++i) f();
      cudaDeviceSynchronize();
      return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    };
    ms[0] = timeit([&] { radix_sort_pair(pol, k.begin(), v.begin(), ko.begin(), vo.begin(), (std::ptrdiff_t)n); });
    ms[1] = timeit([&] { exclusive_scan(pol, v.begin(), v.end(), vo.begin()); });
    ms[2] = timeit([&] { reduce(pol, v.begin(), v.end(), red.begin(), 0); });
  };
  if (use_b200) {
    auto pol = b200_exec();
    run(pol);
  } else {
    auto pol = cuda_exec().device(0);
    run(pol);
  }
}

/// compile-time coverage of the SparseGrid seam of the overlay: the reference's SparseGrid<3, f32, 8> on the device, one substep through
/// zs::b200::sg_*; x, v, C, F come back to the host.  (Run by no test yet.)
void zpcrefcuda_overlay_sparsegrid_substep(void *h, int nblocks, float dt, float E, float nu, float volume, float gravity) {
  auto &s = *(RefMpmCuda *)h;
  SparseGrid<3, f32, 8> sg{7, (size_t)nblocks, memsrc_e::device, 0};
  sg.scale(s.dx);
  auto pol = b200_exec();
  FixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.volume = volume;
  b200::sg_partition_for_particles(pol, sg, s.pars);
  b200::sg_clean(pol, sg);
  b200::sg_p2g(pol, dt, model, s.pars, sg);
  s.maxVel.setVal(0.f);
  b200::sg_grid_update(pol, sg, dt, gravity, s.maxVel.data(), 1);
  b200::sg_g2p(pol, dt, sg, s.pars);
}

/// The fast path of the overlay on the reference's containers: bin the particles (TileVector<f32, 32>), `steps` substeps with a
/// re-bin every `rebinEvery`, unbin back into the reference's Particles (slot order = bin order; order[] maps slots to the
/// original particles).  The partition is rebuilt every substep with the reference's EnlargeSparsity{0, 2}.
void zpcrefcuda_overlay_binned_substeps(void *h, int steps, int rebinEvery, float dt, float E, float nu, float volume, float gravity, int *order) {
  auto &s = *(RefMpmCuda *)h;
  auto pol = b200_exec();
  FixedCorotatedConfig model{};
  model.E = E;
  model.nu = nu;
  model.volume = volume;
  b200::partition_for_particles(pol, s.table, s.pars, s.dx);
  b200::BinnedParticles bins{(size_t)s.n, (int)(s.table._tableSize / 16)};
  b200::bin_particles(pol, s.pars, s.table, s.dx, bins);
  std::vector<int> perm(s.n), tmp(s.n);
  d2h(perm.data(), bins.order.data(), sizeof(int) * s.n);
  for (int i = 0; i < steps; ++i) {
    b200::partition_for_particles(pol, s.table, bins, s.dx);
    if (i > 0 && rebinEvery > 0 && i % rebinEvery == 0) {
      // the re-bin permutes the slots again: compose the maps through the sort's own output? the C entry returns none, so the
      // identity of a particle is carried by the caller (the test tags particles by their mass)
      b200::rebin_particles(pol, s.table, s.dx, bins);
    }
    b200::clean_grid_blocks(pol, s.table, s.grids);
    b200::p2g(pol, dt, model, bins, s.table, s.grids);
    s.maxVel.setVal(0.f);
    b200::compute_grid_block_velocity(pol, s.grids, s.table, dt, gravity, s.maxVel.data(), 1);
    b200::g2p(pol, dt, s.grids, s.table, bins);
  }
  b200::unbin_particles(pol, bins, s.pars);
  // mass travels with the particle: hand it back too (unbin writes M when the view has it)
  std::memcpy(order, perm.data(), sizeof(int) * s.n);
}
void zpcrefcuda_mpm_get_mass(void *h, float *m) {
  auto &s = *(RefMpmCuda *)h;
  d2h(m, s.pars.attrScalar("m").data(), sizeof(float) * s.n);
}
}

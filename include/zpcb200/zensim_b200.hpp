// C++ host-side mirror of the reference interface for the MPM transfer path, over the C ABI (zpcb200.h).
//
// Header-only; needs only <cuda_runtime.h> and libzpcb200.so.  Names, argument meaning and error behaviour follow
// zenustech/zpc (paths relative to include/zensim/):
//   CudaExecutionPolicy + free functions reduce / exclusive_scan / inclusive_scan / radix_sort / radix_sort_pair
//       <- cuda/execution/ExecutionPolicy.cuh:362-912, execution/ExecutionPolicy.hpp:684-781
//          (chained setters device().stream().sync(); sync defaults to true; errors are latched, never thrown)
//   plus<T>, multiplies<T>, getmax<T>, getmin<T>  <- ZpcFunctional.hpp:60-117
//   Vector<T>                               <- container/Vector.hpp            (device memory, getVal/setVal)
//   HashTable (i32,3,int)                   <- container/HashTable.hpp:15-206  (tableSize = next_2pow(n) * 16)
//   Grids (f32,3,4) {m, v, rhs}             <- geometry/Structure.hpp:140-260
//   Particles (f32,3) AoS x,v,m,C,F         <- geometry/Structurefree.hpp:22-224
//   Bht (i32,3,int,16), SparseGrid (3,f32,8) <- container/Bht.hpp, geometry/SparseGrid.hpp  (+ Sg* functors on side-8 blocks)
//   merge_sort, merge_sort_pair             <- cuda/execution/ExecutionPolicy.cuh:686-760
//   partition_for_particles, CleanGridBlocks, P2GTransfer, ComputeGridBlockVelocity, G2PTransfer
//       <- simulation/{sparsity,grid,transfer}/*.hpp; invoked as pol(functor) instead of pol(range, functor).
// There is no host fallback: every call lands in the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../zpcb200.h"

namespace zsb200 {

template <typename T = void> struct plus {};
template <typename T = void> struct getmax {};
template <typename T = void> struct getmin {};
template <typename T = void> struct multiplies {};

template <typename T> struct Vector {  // device vector
  T *_ptr{nullptr};
  size_t _n{0};
  Vector() = default;
  explicit Vector(size_t n) : _n{n} {
    if (cudaMalloc((void **)&_ptr, sizeof(T) * (n ? n : 1)) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
  }
  explicit Vector(const std::vector<T> &h) : Vector(h.size()) { cudaMemcpy(_ptr, h.data(), sizeof(T) * _n, cudaMemcpyHostToDevice); }
  Vector(const Vector &) = delete;
  Vector &operator=(const Vector &) = delete;
  Vector(Vector &&o) noexcept : _ptr{o._ptr}, _n{o._n} { o._ptr = nullptr; o._n = 0; }
  Vector &operator=(Vector &&o) noexcept {
    if (this != &o) { if (_ptr) cudaFree(_ptr); _ptr = o._ptr; _n = o._n; o._ptr = nullptr; o._n = 0; }
    return *this;
  }
  ~Vector() { if (_ptr) cudaFree(_ptr); }
  T *data() { return _ptr; }
  const T *data() const { return _ptr; }
  T *begin() { return _ptr; }
  T *end() { return _ptr + _n; }
  size_t size() const { return _n; }
  T getVal(size_t i = 0) const { T v; cudaMemcpy(&v, _ptr + i, sizeof(T), cudaMemcpyDeviceToHost); return v; }  // Vector.hpp getVal
  void setVal(T v, size_t i = 0) { cudaMemcpy(_ptr + i, &v, sizeof(T), cudaMemcpyHostToDevice); }
  std::vector<T> toHost() const { std::vector<T> h(_n); cudaMemcpy(h.data(), _ptr, sizeof(T) * _n, cudaMemcpyDeviceToHost); return h; }
};

template <typename T> inline zpc_port make_port(const T *p) { return zpc_port{(void *)p, 0, 0, 0, 1}; }

struct CudaExecutionPolicy {
  int _device{0};
  cudaStream_t _stream{nullptr};
  bool _sync{true};
  mutable int _lastError{0};
  mutable void *_scratch{nullptr};
  mutable size_t _scratchBytes{0};
  CudaExecutionPolicy() = default;
  CudaExecutionPolicy(const CudaExecutionPolicy &o) : _device{o._device}, _stream{o._stream}, _sync{o._sync} {}
  ~CudaExecutionPolicy() { if (_scratch) cudaFree(_scratch); }
  CudaExecutionPolicy &device(int d) { _device = d; return *this; }
  CudaExecutionPolicy &stream(cudaStream_t s) { _stream = s; return *this; }
  CudaExecutionPolicy &sync(bool b) { _sync = b; return *this; }
  bool shouldSync() const { return _sync; }
  void *getStream() const { return (void *)_stream; }
  int lastError() const { return _lastError; }

  void *scratch(size_t bytes) const {
    if (bytes > _scratchBytes) {
      if (_scratch) cudaFree(_scratch);
      _scratchBytes = bytes + bytes / 4 + 256;
      if (cudaMalloc(&_scratch, _scratchBytes) != cudaSuccess) { _scratch = nullptr; _scratchBytes = 0; }
    }
    return _scratch;
  }
  template <typename Fn> void twoPhase(Fn fn) const {  // CUB-style size query then run (ExecutionPolicy.cuh:803-812)
    cudaSetDevice(_device);
    size_t bytes = 0;
    int rc = fn(nullptr, &bytes);
    if (!rc) {
      void *t = scratch(bytes ? bytes : 1);
      size_t cap = _scratchBytes;
      rc = t ? fn(t, &cap) : (int)cudaErrorMemoryAllocation;
    }
    finish(rc);
  }
  void finish(int rc) const {
    if (rc) _lastError = rc;  // latched like CudaContext::errorStatus (cuda/Cuda.h:291-312)
    if (_sync) { cudaError_t e = cudaStreamSynchronize(_stream); if (e != cudaSuccess) _lastError = (int)e; }
  }

  // ---- primitives -----------------------------------------------------------------------------------------
#define ZSB_DISPATCH_T(CALL_I32, CALL_U32, CALL_I64, CALL_F32)                                              \
  if constexpr (std::is_same_v<T, int32_t>) { twoPhase([&](void *t, size_t *b) { return CALL_I32; }); }      \
  else if constexpr (std::is_same_v<T, uint32_t>) { twoPhase([&](void *t, size_t *b) { return CALL_U32; }); } \
  else if constexpr (std::is_same_v<T, int64_t>) { twoPhase([&](void *t, size_t *b) { return CALL_I64; }); }  \
  else if constexpr (std::is_same_v<T, float>) { twoPhase([&](void *t, size_t *b) { return CALL_F32; }); }    \
  else static_assert(sizeof(T) == 0, "unsupported element type");
  template <typename T, template <class> class Op, typename U>
  void reduce(const T *first, const T *last, T *d_first, T /*init = identity*/, Op<U>) const {
    const size_t n = (size_t)(last - first);
    const zpc_port in = make_port(first), out = make_port(d_first);
    if constexpr (std::is_same_v<Op<U>, plus<U>>) {
      ZSB_DISPATCH_T(zpcb200_reduce_sum_i32(t, b, in, out, n, _stream), zpcb200_reduce_sum_u32(t, b, in, out, n, _stream),
                     zpcb200_reduce_sum_i64(t, b, in, out, n, _stream), zpcb200_reduce_sum_f32(t, b, in, out, n, _stream))
    } else if constexpr (std::is_same_v<Op<U>, multiplies<U>>) {
      ZSB_DISPATCH_T(zpcb200_reduce_prod_i32(t, b, in, out, n, _stream), zpcb200_reduce_prod_u32(t, b, in, out, n, _stream),
                     zpcb200_reduce_prod_i64(t, b, in, out, n, _stream), zpcb200_reduce_prod_f32(t, b, in, out, n, _stream))
    } else if constexpr (std::is_same_v<Op<U>, getmax<U>>) {
      ZSB_DISPATCH_T(zpcb200_reduce_max_i32(t, b, in, out, n, _stream), zpcb200_reduce_max_u32(t, b, in, out, n, _stream),
                     zpcb200_reduce_max_i64(t, b, in, out, n, _stream), zpcb200_reduce_max_f32(t, b, in, out, n, _stream))
    } else {
      ZSB_DISPATCH_T(zpcb200_reduce_min_i32(t, b, in, out, n, _stream), zpcb200_reduce_min_u32(t, b, in, out, n, _stream),
                     zpcb200_reduce_min_i64(t, b, in, out, n, _stream), zpcb200_reduce_min_f32(t, b, in, out, n, _stream))
    }
  }
  template <typename T> void exclusive_scan(const T *first, const T *last, T *d_first) const {
    const size_t n = (size_t)(last - first);
    const zpc_port in = make_port(first), out = make_port(d_first);
    ZSB_DISPATCH_T(zpcb200_exclusive_scan_sum_i32(t, b, in, out, n, _stream), zpcb200_exclusive_scan_sum_u32(t, b, in, out, n, _stream),
                   zpcb200_exclusive_scan_sum_i64(t, b, in, out, n, _stream), zpcb200_exclusive_scan_sum_f32(t, b, in, out, n, _stream))
  }
  template <typename T> void inclusive_scan(const T *first, const T *last, T *d_first) const {
    const size_t n = (size_t)(last - first);
    const zpc_port in = make_port(first), out = make_port(d_first);
    ZSB_DISPATCH_T(zpcb200_inclusive_scan_sum_i32(t, b, in, out, n, _stream), zpcb200_inclusive_scan_sum_u32(t, b, in, out, n, _stream),
                   zpcb200_inclusive_scan_sum_i64(t, b, in, out, n, _stream), zpcb200_inclusive_scan_sum_f32(t, b, in, out, n, _stream))
  }
#undef ZSB_DISPATCH_T
  template <typename K>
  void radix_sort_pair(const K *keysIn, const int *valsIn, K *keysOut, int *valsOut, size_t count, int sbit = 0,
                       int ebit = sizeof(K) * 8) const {
    const zpc_port ki = make_port(keysIn), vi = make_port(valsIn), ko = make_port(keysOut), vo = make_port(valsOut);
    if constexpr (std::is_same_v<K, uint32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_pair_u32(t, b, ki, vi, ko, vo, count, sbit, ebit, _stream); });
    else if constexpr (std::is_same_v<K, int32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_pair_i32(t, b, ki, vi, ko, vo, count, sbit, ebit, _stream); });
    else if constexpr (std::is_same_v<K, uint64_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_pair_u64(t, b, ki, vi, ko, vo, count, sbit, ebit, _stream); });
    else static_assert(sizeof(K) == 0, "radix sort keys: u32, i32, u64");
  }
  template <typename K> void radix_sort(const K *first, const K *last, K *d_first, int sbit = 0, int ebit = sizeof(K) * 8) const {
    const size_t n = (size_t)(last - first);
    const zpc_port ki = make_port(first), ko = make_port(d_first);
    if constexpr (std::is_same_v<K, uint32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_u32(t, b, ki, ko, n, sbit, ebit, _stream); });
    else if constexpr (std::is_same_v<K, int32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_i32(t, b, ki, ko, n, sbit, ebit, _stream); });
    else if constexpr (std::is_same_v<K, uint64_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_radix_sort_u64(t, b, ki, ko, n, sbit, ebit, _stream); });
    else static_assert(sizeof(K) == 0, "radix sort keys: u32, i32, u64");
  }
  // merge_sort_pair / merge_sort: stable, ascending, in place (ExecutionPolicy.cuh:686-760); keys int | float | double
  template <typename K> void merge_sort_pair(K *keys, int *vals, size_t count) const {
    const zpc_port k = make_port(keys), v = make_port(vals);
    if constexpr (std::is_same_v<K, int32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_pair_i32(t, b, k, v, count, _stream); });
    else if constexpr (std::is_same_v<K, float>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_pair_f32(t, b, k, v, count, _stream); });
    else if constexpr (std::is_same_v<K, double>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_pair_f64(t, b, k, v, count, _stream); });
    else static_assert(sizeof(K) == 0, "merge sort keys: int, float, double");
  }
  template <typename K> void merge_sort(K *first, K *last) const {
    const size_t n = (size_t)(last - first);
    const zpc_port k = make_port(first);
    if constexpr (std::is_same_v<K, int32_t>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_i32(t, b, k, n, _stream); });
    else if constexpr (std::is_same_v<K, float>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_f32(t, b, k, n, _stream); });
    else if constexpr (std::is_same_v<K, double>) twoPhase([&](void *t, size_t *b) { return zpcb200_merge_sort_f64(t, b, k, n, _stream); });
    else static_assert(sizeof(K) == 0, "merge sort keys: int, float, double");
  }
  // pol(functor): one C call per functor (the reference writes pol(range, functor))
  template <typename F> void operator()(F &&f) const { finish(f.launch(*this)); }
};
inline CudaExecutionPolicy cuda_exec() { return CudaExecutionPolicy{}; }

// free-function wrappers (execution/ExecutionPolicy.hpp:684-781)
template <typename T, typename Op> void reduce(const CudaExecutionPolicy &p, const T *f, const T *l, T *o, T init, Op op) { p.reduce(f, l, o, init, op); }
template <typename T> void exclusive_scan(const CudaExecutionPolicy &p, const T *f, const T *l, T *o) { p.exclusive_scan(f, l, o); }
template <typename T> void inclusive_scan(const CudaExecutionPolicy &p, const T *f, const T *l, T *o) { p.inclusive_scan(f, l, o); }
template <typename K> void radix_sort(const CudaExecutionPolicy &p, const K *f, const K *l, K *o, int sbit = 0, int ebit = sizeof(K) * 8) { p.radix_sort(f, l, o, sbit, ebit); }
template <typename K>
void radix_sort_pair(const CudaExecutionPolicy &p, const K *ki, const int *vi, K *ko, int *vo, size_t count, int sbit = 0, int ebit = sizeof(K) * 8) {
  p.radix_sort_pair(ki, vi, ko, vo, count, sbit, ebit);
}

template <typename K> void merge_sort_pair(const CudaExecutionPolicy &p, K *keys, int *vals, size_t count) { p.merge_sort_pair(keys, vals, count); }
template <typename K> void merge_sort(const CudaExecutionPolicy &p, K *first, K *last) { p.merge_sort(first, last); }

// ---- containers of the MPM path -----------------------------------------------------------------------------
inline size_t next_2pow(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }

struct HashTable {  // HashTable<i32,3,int>
  int _tableSize;
  Vector<int> keys, indices, status, _activeKeys, _cnt, _overflow;
  explicit HashTable(size_t numExpectedEntries)
      : _tableSize{(int)(next_2pow(numExpectedEntries) * 16)}, keys((size_t)_tableSize * 3), indices(_tableSize), status(_tableSize),
        _activeKeys((size_t)_tableSize * 3), _cnt(1), _overflow(1) { _cnt.setVal(0); _overflow.setVal(0); }
  int size() const { return _cnt.getVal(); }  // the D2H read of HashTable.hpp:152
  zpc_hashtable_view view() { return zpc_hashtable_view{keys.data(), indices.data(), status.data(), _activeKeys.data(), _tableSize, _cnt.data()}; }
};
struct Grids {  // Grids<f32,3,4>, channels {"m",1},{"v",3},{"rhs",3}
  float _dx;
  size_t _numBlocks;
  Vector<float> blocks;
  Grids(float dx, size_t numBlocks) : _dx{dx}, _numBlocks{numBlocks}, blocks(numBlocks * 7 * 64) {}
  zpc_grids_view view() { return zpc_grids_view{blocks.data(), _numBlocks, 7, _dx}; }
};
struct Particles {  // Particles<f32,3>, AoS attributes
  size_t _n;
  Vector<float> X, V, M, C, F, logJp;  // logJp: addAttr("logJp", scalar) of the plastic models, allocated on request
  explicit Particles(size_t n, bool withLogJp = false) : _n{n}, X(3 * n), V(3 * n), M(n), C(9 * n), F(9 * n) {
    if (withLogJp) logJp = Vector<float>(n);
  }
  size_t size() const { return _n; }
  zpc_particles_view view() { return zpc_particles_view{M.data(), X.data(), V.data(), nullptr, nullptr, F.data(), C.data(), logJp.data(), _n}; }
};
struct FixedCorotatedConfig { float rho{1e3f}, volume{1.f}; int dim{3}; float E{5e4f}, nu{0.4f}; };  // ConstitutiveModel.hpp:739-742
struct VonMisesFixedCorotatedConfig { float rho{1e3f}, volume{1.f}; int dim{3}; float E{5e4f}, nu{0.4f}, yieldStress{240e6f}; };  // :743-747

struct DruckerPragerConfig {  // :748-757
  float rho{1e3f}, volume{1.f}; int dim{3}; float E{5e4f}, nu{0.4f}, logJp0{0.f}, fa{30.f}, cohesion{0.f}, beta{1.f};
  bool volumeCorrection{true}; float yieldSurface{0.816496580927726f * 2.f * 0.5f / (3.f - 0.5f)};
};
struct NACCConfig {  // :758-776
  float rho{1e3f}, volume{1.f}; int dim{3}; float E{5e4f}, nu{0.4f}, logJp0{-0.01f}, fa{45.f}, xi{0.8f}, beta{0.5f}; bool hardeningOn{true};
};

// bht<i32,3,int,16> (container/Bht.hpp): buckets of 16, three universal hashes from std::mt19937(2), 16-byte key slots
struct Bht {
  size_t _tableSize;
  uint32_t _hf[6];
  Vector<int> keys, indices, status, _activeKeys, _cnt, _buildSuccess, _overflow;
  explicit Bht(size_t numExpectedEntries)
      : _tableSize{zpcb200_bht_table_size(numExpectedEntries)}, keys((_tableSize ? _tableSize : 1) * 4), indices(_tableSize ? _tableSize : 1),
        status(_tableSize ? _tableSize : 1), _activeKeys((_tableSize ? _tableSize : 1) * 3), _cnt(1), _buildSuccess(1), _overflow(1) {
    zpcb200_bht_params(_hf);
    _cnt.setVal(0); _buildSuccess.setVal(1); _overflow.setVal(0);
  }
  int size() const { return _cnt.getVal(); }
  zpc_bht_view view() {
    zpc_bht_view v{keys.data(), indices.data(), status.data(), _activeKeys.data(), (uint32_t)_tableSize, (uint32_t)(_tableSize / 16),
                   _cnt.data(), _buildSuccess.data(), {}};
    for (int i = 0; i < 6; ++i) v.hf[i] = _hf[i];
    return v;
  }
};
// SparseGrid<3,f32,8> (geometry/SparseGrid.hpp:16-188): bht keyed by block origins + TileVector<f32,512> + index-to-world transform
struct SparseGrid {
  Bht _table;
  size_t _numBlocks;
  int _numChannels;
  Vector<float> _grid;
  float _transform[16];
  float _background{0.f};
  SparseGrid(int numChns, size_t numBlocks) : _table(numBlocks), _numBlocks{numBlocks}, _numChannels{numChns}, _grid((numBlocks ? numBlocks : 1) * numChns * 512) {
    for (int i = 0; i < 16; ++i) _transform[i] = (i % 5 == 0) ? 1.f : 0.f;
  }
  void scale(float s) {  // Transform::preScale (uniform)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) _transform[4 * i + j] *= s;
  }
  void translate(const float t[3]) { for (int d = 0; d < 3; ++d) _transform[12 + d] += t[d]; }  // postTranslate
  size_t numBlocks() const { return (size_t)_table.size(); }
  zpc_sparsegrid_view view() {
    zpc_sparsegrid_view v{_table.view(), _grid.data(), _numBlocks, _numChannels, {}, _background};
    for (int i = 0; i < 16; ++i) v.transform[i] = _transform[i];
    return v;
  }
};

// ---- functors ---------------------------------------------------------------------------------------------------
struct PartitionForParticles {  // CleanSparsity + ComputeSparsity + EnlargeSparsity{lo,hi}
  Particles &pars; float dx; HashTable &table; int lo{0}, hi{2};
  int launch(const CudaExecutionPolicy &pol) {
    zpc_port x{pars.X.data(), 0, 0, 0, 3};
    size_t bytes = 0;
    int rc = zpcb200_partition_build(nullptr, &bytes, x, pars.size(), dx, table.view(), lo, hi, table._overflow.data(), pol._stream);
    if (rc) return rc;
    void *t = pol.scratch(bytes);
    size_t cap = pol._scratchBytes;
    return t ? zpcb200_partition_build(t, &cap, x, pars.size(), dx, table.view(), lo, hi, table._overflow.data(), pol._stream)
             : (int)cudaErrorMemoryAllocation;
  }
};
struct SgPartitionForParticles {  // the same convention on side-8 blocks, table = the SparseGrid's bht
  Particles &pars; SparseGrid &sg; int lo{0}, hi{2};
  int launch(const CudaExecutionPolicy &pol) {
    zpc_port x{pars.X.data(), 0, 0, 0, 3};
    size_t bytes = 0;
    int rc = zpcb200_sg_partition_build(nullptr, &bytes, x, pars.size(), sg.view(), lo, hi, sg._table._overflow.data(), pol._stream);
    if (rc) return rc;
    void *t = pol.scratch(bytes);
    size_t cap = pol._scratchBytes;
    return t ? zpcb200_sg_partition_build(t, &cap, x, pars.size(), sg.view(), lo, hi, sg._table._overflow.data(), pol._stream)
             : (int)cudaErrorMemoryAllocation;
  }
};
struct SgCleanGridBlocks { SparseGrid &sg; int launch(const CudaExecutionPolicy &pol) { return zpcb200_sg_clean(sg.view(), pol._stream); } };
struct SgP2GTransfer {
  float dt; FixedCorotatedConfig model; Particles &pars; SparseGrid &sg;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
    return zpcb200_sg_p2g_apic_fcr(pars.view(), sg.view(), dt, m, pol._stream);
  }
};
struct SgComputeGridBlockVelocity {
  SparseGrid &sg; float dt; float gravity; float *maxVel; int mode{0};
  int launch(const CudaExecutionPolicy &pol) {
    const float extf[3] = {0.f, gravity, 0.f};
    return zpcb200_sg_grid_update(sg.view(), dt, extf, mode, maxVel, pol._stream);
  }
};
struct SgG2PTransfer {
  float dt; SparseGrid &sg; Particles &pars;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_sg_g2p_apic(pars.view(), sg.view(), dt, pol._stream); }
};
struct CleanGridBlocks {
  Grids &grids; HashTable &table;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_clean_grid(grids.view(), table._cnt.data(), pol._stream); }
};
struct P2GTransfer {  // P2GTransfer{cuda_c, wrapv<apic>{}, dt, model, pars, table, grids}
  float dt; FixedCorotatedConfig model; Particles &pars; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
    return zpcb200_p2g_apic_fcr(pars.view(), table.view(), grids.view(), dt, m, pol._stream);
  }
};
struct P2GTransferVonMises {  // P2GTransfer with VonMisesFixedCorotatedConfig (P2G.hpp:89-90)
  float dt; VonMisesFixedCorotatedConfig model; Particles &pars; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_vonmises_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu, model.yieldStress};
    return zpcb200_p2g_apic_vonmises(pars.view(), table.view(), grids.view(), dt, m, pol._stream);
  }
};
struct P2GTransferDruckerPrager {  // P2GTransfer with DruckerPragerConfig (P2G.hpp:92-96); pars needs logJp
  float dt; DruckerPragerConfig model; Particles &pars; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_drucker_prager m{model.rho, model.volume, model.dim, model.E, model.nu, model.logJp0, model.fa, model.cohesion, model.beta,
                         model.volumeCorrection ? 1 : 0, model.yieldSurface};
    return zpcb200_p2g_apic_drucker_prager(pars.view(), table.view(), grids.view(), dt, m, pol._stream);
  }
};
struct P2GTransferNACC {  // P2GTransfer with NACCConfig (P2G.hpp:97-100); pars needs logJp
  float dt; NACCConfig model; Particles &pars; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_nacc m{model.rho, model.volume, model.dim, model.E, model.nu, model.logJp0, model.fa, model.xi, model.beta, model.hardeningOn ? 1 : 0};
    return zpcb200_p2g_apic_nacc(pars.view(), table.view(), grids.view(), dt, m, pol._stream);
  }
};
struct ComputeGridBlockVelocity {  // {cuda_c, wrapv<apic>{}, grids, dt, gravity, maxVel}; mode 1 adds rhs (explicit update)
  Grids &grids; HashTable &table; float dt; float gravity; float *maxVel; int mode{0};
  int launch(const CudaExecutionPolicy &pol) {
    const float extf[3] = {0.f, gravity, 0.f};
    return zpcb200_grid_update(grids.view(), table._cnt.data(), dt, extf, mode, maxVel, pol._stream);
  }
};
struct GridMomentumToVelocity {  // {cuda_c, grid, mChn, mvChn, maxVel} (GridOp.hpp:184-214): v = mv / m, no gravity
  Grids &grids; HashTable &table; int mChn; int mvChn; float *maxVel;
  int launch(const CudaExecutionPolicy &pol) {
    return zpcb200_grid_momentum_to_velocity(grids.view(), table._cnt.data(), mChn, mvChn, maxVel, pol._stream);
  }
};
struct GridAngularMomentum {  // {cuda_c, table, grid, mChn, mvChn, sum} (GridOp.hpp:216-262): six doubles on the device, added to
  HashTable &table; Grids &grids; int mChn; int mvChn; double *sumAngularMomentum;
  int launch(const CudaExecutionPolicy &pol) {
    return zpcb200_grid_angular_momentum(grids.view(), table.view(), mChn, mvChn, sumAngularMomentum, pol._stream);
  }
};
// LBvh<3, int, f32> (container/Bvh.hpp:82-174): build(pol, primBvs, refit) / refit(pol, primBvs); boxes = 6 floats {min, max}
struct LBvh {
  size_t _numLeaves{0};
  Vector<float> orderedBvs;
  Vector<int> auxIndices, parents, levels, leafInds;
  size_t getNumLeaves() const { return _numLeaves; }
  size_t getNumNodes() const { return _numLeaves > 2 ? _numLeaves * 2 - 1 : _numLeaves; }
  zpc_lbvh_view view() { return zpc_lbvh_view{orderedBvs.data(), auxIndices.data(), parents.data(), levels.data(), leafInds.data()}; }
  int build(const CudaExecutionPolicy &pol, const Vector<float> &primBvs, bool refit = true) {
    _numLeaves = primBvs.size() / 6;
    const size_t nn = getNumNodes();
    orderedBvs = Vector<float>(6 * nn); auxIndices = Vector<int>(nn); parents = Vector<int>(nn); levels = Vector<int>(nn);
    leafInds = Vector<int>(_numLeaves);
    size_t bytes = 0;
    int rc = zpcb200_lbvh_build(nullptr, &bytes, primBvs.data(), _numLeaves, view(), refit, pol._stream);
    if (rc) return rc;
    Vector<char> tmp(bytes);
    rc = zpcb200_lbvh_build(tmp.data(), &bytes, primBvs.data(), _numLeaves, view(), refit, pol._stream);
    cudaStreamSynchronize(pol._stream);  // tmp is freed on return
    return rc;
  }
  int refit(const CudaExecutionPolicy &pol, const Vector<float> &primBvs) {
    if (primBvs.size() / 6 != _numLeaves) throw std::runtime_error("bvh topology changes, require rebuild!");  // Bvh.hpp:1239-1240
    size_t bytes = 0;
    int rc = zpcb200_lbvh_refit(nullptr, &bytes, primBvs.data(), _numLeaves, view(), pol._stream);
    if (rc) return rc;
    Vector<char> tmp(bytes);
    rc = zpcb200_lbvh_refit(tmp.data(), &bytes, primBvs.data(), _numLeaves, view(), pol._stream);
    cudaStreamSynchronize(pol._stream);
    return rc;
  }
};

// Collider{AnalyticLevelSet<Plane | Sphere | Cuboid>, collider_e} with its rigid motion (geometry/Collider.h:10-143,
// geometry/AnalyticLevelSet.h) and ApplyBoundaryConditionOnGridBlocks{cuda_c, collider, table, grids} (GridOp.hpp:112-164)
enum class collider_e : int { Sticky = 0, Slip = 1, Separate = 2 };
struct Collider {
  zpc_collider c;
  static Collider make(int geom, collider_e type, const float (&p0)[3], const float (&p1)[3]) {
    Collider r{zpcb200_collider_static(geom, (int)type, p0, p1)};
    return r;
  }
  static Collider plane(const float (&origin)[3], const float (&normal)[3], collider_e t = collider_e::Sticky) { return make(ZPC_GEOM_PLANE, t, origin, normal); }
  static Collider sphere(const float (&center)[3], float radius, collider_e t = collider_e::Sticky) {
    const float r[3] = {radius, 0.f, 0.f};
    return make(ZPC_GEOM_SPHERE, t, center, r);
  }
  static Collider cuboid(const float (&mn)[3], const float (&mx)[3], collider_e t = collider_e::Sticky) { return make(ZPC_GEOM_CUBOID, t, mn, mx); }
  void setTranslation(const float (&b)[3], const float (&dbdt)[3]) { for (int d = 0; d < 3; ++d) { c.b[d] = b[d]; c.dbdt[d] = dbdt[d]; } }
  void setRotation(const float (&R)[9], const float (&omega)[3]) {  // R row-major
    for (int d = 0; d < 9; ++d) c.R[d] = R[d];
    for (int d = 0; d < 3; ++d) c.omega[d] = omega[d];
  }
};
struct ApplyBoundaryConditionOnGridBlocks {
  Collider collider; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_apply_boundary(grids.view(), table.view(), collider.c, pol._stream); }
};
// ---- the block-binned fast path --------------------------------------------------------------------
// Particles::particleBins (geometry/Structurefree.hpp:220: TileVector<f32,32>, channels m x(3) v(3) C(9) F(9)) sorted by home block,
// plus the bin metadata, the cell-order cache G2P leaves for the next P2G, and the status word of include/zpcb200.h (ZPC_BINS_*).
struct BinnedParticles {
  size_t _n;
  int _binCapacity;
  Vector<float> tiles;
  Vector<int> binStart, binKey, numBins, cellOrderValid, status;
  Vector<unsigned short> cellOrder, cellStart;
  BinnedParticles(size_t n, int binCapacity)
      : _n{n}, _binCapacity{binCapacity}, tiles(((n + 31) / 32) * ZPC_PB_NCH * 32), binStart((size_t)binCapacity + 1),
        binKey((size_t)binCapacity * 3), numBins(1), cellOrderValid(1), status(1), cellOrder(n ? n : 1),
        cellStart((size_t)binCapacity * ZPCB200_CELL_GROUPS_PAD) {
    numBins.setVal(0); cellOrderValid.setVal(0); status.setVal(0);
    cudaMemset(tiles.data(), 0, sizeof(float) * tiles.size());
  }
  size_t size() const { return _n; }
  int bins() const { return numBins.getVal(); }
  int statusWord() const { return status.getVal(); }   // ZPC_BINS_* bits, 0 = fine
  zpc_bins_view view() {
    return zpc_bins_view{zpc_tilevector_view{tiles.data(), _n, ZPC_PB_NCH}, binStart.data(), binKey.data(), numBins.data(), _binCapacity,
                         cellOrder.data(), cellStart.data(), cellOrderValid.data(), status.data()};
  }
};
namespace detail {
template <class Fn> inline int two_phase(const CudaExecutionPolicy &pol, Fn fn) {
  size_t bytes = 0;
  int rc = fn(nullptr, &bytes);
  if (rc) return rc;
  void *t = pol.scratch(bytes);
  size_t cap = pol._scratchBytes;
  return t ? fn(t, &cap) : (int)cudaErrorMemoryAllocation;
}
}  // namespace detail
struct BinParticles {  // AoS Particles -> bins (needs a partition built from the same positions); orderOut[i] = AoS index of binned particle i
  Particles &pars; HashTable &table; float dx; BinnedParticles &bins; int *orderOut{nullptr};
  int launch(const CudaExecutionPolicy &pol) {
    return detail::two_phase(pol, [&](void *t, size_t *b) { return zpcb200_bin_particles(t, b, pars.view(), table.view(), dx, bins.view(), orderOut, pol._stream); });
  }
};
struct RebinParticles {  // after the particles moved: src -> dst, both binned (partition rebuilt from src's positions first)
  BinnedParticles &src; HashTable &table; float dx; BinnedParticles &dst;
  int launch(const CudaExecutionPolicy &pol) {
    return detail::two_phase(pol, [&](void *t, size_t *b) { return zpcb200_rebin_particles(t, b, src.view(), table.view(), dx, dst.view(), pol._stream); });
  }
};
struct PartitionForBinnedParticles {  // the partition from the positions of binned particles (AoSoA port on channel x)
  BinnedParticles &bins; float dx; HashTable &table; int lo{0}, hi{2};
  int launch(const CudaExecutionPolicy &pol) {
    zpc_port x{bins.tiles.data() + ZPC_PB_X * 32, 0, 5, 31, ZPC_PB_NCH};
    return detail::two_phase(pol, [&](void *t, size_t *b) {
      return zpcb200_partition_build(t, b, x, bins.size(), dx, table.view(), lo, hi, table._overflow.data(), pol._stream);
    });
  }
};
struct UnbinParticles {  // bins -> AoS Particles, in bin order
  BinnedParticles &bins; Particles &pars;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_unbin_particles(bins.view(), pars.view(), pol._stream); }
};
struct P2GTransferBinned {  // P2GTransfer<apic, FixedCorotated> on the bins: smem arena, TMA bulk reduce write-back
  float dt; FixedCorotatedConfig model; BinnedParticles &bins; HashTable &table; Grids &grids;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
    return zpcb200_p2g_apic_fcr_binned(bins.view(), table.view(), grids.view(), dt, m, pol._stream);
  }
};
struct G2PTransferBinned {  // G2PTransfer<apic> on the bins: TMA-staged arena and particle rows; leaves the cell-order cache
  float dt; Grids &grids; HashTable &table; BinnedParticles &bins;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_g2p_apic_binned(bins.view(), table.view(), grids.view(), dt, pol._stream); }
};
// the same on SparseGrid<3,f32,8>: bins = octants of the side-8 blocks; binCapacity >= 8 x the number of active blocks
struct SgBinParticles {
  Particles &pars; SparseGrid &sg; BinnedParticles &bins; int *orderOut{nullptr};
  int launch(const CudaExecutionPolicy &pol) {
    return detail::two_phase(pol, [&](void *t, size_t *b) { return zpcb200_sg_bin_particles(t, b, pars.view(), sg.view(), bins.view(), orderOut, pol._stream); });
  }
};
struct SgRebinParticles {
  BinnedParticles &src; SparseGrid &sg; BinnedParticles &dst;
  int launch(const CudaExecutionPolicy &pol) {
    return detail::two_phase(pol, [&](void *t, size_t *b) { return zpcb200_sg_rebin_particles(t, b, src.view(), sg.view(), dst.view(), nullptr, pol._stream); });
  }
};
struct SgP2GTransferBinned {
  float dt; FixedCorotatedConfig model; BinnedParticles &bins; SparseGrid &sg;
  int launch(const CudaExecutionPolicy &pol) {
    zpc_fixed_corotated m{model.rho, model.volume, model.dim, model.E, model.nu};
    return zpcb200_sg_p2g_apic_fcr_binned(bins.view(), sg.view(), dt, m, pol._stream);
  }
};
struct SgG2PTransferBinned {
  float dt; SparseGrid &sg; BinnedParticles &bins;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_sg_g2p_apic_binned(bins.view(), sg.view(), dt, pol._stream); }
};
struct G2PTransfer {
  float dt; Grids &grids; HashTable &table; Particles &pars;
  int launch(const CudaExecutionPolicy &pol) { return zpcb200_g2p_apic(pars.view(), table.view(), grids.view(), dt, pol._stream); }
};

}  // namespace zsb200

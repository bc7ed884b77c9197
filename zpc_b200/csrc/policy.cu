// Reference-style object ABI: an opaque execution policy + primitives named <op>__b200_<T>_1 that take
// iterator ports by value — the shape of zenustech/zpc's own C layer
// (include/zensim/py_interop/cuda/ExecutionPolicy.cpp:8-9, 39-134), so a binding written against that file
// can be repointed by renaming "cuda" -> "b200".  Scratch comes from the stream-ordered pool and is
// released on the same stream (reference cuda/execution/ExecutionPolicy.cuh:806-815); sync(true) is the
// default like the reference's policies (execution/ExecutionPolicy.hpp:125).
#include "common.cuh"

struct zpcb200_policy {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sync = 1;
  int last_error = 0;
};

namespace {
template <typename Fn> void run_with_scratch(zpcb200_policy *p, Fn fn) {
  if (!p) return;
  cudaSetDevice(p->device);
  size_t bytes = 0;
  int rc = fn(nullptr, &bytes);
  if (rc) { p->last_error = rc; return; }
  void *tmp = nullptr;
  cudaError_t e = cudaMallocAsync(&tmp, bytes ? bytes : 1, p->stream);
  if (e != cudaSuccess) { p->last_error = (int)e; return; }
  rc = fn(tmp, &bytes);
  cudaFreeAsync(tmp, p->stream);
  if (rc) { p->last_error = rc; return; }
  if (p->sync) {
    e = cudaStreamSynchronize(p->stream);
    if (e != cudaSuccess) p->last_error = (int)e;
  }
}
size_t port_distance(const zpc_port &first, const zpc_port &last) { return (size_t)(last.idx - first.idx); }
}  // namespace

extern "C" {

zpcb200_policy *policy__b200(void) { return new zpcb200_policy; }
void del_policy__b200(zpcb200_policy *p) { delete p; }
void policy_set__b200(zpcb200_policy *p, int device, zpc_stream_t stream, int sync) {
  if (!p) return;
  p->device = device;
  p->stream = (cudaStream_t)stream;
  p->sync = sync;
}
int policy_last_error__b200(const zpcb200_policy *p) { return p ? p->last_error : ZPCB200_E_BADARG; }

#define ZPC_DEF_POLICY_PRIMS(T, S)                                                                                    \
  void reduce_sum__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {                     \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_reduce_sum_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }                                                                                                                   \
  void reduce_prod__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {                    \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_reduce_prod_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }                                                                                                                   \
  void reduce_min__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {                     \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_reduce_min_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }                                                                                                                   \
  void reduce_max__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {                     \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_reduce_max_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }                                                                                                                   \
  void exclusive_scan_sum__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {             \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_exclusive_scan_sum_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }                                                                                                                   \
  void inclusive_scan_sum__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {             \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_inclusive_scan_sum_##S(t, b, first, out, port_distance(first, last), p->stream); }); \
  }
ZPC_DEF_POLICY_PRIMS(int, i32)
ZPC_DEF_POLICY_PRIMS(float, f32)
ZPC_DEF_POLICY_PRIMS(double, f64)

#define ZPC_DEF_POLICY_MERGE(T, S)                                                                                   \
  void merge_sort__b200_##T##_1(zpcb200_policy *p, zpc_port first, zpc_port last) {                                  \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_merge_sort_##S(t, b, first, port_distance(first, last), p->stream); }); \
  }                                                                                                                  \
  void merge_sort_pair__b200_##T##_1(zpcb200_policy *p, zpc_port keys, zpc_port vals, size_t count) {                \
    run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_merge_sort_pair_##S(t, b, keys, vals, count, p->stream); }); \
  }
ZPC_DEF_POLICY_MERGE(int, i32)
ZPC_DEF_POLICY_MERGE(float, f32)
ZPC_DEF_POLICY_MERGE(double, f64)

void radix_sort__b200_int_1(zpcb200_policy *p, zpc_port first, zpc_port last, zpc_port out) {
  run_with_scratch(p, [&](void *t, size_t *b) { return zpcb200_radix_sort_i32(t, b, first, out, port_distance(first, last), 0, 32, p->stream); });
}
void radix_sort_pair__b200_int_1(zpcb200_policy *p, zpc_port keysIn, zpc_port valsIn, zpc_port keysOut, zpc_port valsOut,
                                 size_t count) {
  run_with_scratch(p, [&](void *t, size_t *b) {
    return zpcb200_radix_sort_pair_i32(t, b, keysIn, valsIn, keysOut, valsOut, count, 0, 32, p->stream);
  });
}
}

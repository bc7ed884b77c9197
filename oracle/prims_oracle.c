/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See oracle/oracle.h for the rules.
 *
 * CPU restatement of the reference's SERIAL parallel-primitives
 * (include/zensim/execution/ExecutionPolicy.hpp): inclusive_scan :245-253, exclusive_scan
 * :254-264, reduce :265-274, radix_sort :485-526, radix_sort_pair :527-608 (stable LSD radix sort,
 * 8-bit digits, signed keys biased by flipping the sign bit, a digit pass is skipped when a single
 * bin holds every key).  A stable sort's output is unique, so these are also the expected results
 * of the reference's OpenMP (omp/execution/ExecutionPolicy.hpp:1028-1160) and CUDA/CUB
 * (cuda/execution/ExecutionPolicy.cuh:755-826) paths.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define ZO_RADIX_IMPL(NAME, KT, UKT, SIGNED, PAIR)                                                  \
  void NAME(const KT *kin, PAIR(const int32_t *vin, ) KT *kout, PAIR(int32_t *vout, ) size_t n,     \
            int sbit, int ebit) {                                                                   \
    const UKT flip = SIGNED ? ((UKT)1 << (sizeof(KT) * 8 - 1)) : 0;                                  \
    UKT *cur = (UKT *)malloc(sizeof(UKT) * (n ? n : 1)), *nxt = (UKT *)malloc(sizeof(UKT) * (n ? n : 1)); \
    PAIR(int32_t *cv = (int32_t *)malloc(4 * (n ? n : 1)); int32_t *nv = (int32_t *)malloc(4 * (n ? n : 1));) \
    size_t sizes[256], offs[256];                                                                   \
    int binCount = 256, binMask = 255;                                                              \
    for (size_t i = 0; i < n; ++i) { cur[i] = (UKT)kin[i] ^ flip; PAIR(cv[i] = vin[i];) }            \
    for (int st = sbit; st < ebit; st += 8) {                                                       \
      if (st + 8 > ebit) { binMask >>= (st + 8 - ebit); binCount >>= (st + 8 - ebit); }             \
      memset(sizes, 0, sizeof sizes);                                                               \
      for (size_t i = 0; i < n; ++i) sizes[(cur[i] >> st) & binMask]++;                             \
      int skip = sizes[0] == n;                                                                     \
      offs[0] = 0;                                                                                  \
      for (int b = 1; b < binCount; ++b) {                                                          \
        if (sizes[b] == n) { skip = 1; break; }                                                     \
        offs[b] = offs[b - 1] + sizes[b - 1];                                                       \
      }                                                                                             \
      if (skip) continue;                                                                           \
      for (int b = 0; b < binCount; ++b) sizes[b] += offs[b];                                       \
      for (size_t i = n; i-- > 0;) { /* backward, stable (:592-597) */                              \
        size_t loc = --sizes[(cur[i] >> st) & binMask];                                             \
        nxt[loc] = cur[i]; PAIR(nv[loc] = cv[i];)                                                   \
      }                                                                                             \
      { UKT *t = cur; cur = nxt; nxt = t; } PAIR({ int32_t *t = cv; cv = nv; nv = t; })              \
    }                                                                                               \
    for (size_t i = 0; i < n; ++i) { kout[i] = (KT)(cur[i] ^ flip); PAIR(vout[i] = cv[i];) }         \
    free(cur); free(nxt); PAIR(free(cv); free(nv);)                                                 \
  }
#define ZO_YES(...) __VA_ARGS__
#define ZO_NO(...)
ZO_RADIX_IMPL(zo_radix_sort_pair_u32, uint32_t, uint32_t, 0, ZO_YES)
ZO_RADIX_IMPL(zo_radix_sort_pair_i32, int32_t, uint32_t, 1, ZO_YES)
ZO_RADIX_IMPL(zo_radix_sort_pair_u64, uint64_t, uint64_t, 0, ZO_YES)
ZO_RADIX_IMPL(zo_radix_sort_u32, uint32_t, uint32_t, 0, ZO_NO)
ZO_RADIX_IMPL(zo_radix_sort_i32, int32_t, uint32_t, 1, ZO_NO)
ZO_RADIX_IMPL(zo_radix_sort_u64, uint64_t, uint64_t, 0, ZO_NO)

/* scan / reduce: strictly left-to-right folds (ExecutionPolicy.hpp:245-274).  Identities as the
 * reference's C ABI passes them (py_interop/cuda/ExecutionPolicy.cpp:41-68): 0 for sum, 1 for prod,
 * numeric max for min, numeric lowest for max. */
#define ZO_SCAN_REDUCE_IMPL(S, T, TMAX, TLOW)                                  \
  void zo_exclusive_scan_sum_##S(const T *in, T *out, size_t n) {             \
    T acc = 0;                                                                \
    for (size_t i = 0; i < n; ++i) { T x = in[i]; out[i] = acc; acc = acc + x; } \
  }                                                                           \
  void zo_inclusive_scan_sum_##S(const T *in, T *out, size_t n) {             \
    if (!n) return;                                                           \
    T acc = in[0]; out[0] = acc;                                              \
    for (size_t i = 1; i < n; ++i) { acc = acc + in[i]; out[i] = acc; }       \
  }                                                                           \
  void zo_reduce_sum_##S(const T *in, T *out, size_t n) {                     \
    T acc = 0;                                                                \
    for (size_t i = 0; i < n; ++i) acc = acc + in[i];                         \
    *out = acc;                                                               \
  }                                                                           \
  void zo_reduce_prod_##S(const T *in, T *out, size_t n) {                    \
    T acc = 1;                                                                \
    for (size_t i = 0; i < n; ++i) acc = acc * in[i];                         \
    *out = acc;                                                               \
  }                                                                           \
  void zo_reduce_min_##S(const T *in, T *out, size_t n) {                     \
    T acc = TMAX;                                                             \
    for (size_t i = 0; i < n; ++i) acc = acc < in[i] ? acc : in[i];           \
    *out = acc;                                                               \
  }                                                                           \
  void zo_reduce_max_##S(const T *in, T *out, size_t n) {                     \
    T acc = TLOW;                                                             \
    for (size_t i = 0; i < n; ++i) acc = acc > in[i] ? acc : in[i];           \
    *out = acc;                                                               \
  }
#include <float.h>
#include <limits.h>
ZO_SCAN_REDUCE_IMPL(i32, int32_t, INT32_MAX, INT32_MIN)
ZO_SCAN_REDUCE_IMPL(u32, uint32_t, UINT32_MAX, 0u)
ZO_SCAN_REDUCE_IMPL(i64, int64_t, INT64_MAX, INT64_MIN)
ZO_SCAN_REDUCE_IMPL(f32, float, FLT_MAX, -FLT_MAX)
ZO_SCAN_REDUCE_IMPL(f64, double, DBL_MAX, -DBL_MAX)

/* merge_sort_pair (stable, ascending, operator<): execution/ExecutionPolicy.hpp:311-455 — bottom-up merge; the result
 * of a STABLE sort under a strict weak order is unique, so any stable merge gives the reference's output. */
#define ZO_MERGE_IMPL(S, T)                                                                       \
  void zo_merge_sort_pair_##S(T *keys, int32_t *vals, size_t n) {                                 \
    T *k2 = (T *)malloc(sizeof(T) * (n ? n : 1));                                                 \
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));                               \
    T *ka = keys, *kb = k2;                                                                       \
    int32_t *va = vals, *vb = v2;                                                                 \
    for (size_t w = 1; w < n; w *= 2) {                                                           \
      for (size_t lo = 0; lo < n; lo += 2 * w) {                                                  \
        size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;               \
        size_t i = lo, j = mid, o = lo;                                                           \
        while (i < mid && j < hi) {                                                               \
          if (ka[j] < ka[i]) { kb[o] = ka[j]; vb[o++] = va[j++]; }                                \
          else { kb[o] = ka[i]; vb[o++] = va[i++]; }                                              \
        }                                                                                         \
        while (i < mid) { kb[o] = ka[i]; vb[o++] = va[i++]; }                                     \
        while (j < hi) { kb[o] = ka[j]; vb[o++] = va[j++]; }                                      \
      }                                                                                           \
      { T *t = ka; ka = kb; kb = t; }                                                             \
      { int32_t *t = va; va = vb; vb = t; }                                                       \
    }                                                                                             \
    if (ka != keys) for (size_t i = 0; i < n; ++i) { keys[i] = ka[i]; vals[i] = va[i]; }          \
    free(k2); free(v2);                                                                           \
  }
ZO_MERGE_IMPL(i32, int32_t)
ZO_MERGE_IMPL(u32, uint32_t)
ZO_MERGE_IMPL(i64, int64_t)
ZO_MERGE_IMPL(u64, uint64_t)
ZO_MERGE_IMPL(f32, float)
ZO_MERGE_IMPL(f64, double)

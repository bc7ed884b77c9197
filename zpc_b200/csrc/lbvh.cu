// LBvh<3, int, f32>::build / refit (reference container/Bvh.hpp:835-1000, 1229-1259) on the sm_100a primitives of prims.cu —
// the biggest in-tree consumer of reduce / radix_sort_pair / exclusive_scan.
//
//   whole box     six strided reduce_min / reduce_max over the primitive boxes (compute_bounding_box, Bvh.hpp:42-80; the
//                 10-eps padding commutes with min / max, so it is applied once to the result)
//   Morton codes  one thread per primitive (:177-187), keys u32, values = primitive ids
//   sort          zpcb200_radix_sort_pair_u32 (stable, all 32 bits — :890-893)
//   topology      Karras 2012, one thread per internal node (:198-287); a leaf's depth = 1 + the number of internal nodes
//                 whose range starts at it, counted with atomics
//   layout        exclusive_scan of the depths (:915) gives every leaf its slot in DFS pre-order; the internal nodes above a
//                 leaf fill the slots before it (:288-337)
//   refit         bottom-up, the second thread to reach a node merges its children (:467-491)
//
// Results are bit-identical to the reference: the topology is a function of the sorted codes, box merging is min / max.
// The whole box stays on the device (the reference reads it back to the host, :887).
#include <cfloat>
#include <climits>

#include "common.cuh"
#include "lbvh_core.cuh"

namespace {

// box = {min of mins, max of maxs} -> padded by 10 eps (Bvh.hpp:18-21)
__global__ void lbvh_pad_box_kernel(float *box) {
  const int d = threadIdx.x;
  if (d < 3) box[d] = box[d] - 10 * FLT_EPSILON;
  else if (d < 6) box[d] = box[d] + 10 * FLT_EPSILON;
}
__global__ void lbvh_morton_kernel(const float *__restrict__ prims, const float *__restrict__ box, int n, unsigned *codes, int *ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  codes[i] = zpcb::morton_of(prims, box, i);
  ids[i] = i;
}
// _build_init_depths (:188-197)
__global__ void lbvh_init_depths_kernel(int n, int *lDepths) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lDepths[i] = 1;
  else if (i == n) lDepths[n] = 0;
}
__global__ void lbvh_topo_kernel(const unsigned *__restrict__ mcs, int numTrunk, int *tPars, int *tRs, int *lPars, int *lDepths) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < numTrunk) zpcb::topo_node(idx, mcs, numTrunk, tPars, tRs, lPars, lDepths);
}
__global__ void lbvh_supp_topo_kernel(int n, const int *__restrict__ lOffsets, const int *__restrict__ lPars, const int *tPars,
                                      const int *__restrict__ pInds, int *tDst, int *lLcas, int *levels, int *auxIndices, int *leafInds) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) zpcb::supp_topo_leaf(idx, n, lOffsets, lPars, tPars, pInds, tDst, lLcas, levels, auxIndices, leafInds);
}
__global__ void lbvh_reorder_kernel(int n, const int *__restrict__ lOffsets, const int *__restrict__ lPars, const int *__restrict__ lLcas,
                                    const int *__restrict__ tPars, const int *__restrict__ tRs, const int *__restrict__ tDst,
                                    int *auxIndices, int *parents) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) zpcb::reorder_node(idx, n, lOffsets, lPars, lLcas, tPars, tRs, tDst, auxIndices, parents);
}

// _refit_bottom_up (:467-491).  The first thread to reach an internal node leaves; the second one finds both children
// written (release: fence before the flag; acquire: fence after it, children read past L1).
__global__ void lbvh_refit_kernel(int n, const float *__restrict__ prims, float *bvs, const int *__restrict__ auxIndices,
                                  const int *__restrict__ leafInds, const int *__restrict__ parents, const int *__restrict__ levels,
                                  int *flags) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int node = leafInds[idx];
  const int prim = auxIndices[node];
#pragma unroll
  for (int d = 0; d < 6; ++d) bvs[6 * (size_t)node + d] = prims[6 * (size_t)prim + d];
  node = parents[node];
  for (int guard = 0; node >= 0 && guard < 2 * n; ++guard) {  // a parent chain is at most n - 1 long; the bound keeps garbage finite
    __threadfence();
    if (atomicCAS(&flags[node], 0, 1) == 0) break;
    __threadfence();
    const int lc = node + 1, rc = levels[lc] ? auxIndices[lc] : lc + 1;
    float out[6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float a0 = __ldcg(bvs + 6 * (size_t)lc + d), a1 = __ldcg(bvs + 6 * (size_t)lc + 3 + d);
      const float b0 = __ldcg(bvs + 6 * (size_t)rc + d), b1 = __ldcg(bvs + 6 * (size_t)rc + 3 + d);
      out[d] = fminf(fminf(a0, b0), b1);       // merge(bv, rbv._min); merge(bv, rbv._max)
      out[3 + d] = fmaxf(fmaxf(a1, b0), b1);
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) bvs[6 * (size_t)node + d] = out[d];
    node = parents[node];
  }
}

// Batched LBvhView::iter_neighbors: one thread per query box.  out == nullptr: counts[q] = number of overlapping primitives;
// otherwise the primitive ids go to out[offsets[q] ...] in visiting order (count -> exclusive_scan -> fill).
__global__ void lbvh_query_kernel(int numLeaves, const float *__restrict__ bvs, const int *__restrict__ auxIndices,
                                  const int *__restrict__ levels, const float *__restrict__ queries, int nq, int *counts,
                                  const int *__restrict__ offsets, int *out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float bv[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) bv[d] = queries[6 * (size_t)q + d];
  int c = 0;
  if (out) {
    int *dst = out + offsets[q];
    zpcb::iter_neighbors(numLeaves, bvs, auxIndices, levels, bv, [&](int prim) { dst[c++] = prim; });
  } else {
    zpcb::iter_neighbors(numLeaves, bvs, auxIndices, levels, bv, [&](int) { ++c; });
    counts[q] = c;
  }
}

// n <= 2 (Bvh.hpp:845-853): no tree, the boxes in input order
__global__ void lbvh_tiny_kernel(int n, const float *prims, float *bvs, int *auxIndices, int *leafInds, int ids) {
  const int i = threadIdx.x;
  if (i < 6 * n) bvs[i] = prims[i];
  if (ids && i < n) { auxIndices[i] = i; leafInds[i] = i; }
}

struct Layout {
  size_t codes, ids, smcs, pInds, tPars, tRs, tDst, lPars, lLcas, lDepths, lOffsets, box, flags, scan, sort, red, total;
  size_t scan_bytes, sort_bytes, red_bytes;
};
int lbvh_layout(size_t n, Layout &L) {
  zpc_port none = {nullptr, 0, 0, 0, 1};
  int rc = zpcb200_radix_sort_pair_u32(nullptr, &L.sort_bytes, none, none, none, none, n, 0, 32, nullptr);
  if (rc) return rc;
  rc = zpcb200_exclusive_scan_sum_i32(nullptr, &L.scan_bytes, none, none, n + 1, nullptr);
  if (rc) return rc;
  rc = zpcb200_reduce_min_f32(nullptr, &L.red_bytes, none, none, n, nullptr);
  if (rc) return rc;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = zpc_align_up(o + bytes, 256); return at; };
  L.codes = take(4 * n); L.ids = take(4 * n); L.smcs = take(4 * n); L.pInds = take(4 * n);
  L.tPars = take(4 * n); L.tRs = take(4 * n); L.tDst = take(4 * n); L.lPars = take(4 * n); L.lLcas = take(4 * n);
  L.lDepths = take(4 * (n + 1)); L.lOffsets = take(4 * (n + 1)); L.box = take(32); L.flags = take(4 * 2 * n);
  L.scan = take(L.scan_bytes); L.sort = take(L.sort_bytes); L.red = take(L.red_bytes);
  L.total = o;
  return ZPCB200_OK;
}

int lbvh_refit_launch(const float *prims, int n, zpc_lbvh_view bvh, int *flags, cudaStream_t s) {
  ZPC_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (2 * (size_t)n - 1), s));
  lbvh_refit_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, prims, bvh.orderedBvs, bvh.auxIndices, bvh.leafInds, bvh.parents, bvh.levels, flags);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}

}  // namespace

extern "C" {

int zpcb200_lbvh_build(void *temp, size_t *temp_bytes, const float *primBvs, size_t numLeaves, zpc_lbvh_view bvh, int refit,
                       zpc_stream_t stream) {
  if (!temp_bytes) return ZPCB200_E_BADARG;
  if (numLeaves > ((size_t)1 << 30)) return ZPCB200_E_UNSUPPORTED;
  Layout L;
  if (numLeaves <= 2) {
    if (!temp) { *temp_bytes = 256; return ZPCB200_OK; }
  } else {
    const int rc = lbvh_layout(numLeaves, L);
    if (rc) return rc;
    if (!temp) { *temp_bytes = L.total; return ZPCB200_OK; }
    if (*temp_bytes < L.total) return ZPCB200_E_TEMP_TOO_SMALL;
  }
  if (numLeaves == 0) return ZPCB200_OK;
  if (!primBvs || !bvh.orderedBvs || !bvh.auxIndices || !bvh.leafInds) return ZPCB200_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)numLeaves;
  if (n <= 2) {
    lbvh_tiny_kernel<<<1, 32, 0, s>>>(n, primBvs, bvh.orderedBvs, bvh.auxIndices, bvh.leafInds, 1);
    ZPC_CHECK_LAUNCH();
    return ZPCB200_OK;
  }
  if (!bvh.parents || !bvh.levels) return ZPCB200_E_BADARG;
  char *t = (char *)temp;
  unsigned *codes = (unsigned *)(t + L.codes), *smcs = (unsigned *)(t + L.smcs);
  int *ids = (int *)(t + L.ids), *pInds = (int *)(t + L.pInds), *tPars = (int *)(t + L.tPars), *tRs = (int *)(t + L.tRs),
      *tDst = (int *)(t + L.tDst), *lPars = (int *)(t + L.lPars), *lLcas = (int *)(t + L.lLcas), *lDepths = (int *)(t + L.lDepths),
      *lOffsets = (int *)(t + L.lOffsets), *flags = (int *)(t + L.flags);
  float *box = (float *)(t + L.box);
  for (int d = 0; d < 6; ++d) {  // component d of every box: a port with six channels and tile length 1
    zpc_port in = {(void *)(primBvs + d), 0, 0, 0, 6}, out = {box + d, 0, 0, 0, 1};
    size_t rb = L.red_bytes;
    const int rc = d < 3 ? zpcb200_reduce_min_f32(t + L.red, &rb, in, out, numLeaves, stream)
                         : zpcb200_reduce_max_f32(t + L.red, &rb, in, out, numLeaves, stream);
    if (rc) return rc;
  }
  lbvh_pad_box_kernel<<<1, 32, 0, s>>>(box);
  ZPC_CHECK_LAUNCH();
  const unsigned gl = (unsigned)((n + 1 + 255) / 256);
  lbvh_morton_kernel<<<gl, 256, 0, s>>>(primBvs, box, n, codes, ids);
  ZPC_CHECK_LAUNCH();
  {
    zpc_port pk = {codes, 0, 0, 0, 1}, pv = {ids, 0, 0, 0, 1}, psk = {smcs, 0, 0, 0, 1}, psv = {pInds, 0, 0, 0, 1};
    size_t sb = L.sort_bytes;
    const int rc = zpcb200_radix_sort_pair_u32(t + L.sort, &sb, pk, pv, psk, psv, numLeaves, 0, 32, stream);
    if (rc) return rc;
  }
  lbvh_init_depths_kernel<<<gl, 256, 0, s>>>(n, lDepths);
  ZPC_CHECK_LAUNCH();
  lbvh_topo_kernel<<<gl, 256, 0, s>>>(smcs, n - 1, tPars, tRs, lPars, lDepths);
  ZPC_CHECK_LAUNCH();
  {
    zpc_port pi = {lDepths, 0, 0, 0, 1}, po = {lOffsets, 0, 0, 0, 1};
    size_t sb = L.scan_bytes;
    const int rc = zpcb200_exclusive_scan_sum_i32(t + L.scan, &sb, pi, po, numLeaves + 1, stream);
    if (rc) return rc;
  }
  lbvh_supp_topo_kernel<<<gl, 256, 0, s>>>(n, lOffsets, lPars, tPars, pInds, tDst, lLcas, bvh.levels, bvh.auxIndices, bvh.leafInds);
  ZPC_CHECK_LAUNCH();
  lbvh_reorder_kernel<<<gl, 256, 0, s>>>(n, lOffsets, lPars, lLcas, tPars, tRs, tDst, bvh.auxIndices, bvh.parents);
  ZPC_CHECK_LAUNCH();
  if (refit) return lbvh_refit_launch(primBvs, n, bvh, flags, s);
  return ZPCB200_OK;
}

int zpcb200_lbvh_refit(void *temp, size_t *temp_bytes, const float *primBvs, size_t numLeaves, zpc_lbvh_view bvh, zpc_stream_t stream) {
  if (!temp_bytes) return ZPCB200_E_BADARG;
  if (numLeaves > ((size_t)1 << 30)) return ZPCB200_E_UNSUPPORTED;
  const size_t need = zpc_align_up(4 * 2 * (numLeaves ? numLeaves : 1), 256);
  if (!temp) { *temp_bytes = need; return ZPCB200_OK; }
  if (*temp_bytes < need) return ZPCB200_E_TEMP_TOO_SMALL;
  if (numLeaves == 0) return ZPCB200_OK;
  if (!primBvs || !bvh.orderedBvs) return ZPCB200_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (numLeaves <= 2) {  // :1241-1244
    lbvh_tiny_kernel<<<1, 32, 0, s>>>((int)numLeaves, primBvs, bvh.orderedBvs, nullptr, nullptr, 0);
    ZPC_CHECK_LAUNCH();
    return ZPCB200_OK;
  }
  if (!bvh.auxIndices || !bvh.leafInds || !bvh.parents || !bvh.levels) return ZPCB200_E_BADARG;
  return lbvh_refit_launch(primBvs, (int)numLeaves, bvh, (int *)temp, s);
}

int zpcb200_lbvh_query(zpc_lbvh_view bvh, size_t numLeaves, const float *queryBvs, size_t numQueries, int *counts, const int *offsets,
                       int *out, zpc_stream_t stream) {
  if (numLeaves > ((size_t)1 << 30) || numQueries > (size_t)INT_MAX) return ZPCB200_E_UNSUPPORTED;
  if (numQueries == 0) return ZPCB200_OK;
  if (!queryBvs || (!out && !counts) || (out && !offsets)) return ZPCB200_E_BADARG;
  if (numLeaves && (!bvh.orderedBvs || !bvh.auxIndices || (numLeaves > 2 && !bvh.levels))) return ZPCB200_E_BADARG;
  lbvh_query_kernel<<<(unsigned)((numQueries + 127) / 128), 128, 0, (cudaStream_t)stream>>>((int)numLeaves, bvh.orderedBvs, bvh.auxIndices,
                                                                                          bvh.levels, queryBvs, (int)numQueries, counts,
                                                                                          offsets, out);
  ZPC_CHECK_LAUNCH();
  return ZPCB200_OK;
}
}

/* TEST INFRASTRUCTURE — NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of LBvh<3, int, f32>::build / refit and of the stack-free traversal of LBvhView
 * (reference container/Bvh.hpp; paths relative to include/zensim/) — the biggest in-tree consumer of the parallel
 * primitives: Morton codes -> radix_sort_pair -> Karras topology -> exclusive_scan -> DFS-order layout.
 * Boxes are AABBBox<3,f32> = {min[3], max[3]}, six floats each.
 */
#include "oracle.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

static uint32_t expand_bits_32(uint32_t v) { /* math/bit/Bits.h:83-89 */
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static uint32_t morton_3d_32(float x, float y, float z) { /* :121-124 */
  return (expand_bits_32((uint32_t)(x * 1024.f)) << 2) | (expand_bits_32((uint32_t)(y * 1024.f)) << 1)
         | expand_bits_32((uint32_t)(z * 1024.f));
}
static int count_lz(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); } /* execution/Intrinsics.hpp:330-351 */

/* compute_bounding_box (Bvh.hpp:11-24, 42-80): every box padded by 10 eps, min / max over all of them */
void zo_lbvh_whole_box(int n, const float *bvs, float box[6]) {
  for (int d = 0; d < 3; ++d) { box[d] = FLT_MAX; box[3 + d] = -FLT_MAX; }
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      const float lo = bvs[6 * i + d] - 10 * FLT_EPSILON, hi = bvs[6 * i + 3 + d] + 10 * FLT_EPSILON;
      if (lo < box[d]) box[d] = lo;
      if (hi > box[3 + d]) box[3 + d] = hi;
    }
}
/* _build_init_mc_id (Bvh.hpp:177-187): Morton code of the box centre in the whole box's unit coordinates
 * (geometry/BoundingVolumeInterface.hpp:12-15, 24-31) */
uint32_t zo_lbvh_morton(const float box[6], const float bv[6]) {
  float coord[3];
  for (int d = 0; d < 3; ++d) {
    const float c = (bv[d] + bv[3 + d]) / 2;
    const float length = box[3 + d] - box[d];
    float off = c - box[d];
    off = off < 0.f ? 0.f : (off > length ? length : off); /* math::clamp */
    coord[d] = off / length;
  }
  return morton_3d_32(coord[0], coord[1], coord[2]);
}

/* _refit_bottom_up (Bvh.hpp:467-491), serial: leaves first, then every trunk node after its two children — the node
 * order is DFS pre-order, so a reverse sweep over the nodes visits children before parents */
void zo_lbvh_refit(int n, const float *primBvs, float *orderedBvs, const int *auxIndices, const int *parents,
                   const int *levels, const int *leafInds) {
  (void)parents;
  if (n <= 2) { memcpy(orderedBvs, primBvs, sizeof(float) * 6 * (size_t)n); return; } /* :1241-1244 */
  const int numNodes = 2 * n - 1;
  for (int i = 0; i < n; ++i) memcpy(orderedBvs + 6 * (size_t)leafInds[i], primBvs + 6 * (size_t)auxIndices[leafInds[i]], sizeof(float) * 6);
  for (int node = numNodes - 1; node >= 0; --node) {
    if (levels[node] == 0) continue;
    const int lc = node + 1, rc = levels[lc] ? auxIndices[lc] : lc + 1;   /* :478-479 */
    float *bv = orderedBvs + 6 * (size_t)node;
    const float *a = orderedBvs + 6 * (size_t)lc, *b = orderedBvs + 6 * (size_t)rc;
    for (int d = 0; d < 3; ++d) {                                          /* merge(bv, rbv._min); merge(bv, rbv._max) */
      float lo = a[d], hi = a[3 + d];
      if (b[d] < lo) lo = b[d];
      if (b[d] > hi) hi = b[d];
      if (b[3 + d] < lo) lo = b[3 + d];
      if (b[3 + d] > hi) hi = b[3 + d];
      bv[d] = lo; bv[3 + d] = hi;
    }
  }
}

/* LBvh::build (Bvh.hpp:835-1000) with the functors of :177-337.  Outputs: auxIndices / parents / levels [2n-1], leafInds [n],
 * orderedBvs [2n-1][6] (filled when refit != 0).  n <= 2: the degenerate layout of :845-853. */
void zo_lbvh_build(int n, const float *primBvs, float *orderedBvs, int *auxIndices, int *parents, int *levels, int *leafInds,
                   int refit) {
  if (n == 0) return;
  if (n <= 2) {
    memcpy(orderedBvs, primBvs, sizeof(float) * 6 * (size_t)n);
    for (int i = 0; i < n; ++i) { leafInds[i] = i; auxIndices[i] = i; }
    return;
  }
  const int numTrunk = n - 1;
  float box[6];
  zo_lbvh_whole_box(n, primBvs, box);
  uint32_t *mcs = malloc(sizeof(uint32_t) * n), *smcs = malloc(sizeof(uint32_t) * n);
  int32_t *ids = malloc(sizeof(int32_t) * n), *pInds = malloc(sizeof(int32_t) * n);
  for (int i = 0; i < n; ++i) { mcs[i] = zo_lbvh_morton(box, primBvs + 6 * (size_t)i); ids[i] = i; }
  zo_radix_sort_pair_u32(mcs, ids, smcs, pInds, (size_t)n, 0, 32);            /* :890-893; pInds = sortedIndices (:188-196) */
  int *tPars = malloc(sizeof(int) * numTrunk), *tLcs = malloc(sizeof(int) * numTrunk), *tRcs = malloc(sizeof(int) * numTrunk),
      *tLs = malloc(sizeof(int) * numTrunk), *tRs = malloc(sizeof(int) * numTrunk), *tDst = malloc(sizeof(int) * numTrunk),
      *lPars = malloc(sizeof(int) * n), *lLcas = malloc(sizeof(int) * n), *lDepths = malloc(sizeof(int) * (n + 1)),
      *lOffsets = malloc(sizeof(int) * (n + 1));
  for (int i = 0; i < n; ++i) lDepths[i] = 1;
  lDepths[numTrunk + 1] = 0;
  const int num_leaves = n;
  for (int idx = 0; idx < numTrunk; ++idx) {                                 /* _build_build_topo, :198-287 (Karras 2012) */
    int i = 0, j = 0;
    if (idx == 0) {
      i = 0;
      j = num_leaves - 1;
    } else {
      int left = idx, right = idx, dir = 0;
      uint32_t minLZ = 0;
      const uint32_t preCode = smcs[idx - 1], curCode = smcs[idx], nxtCode = smcs[idx + 1];
      if (preCode == curCode && curCode == nxtCode) {
        for (++right; right < num_leaves - 1; ++right)
          if (smcs[right] != smcs[right + 1]) break;
        j = right;
        i = left;
      } else {
        const uint32_t lLZ = (uint32_t)count_lz(preCode ^ curCode), rLZ = (uint32_t)count_lz(nxtCode ^ curCode);
        if (lLZ > rLZ) { dir = -1; minLZ = rLZ; }
        else { dir = 1; minLZ = lLZ; }
        int step;
        for (step = 2; right = left + step * dir,
            (right < num_leaves && right >= 0 ? (uint32_t)count_lz(smcs[right] ^ curCode) > minLZ : 0);
             step <<= 1);
        int len;
        for (len = 0, step >>= 1; step >= 1; step >>= 1) {
          right = left + (len + step) * dir;
          if (right < num_leaves && right >= 0)
            if ((uint32_t)count_lz(smcs[right] ^ curCode) > minLZ) len += step;
        }
        if (dir == 1) { i = left; j = left + len; }
        else { i = left - len; j = left; }
      }
    }
    lDepths[i] += 1;
    tLs[idx] = i;
    tRs[idx] = j;
    int gamma;
    const uint32_t lCode = smcs[i], rCode = smcs[j];
    if (lCode == rCode) gamma = i;
    else {
      const int LZ = count_lz(lCode ^ rCode);
      int step, len;
      for (step = (j - i + 1) >> 1, len = 0; 1; step = (step + 1) >> 1) {
        if (i + len + step > numTrunk) continue;
        if (count_lz(smcs[i + len + step] ^ lCode) > LZ) len += step;
        if (step <= 1) break;
      }
      gamma = i + len;
    }
    tLcs[idx] = gamma;
    tRcs[idx] = gamma + 1;
    const int mi = i < j ? i : j, ma = i > j ? i : j;
    if (mi == gamma) { lPars[gamma] = idx; tLcs[idx] += numTrunk; }
    else tPars[gamma] = idx;
    if (ma == gamma + 1) { lPars[gamma + 1] = idx; tRcs[idx] += numTrunk; }
    else tPars[gamma + 1] = idx;
    if (idx == 0) tPars[0] = -1;
  }
  { int run = 0; for (int i = 0; i <= n; ++i) { lOffsets[i] = run; run += lDepths[i]; } }   /* exclusive_scan, :915 */
  for (int idx = 0; idx < n; ++idx) {                                          /* _build_supp_topo, :288-303 */
    int depth = lOffsets[idx + 1] - lOffsets[idx];
    int dst = lOffsets[idx + 1] - 2;
    int node = lPars[idx], ch = idx + numTrunk, level = 0;
    for (; --depth; node = tPars[node], --dst) {
      tDst[node] = dst;
      levels[dst] = ++level;
      ch = node;
    }
    lLcas[idx] = ch;
  }
  for (int idx = 0; idx < n; ++idx) {                                          /* _build_reorder_leaf, :304-318 */
    const int dst = lOffsets[idx + 1] - 1;
    auxIndices[dst] = pInds[idx];
    parents[dst] = tDst[lPars[idx]];
    levels[dst] = 0;
    leafInds[idx] = dst;
  }
  for (int idx = 0; idx < numTrunk; ++idx) {                                   /* _build_reorder_trunk, :319-337 */
    const int dst = tDst[idx], r = tRs[idx];
    if (r != numTrunk) {
      const int lca = lLcas[r + 1];
      auxIndices[dst] = lca < numTrunk ? tDst[lca] : lOffsets[r + 1];
    } else
      auxIndices[dst] = -1;
    parents[dst] = idx != 0 ? tDst[tPars[idx]] : -1;
  }
  free(mcs); free(smcs); free(ids); free(pInds); free(tPars); free(tLcs); free(tRcs); free(tLs); free(tRs); free(tDst);
  free(lPars); free(lLcas); free(lDepths); free(lOffsets);
  if (refit) zo_lbvh_refit(n, primBvs, orderedBvs, auxIndices, parents, levels, leafInds);
}

/* LBvhView::iter_neighbors (Bvh.hpp:660-689): stack-free traversal along the DFS order with escape indices; writes the
 * primitive ids whose boxes overlap bv (in visiting order) and returns their number */
static int overlaps(const float *a, const float *b) { /* geometry/BoundingVolumeInterface.hpp overlaps(AABB, AABB) */
  for (int d = 0; d < 3; ++d)
    if (b[d] > a[3 + d] || b[3 + d] < a[d]) return 0;
  return 1;
}
int zo_lbvh_iter_neighbors(int n, const float *orderedBvs, const int *auxIndices, const int *levels, const float bv[6], int *out,
                           int cap) {
  int cnt = 0;
  if (n <= 2) {
    for (int i = 0; i < n; ++i)
      if (overlaps(orderedBvs + 6 * i, bv)) { if (cnt < cap) out[cnt] = i; ++cnt; }
    return cnt;
  }
  const int numNodes = 2 * n - 1;
  int node = 0;
  while (node != -1 && node != numNodes) {
    int level = levels[node];
    for (; level; --level, ++node)
      if (!overlaps(orderedBvs + 6 * (size_t)node, bv)) break;
    if (level == 0) {
      if (overlaps(orderedBvs + 6 * (size_t)node, bv)) { if (cnt < cap) out[cnt] = auxIndices[node]; ++cnt; }
      node++;
    } else
      node = auxIndices[node];
  }
  return cnt;
}

// Per-node functions of the LBvh build (reference container/Bvh.hpp:177-337), __host__ __device__ so that tests/hostmath can
// run them on the CPU against the oracle; lbvh.cu wraps each in a one-thread-per-index kernel.
#pragma once
#include "mpm_math.cuh"  // ZPC_HD, rn_div / rn_mul

namespace zpcb {

ZPC_HD int clz32(unsigned x) {  // count_lz (execution/Intrinsics.hpp:298-351): 32 for 0 on both sides
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
ZPC_HD void count_one(int *p) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, 1);
#else
  ++*p;
#endif
}
ZPC_HD unsigned expand_bits_32(unsigned v) {  // math/bit/Bits.h:83-89
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// _build_init_mc_id (Bvh.hpp:177-187) with getBoxCenter / getUniformCoord (geometry/BoundingVolumeInterface.hpp:12-31)
ZPC_HD unsigned morton_of(const float *prims, const float *box, int i) {
  unsigned q[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float c = (prims[6 * (size_t)i + d] + prims[6 * (size_t)i + 3 + d]) / 2;
    const float length = box[3 + d] - box[d];
    float off = c - box[d];
    off = off < 0.f ? 0.f : (off > length ? length : off);
    q[d] = (unsigned)(zpcm::rn_mul(zpcm::rn_div(off, length), 1024.f));
  }
  return (expand_bits_32(q[0]) << 2) | (expand_bits_32(q[1]) << 1) | expand_bits_32(q[2]);
}

// _build_build_topo (:198-287): range of internal node idx, its split, the parent links of its two children
ZPC_HD void topo_node(int idx, const unsigned *mcs, int numTrunk, int *tPars, int *tRs, int *lPars, int *lDepths) {
  const int num_leaves = numTrunk + 1;
  int i = 0, j = 0;
  if (idx == 0) {
    j = num_leaves - 1;
  } else {
    const int left = idx;
    int right = idx;
    const unsigned preCode = mcs[idx - 1], curCode = mcs[idx], nxtCode = mcs[idx + 1];
    if (preCode == curCode && curCode == nxtCode) {
      for (++right; right < num_leaves - 1; ++right)
        if (mcs[right] != mcs[right + 1]) break;
      i = left;
      j = right;
    } else {
      const int lLZ = clz32(preCode ^ curCode), rLZ = clz32(nxtCode ^ curCode);
      const int dir = lLZ > rLZ ? -1 : 1, minLZ = lLZ > rLZ ? rLZ : lLZ;
      int step = 2;
      while (true) {  // exponential search for the other end of the range
        right = left + step * dir;
        if (!(right < num_leaves && right >= 0 && clz32(mcs[right] ^ curCode) > minLZ)) break;
        step <<= 1;
      }
      int len = 0;
      for (step >>= 1; step >= 1; step >>= 1) {  // binary search
        right = left + (len + step) * dir;
        if (right < num_leaves && right >= 0 && clz32(mcs[right] ^ curCode) > minLZ) len += step;
      }
      if (dir == 1) { i = left; j = left + len; }
      else { i = left - len; j = left; }
    }
  }
  count_one(&lDepths[i]);
  tRs[idx] = j;
  int gamma;
  const unsigned lCode = mcs[i], rCode = mcs[j];
  if (lCode == rCode) {
    gamma = i;
  } else {
    const int LZ = clz32(lCode ^ rCode);
    int len = 0;
    for (int step = (j - i + 1) >> 1;; step = (step + 1) >> 1) {
      // the reference skips the probe (and the exit test) when i + len + step > numTrunk; len < j - i means that can only
      // happen with step >= 2, where the exit test is false anyway: guarding the probe alone is equivalent
      if (i + len + step <= numTrunk && clz32(mcs[i + len + step] ^ lCode) > LZ) len += step;
      if (step <= 1) break;
    }
    gamma = i + len;
  }
  if (i == gamma) lPars[gamma] = idx;          // i <= j always: mi = i, ma = j
  else tPars[gamma] = idx;
  if (j == gamma + 1) lPars[gamma + 1] = idx;
  else tPars[gamma + 1] = idx;
  if (idx == 0) tPars[0] = -1;
}

// _build_supp_topo + _build_reorder_leaf (:288-318): the internal nodes whose range starts at leaf idx take the slots right
// before the leaf's, top-most first; the leaf records its primitive id and its slot (its parent's slot may be assigned by
// another leaf: reorder_node, afterwards)
ZPC_HD void supp_topo_leaf(int idx, int n, const int *lOffsets, const int *lPars, const int *tPars, const int *pInds, int *tDst,
                           int *lLcas, int *levels, int *auxIndices, int *leafInds) {
  const int numTrunk = n - 1;
  int depth = lOffsets[idx + 1] - lOffsets[idx];
  if (depth < 1 || depth > n) return;  // cannot happen on a consistent scan; never walk on garbage
  int dst = lOffsets[idx + 1] - 2;
  int node = lPars[idx], ch = idx + numTrunk, level = 0;
  for (; --depth; node = tPars[node], --dst) {
    tDst[node] = dst;
    levels[dst] = ++level;
    ch = node;
  }
  lLcas[idx] = ch;
  const int slot = lOffsets[idx + 1] - 1;
  auxIndices[slot] = pInds[idx];
  levels[slot] = 0;
  leafInds[idx] = slot;
}
// parents of the leaves (:311) and _build_reorder_trunk (:319-337): escape index and parent of every internal node
ZPC_HD void reorder_node(int idx, int n, const int *lOffsets, const int *lPars, const int *lLcas, const int *tPars, const int *tRs,
                         const int *tDst, int *auxIndices, int *parents) {
  const int numTrunk = n - 1;
  if (idx < n) parents[lOffsets[idx + 1] - 1] = tDst[lPars[idx]];
  if (idx < numTrunk) {
    const int dst = tDst[idx], r = tRs[idx];
    if (r != numTrunk) {
      const int lca = lLcas[r + 1];
      auxIndices[dst] = lca < numTrunk ? tDst[lca] : lOffsets[r + 1];
    } else {
      auxIndices[dst] = -1;
    }
    parents[dst] = idx != 0 ? tDst[tPars[idx]] : -1;
  }
}

// LBvhView::iter_neighbors (Bvh.hpp:660-689): stack-free traversal along the DFS order — `level` nodes of a left chain are
// consecutive, a miss jumps to the node's escape index.  f(primitive id) for every leaf whose box overlaps bv
// (overlaps(AABB, AABB), geometry/AnalyticLevelSet.h:262-266).
ZPC_HD bool boxes_overlap(const float *a, const float *b) {
  return !(b[0] > a[3] || b[3] < a[0] || b[1] > a[4] || b[4] < a[1] || b[2] > a[5] || b[5] < a[2]);
}
template <class F>
ZPC_HD void iter_neighbors(int numLeaves, const float *bvs, const int *auxIndices, const int *levels, const float *bv, F &&f) {
  if (numLeaves <= 2) {
    for (int i = 0; i < numLeaves; ++i)
      if (boxes_overlap(bvs + 6 * i, bv)) f(i);
    return;
  }
  const int numNodes = 2 * numLeaves - 1;
  int node = 0;
  // node only moves forward in a well-formed tree (escape indices point past the subtree): the bound turns a corrupted array
  // into an early exit instead of an endless loop
  for (int guard = 0; node >= 0 && node < numNodes && guard < numNodes; ++guard) {
    int level = levels[node];
    for (; level; --level, ++node)
      if (!boxes_overlap(bvs + 6 * (size_t)node, bv)) break;
    if (level == 0) {
      if (boxes_overlap(bvs + 6 * (size_t)node, bv)) f(auxIndices[node]);
      node++;
    } else {
      node = auxIndices[node];
    }
  }
}

}  // namespace zpcb

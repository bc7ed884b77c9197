// The sweep phase of the binned P2G (csrc/mpm_binned.cu) and the record layout it reads: host+device so that the channel / operand
// mapping of every variant can be checked on the CPU (tests/hostmath), where the packed operations run one half at a time.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#include "mpm_math.cuh"

namespace zpcs {

// record i of the chunk starts at float4 index rec_at<VAR>(i).  VAR 4 reads three records per LDS.128 (one per lane
// group), typically 8 apart (8 particles per cell): a pad granule every 8 records puts them in different banks.
template <int VAR> ZPC_HD int rec_at(int i) { return VAR >= 4 ? 7 * i + (i >> 3) : 7 * i; }

// lanes = the 27 stencil offsets.  Sweeps the records of sorted positions [lo,hi) (all in one cell) and returns the
// 7 channel sums of this lane's node.
struct LaneCoef {
  float ax, bx, cx, ay, by, cy, az, bz, cz, fx, fy, fz;
};
ZPC_HD void sweep_cell(const float4 *rec, int lo, int hi, const LaneCoef &L, float (&acc)[7]) {
#pragma unroll 1
  for (int p = lo; p < hi; ++p) {
    const float4 *rp = rec + 7 * p;
    const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3], r4 = rp[4], r5 = rp[5], r6 = rp[6];
    const float wx = fmaf(fmaf(L.ax, r0.x, L.bx), r0.x, L.cx), wy = fmaf(fmaf(L.ay, r0.y, L.by), r0.y, L.cy),
                wz = fmaf(fmaf(L.az, r0.z, L.bz), r0.z, L.cz);
    const float W = wx * wy * wz;
    acc[0] = fmaf(W, r0.w, acc[0]);
    // A = r1.xyz ; B row d = (r1.w r2.x r2.y), (r2.z r2.w r3.x), (r3.y r3.z r3.w)
    acc[1] = fmaf(W, fmaf(r2.y, L.fz, fmaf(r2.x, L.fy, fmaf(r1.w, L.fx, r1.x))), acc[1]);
    acc[2] = fmaf(W, fmaf(r3.x, L.fz, fmaf(r2.w, L.fy, fmaf(r2.z, L.fx, r1.y))), acc[2]);
    acc[3] = fmaf(W, fmaf(r3.w, L.fz, fmaf(r3.z, L.fy, fmaf(r3.y, L.fx, r1.z))), acc[3]);
    // a = r4.xyz ; K row d = (r4.w r5.x r5.y), (r5.z r5.w r6.x), (r6.y r6.z r6.w)
    acc[4] = fmaf(W, fmaf(r5.y, L.fz, fmaf(r5.x, L.fy, fmaf(r4.w, L.fx, r4.x))), acc[4]);
    acc[5] = fmaf(W, fmaf(r6.x, L.fz, fmaf(r5.w, L.fy, fmaf(r5.z, L.fx, r4.y))), acc[5]);
    acc[6] = fmaf(W, fmaf(r6.w, L.fz, fmaf(r6.z, L.fy, fmaf(r6.y, L.fx, r4.z))), acc[6]);
  }
}


// ---- v4 sweep: lanes = 3 cells x 9 (ox,oy) node columns -------------------------------------------------------
// A warp takes three non-empty cells at a time.  Lane group gi = lane/9 owns one cell, lane j = lane%9 owns the node
// column (ox,oy) = (j/3, j%3) of that cell's stencil and keeps the sums of its three z-nodes x 7 channels in
// registers: the affine part A + B.o is evaluated once per column (12 FMA) and extended along z with 12 more, the
// x/y weights are shared by the three nodes — 59 FP instructions per (particle, column) for 3 nodes instead of
// 33 per node, and every LDS.128 of a record now serves three different particles.
struct ColCoef {
  float ax, bx, cx, ay, by, cy, fx, fy;
};
ZPC_HD void sweep_cells3(const float4 *rec, int lo, int hi, int nmax, const ColCoef &L,
                                             float (&acc)[7][3]) {
  // lo, hi: record indices relative to the chunk
#pragma unroll 1
  for (int it = 0; it < nmax; ++it) {
    const int p = lo + it;
    if (p < hi) {
      const float4 *rp = rec + rec_at<4>(p);
      const float4 r0 = rp[0];
      const float wx = fmaf(fmaf(L.ax, r0.x, L.bx), r0.x, L.cx), wy = fmaf(fmaf(L.ay, r0.y, L.by), r0.y, L.cy);
      const float wxy = wx * wy;
      const float W0 = wxy * fmaf(fmaf(0.5f, r0.z, -1.5f), r0.z, 1.125f);
      const float W1 = wxy * fmaf(fmaf(-1.0f, r0.z, 2.0f), r0.z, -0.25f);
      const float W2 = wxy * fmaf(fmaf(0.5f, r0.z, -0.5f), r0.z, 0.125f);
      acc[0][0] = fmaf(W0, r0.w, acc[0][0]);
      acc[0][1] = fmaf(W1, r0.w, acc[0][1]);
      acc[0][2] = fmaf(W2, r0.w, acc[0][2]);
#define ZPC_COL3(CH, A0, BX, BY, BZ)                                    \
  {                                                                     \
    const float b0 = fmaf(BY, L.fy, fmaf(BX, L.fx, A0));                \
    acc[CH][0] = fmaf(W0, b0, acc[CH][0]);                              \
    acc[CH][1] = fmaf(W1, b0 + BZ, acc[CH][1]);                         \
    acc[CH][2] = fmaf(W2, fmaf(2.0f, BZ, b0), acc[CH][2]);              \
  }
      // A = r1.xyz ; B row d = (r1.w r2.x r2.y), (r2.z r2.w r3.x), (r3.y r3.z r3.w)
      const float4 r1 = rp[1], r2 = rp[2], r3 = rp[3];
      ZPC_COL3(1, r1.x, r1.w, r2.x, r2.y)
      ZPC_COL3(2, r1.y, r2.z, r2.w, r3.x)
      ZPC_COL3(3, r1.z, r3.y, r3.z, r3.w)
      // a = r4.xyz ; K row d = (r4.w r5.x r5.y), (r5.z r5.w r6.x), (r6.y r6.z r6.w)
      const float4 r4 = rp[4], r5 = rp[5], r6 = rp[6];
      ZPC_COL3(4, r4.x, r4.w, r5.x, r5.y)
      ZPC_COL3(5, r4.y, r5.z, r5.w, r6.x)
      ZPC_COL3(6, r4.z, r6.y, r6.z, r6.w)
#undef ZPC_COL3
    }
  }
}

// ---- v5 sweep: the v4 sweep on packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100: two IEEE-rounded fp32 operations per
// issue slot) -------------------------------------------------------------------------------------------------------------------
// The six vector channels go through the same 7 operations per z-column with the same weights, so channels are paired:
// (1,2), (3,4), (5,6).  The record keeps the v4 size (7 float4) but stores the operands of a pair next to each other —
//   rec[1 + 2q] = (A0_c, A0_c', BX_c, BX_c'),  rec[2 + 2q] = (BY_c, BY_c', BZ_c, BZ_c')   for pair q = (c, c') —
// so that every LDS.128 delivers two aligned register pairs.  Each half of a packed operation is the scalar operation of the v4
// sweep (fma.rn / add.rn per half): the per-lane sums are bit-identical to v4's; 38 issue slots per (particle, column) against 59.
ZPC_HD float2 f2(float a, float b) { return make_float2(a, b); }
// two IEEE-rounded fp32 operations per instruction on the device (FFMA2 / FADD2); the host pass does the two halves one by one
ZPC_HD float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
ZPC_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  return __fadd2_rn(a, b);
#else
  return make_float2(zpcm::rn_add(a.x, b.x), zpcm::rn_add(a.y, b.y));
#endif
}
ZPC_HD void sweep_cells3_packed(const float4 *rec, int lo, int hi, int nmax, const ColCoef &L, float (&accm)[3],
                                                    float2 (&acc)[3][3]) {
  // the lane's eight coefficients are made opaque so that ptxas keeps them in registers instead of rebuilding them from (ox, oy)
  // inside the loop (14 extra instructions per iteration at the 64-register cap)
  float ax = L.ax, bx = L.bx, cx = L.cx, ay = L.ay, by = L.by, cy = L.cy, fx = L.fx, fy = L.fy;
#ifdef __CUDA_ARCH__
  asm volatile("" : "+f"(ax), "+f"(bx), "+f"(cx), "+f"(ay), "+f"(by), "+f"(cy), "+f"(fx), "+f"(fy));
#endif
  const float2 fx2 = f2(fx, fx), fy2 = f2(fy, fy), two2 = f2(2.0f, 2.0f);
#pragma unroll 1
  for (int it = 0; it < nmax; ++it) {
    const int p = lo + it;
    if (p < hi) {
      const float4 *rp = rec + rec_at<5>(p);
      const float4 r0 = rp[0];
      const float2 xy = f2(r0.x, r0.y);
      const float2 w2 = fma2(fma2(f2(ax, ay), xy, f2(bx, by)), xy, f2(cx, cy));
      const float wxy = w2.x * w2.y;   // scalar like v4: packing (wx, wy) made ptxas rebuild the six coefficients inside the loop
      const float W0 = wxy * fmaf(fmaf(0.5f, r0.z, -1.5f), r0.z, 1.125f);
      const float W1 = wxy * fmaf(fmaf(-1.0f, r0.z, 2.0f), r0.z, -0.25f);
      const float W2 = wxy * fmaf(fmaf(0.5f, r0.z, -0.5f), r0.z, 0.125f);
      accm[0] = fmaf(W0, r0.w, accm[0]);
      accm[1] = fmaf(W1, r0.w, accm[1]);
      accm[2] = fmaf(W2, r0.w, accm[2]);
      const float2 W0p = f2(W0, W0), W1p = f2(W1, W1), W2p = f2(W2, W2);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 q0 = rp[1 + 2 * q], q1 = rp[2 + 2 * q];
        const float2 BZ = f2(q1.z, q1.w);
        const float2 b0 = fma2(f2(q1.x, q1.y), fy2, fma2(f2(q0.z, q0.w), fx2, f2(q0.x, q0.y)));
        acc[q][0] = fma2(W0p, b0, acc[q][0]);
        acc[q][1] = fma2(W1p, add2(b0, BZ), acc[q][1]);
        acc[q][2] = fma2(W2p, fma2(two2, BZ, b0), acc[q][2]);
      }
    }
  }
}


// The 28-float record of one particle as the sweep variant VAR reads it: d0 = position in the stencil's frame (three floats) and the
// mass, then for the six vector channels ch the four numbers of  value_ch(o) = A0_ch + BX_ch o_x + BY_ch o_y + BZ_ch o_z :
// channels 1..3 = momentum (A, B), 4..6 = rhs (a, Kd).
template <int VAR>
ZPC_HD void write_record(float4 *dst, const float (&d0)[3], float mass, const float (&A)[3], const float (&a)[3], const float (&B)[9],
                         const float (&Kd)[9]) {
  dst[0] = make_float4(d0[0], d0[1], d0[2], mass);
  if constexpr (VAR == 5) {  // channel pairs (1,2) (3,4) (5,6): (A0, A0', BX, BX'), (BY, BY', BZ, BZ') — see sweep_cells3_packed
    dst[1] = make_float4(A[0], A[1], B[0], B[3]);
    dst[2] = make_float4(B[1], B[4], B[2], B[5]);
    dst[3] = make_float4(A[2], a[0], B[6], Kd[0]);
    dst[4] = make_float4(B[7], Kd[1], B[8], Kd[2]);
    dst[5] = make_float4(a[1], a[2], Kd[3], Kd[6]);
    dst[6] = make_float4(Kd[4], Kd[7], Kd[5], Kd[8]);
  } else {
    dst[1] = make_float4(A[0], A[1], A[2], B[0]);
    dst[2] = make_float4(B[1], B[2], B[3], B[4]);
    dst[3] = make_float4(B[5], B[6], B[7], B[8]);
    dst[4] = make_float4(a[0], a[1], a[2], Kd[0]);
    dst[5] = make_float4(Kd[1], Kd[2], Kd[3], Kd[4]);
    dst[6] = make_float4(Kd[5], Kd[6], Kd[7], Kd[8]);
  }
}


// ---- plane sweep (round 2): lanes = 10 cells x 3 x-planes ----------------------------------------------------------------------
// The column sweep above is bound by the shared-memory -> register path, not by issue slots: every one of the nine lanes that serve a
// particle pulls the whole 112-byte record (7 LDS.128 = 28 data-pipe cycles per three particles; ncu: 13.3 wavefronts per particle,
// data pipe 69 % busy).  Here a lane owns one x-plane of a cell's stencil — nine nodes x 7 channels = 63 sums in registers — so a
// record is pulled by three lanes instead of nine (28 cycles per TEN particles), the y/z weights and the affine part are shared by
// nine nodes (143 FP instructions per (particle, plane) = 429 per particle against 9 x 59 = 531), and a warp takes ten cells per unit.
// Record layout: rec[0] = (d0x, d0y, d0z, m); rec[1 + c] = (A0, BX, BY, BZ) of vector channel c (0..2 momentum, 3..5 rhs):
// value_c(i, j, k) = A0 + BX i + BY j + BZ k.
struct PlaneCoef {
  float ax, bx, cx, fi;   // x weight of this lane's plane i as a polynomial in d0x, and (float)i
};
ZPC_HD int prec_at(int i) { return 7 * i + (i >> 3); }   // same pad granule as rec_at<4>: records of neighbouring cells (8 apart) land in different banks
ZPC_HD void write_plane_record(float4 *dst, const float (&d0)[3], float mass, const float (&A)[3], const float (&a)[3], const float (&B)[9],
                               const float (&Kd)[9]) {
  dst[0] = make_float4(d0[0], d0[1], d0[2], mass);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dst[1 + c] = make_float4(A[c], B[3 * c], B[3 * c + 1], B[3 * c + 2]);
    dst[4 + c] = make_float4(a[c], Kd[3 * c], Kd[3 * c + 1], Kd[3 * c + 2]);
  }
}
// acc[ch][j][k]: channel ch of node (i, j, k) of the lane's plane
ZPC_HD void sweep_plane(const float4 *rec, int lo, int hi, int nmax, const PlaneCoef &L, float (&acc)[7][3][3]) {
#pragma unroll 1
  for (int it = 0; it < nmax; ++it) {
    const int p = lo + it;
    if (p < hi) {
      const float4 *rp = rec + prec_at(p);
      const float4 r0 = rp[0];
      const float wx = fmaf(fmaf(L.ax, r0.x, L.bx), r0.x, L.cx);
      float wxy[3], wz[3], W[3][3];
      wxy[0] = wx * fmaf(fmaf(0.5f, r0.y, -1.5f), r0.y, 1.125f);
      wxy[1] = wx * fmaf(fmaf(-1.0f, r0.y, 2.0f), r0.y, -0.25f);
      wxy[2] = wx * fmaf(fmaf(0.5f, r0.y, -0.5f), r0.y, 0.125f);
      wz[0] = fmaf(fmaf(0.5f, r0.z, -1.5f), r0.z, 1.125f);
      wz[1] = fmaf(fmaf(-1.0f, r0.z, 2.0f), r0.z, -0.25f);
      wz[2] = fmaf(fmaf(0.5f, r0.z, -0.5f), r0.z, 0.125f);
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          W[j][k] = wxy[j] * wz[k];
          acc[0][j][k] = fmaf(W[j][k], r0.w, acc[0][j][k]);
        }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float4 q = rp[1 + c];
        const float base = fmaf(q.y, L.fi, q.x);
        const float bj[3] = {base, base + q.z, fmaf(2.0f, q.z, base)};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          acc[1 + c][j][0] = fmaf(W[j][0], bj[j], acc[1 + c][j][0]);
          acc[1 + c][j][1] = fmaf(W[j][1], bj[j] + q.w, acc[1 + c][j][1]);
          acc[1 + c][j][2] = fmaf(W[j][2], fmaf(2.0f, q.w, bj[j]), acc[1 + c][j][2]);
        }
      }
    }
  }
}

}  // namespace zpcs

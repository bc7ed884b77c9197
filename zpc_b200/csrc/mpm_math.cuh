// Per-particle math of the MPM transfer path (registers only).
//   svd3             : McAdams et al. 3x3 SVD, 4 cyclic Jacobi sweeps + Givens QR — the algorithm of
//                      reference math/matrix/SVD.hpp:15-1026 (constants :27-32), with ::rsqrtf as the
//                      reference's device build uses (ZpcMathUtils.hpp:812-816).
//   stress_fcr       : fixed-corotated  P F^T vol  (physics/ConstitutiveModel_Vol_dP.hpp:10-47).
//   bspline weights  : quadratic B-spline (math/curve/InterpolationKernel.hpp:91-128) around
//                      base_node<1> (:46-55) as used by LocalArena (simulation/Utils.hpp:51-70).
// Matrices are column-major 9-vectors M[3*col+row], exactly the reference's vec9 convention.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "../../include/zpcb200.h"

// The per-particle math is __host__ __device__: tests/hostmath compiles it for the CPU and checks it against the
// oracle without a GPU (transcription errors show up there; the device build is the one that ships).
#define ZPC_HD __host__ __device__ __forceinline__

namespace zpcm {

ZPC_HD float bits_to_float(unsigned u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// float add to the grid: a RED on the device; the host build (tests/hostmath) runs one particle at a time
ZPC_HD void grid_add(float *p, float v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
}
ZPC_HD float grid_load(const float *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

template <int AXIS>
ZPC_HD void jacobi_step(float &pp, float &qq, float &off, float &rr, float &a, float &b,
                                            float (&q)[4]) {
  const float tiny = 1.e-20f, gamma = 5.8284273147583007813f;
  const float sin_pi8 = bits_to_float(1053028117u), cos_pi8 = bits_to_float(1064076127u);
  float sh = off * 0.5f;
  float t5 = pp - qq;
  const bool big = sh * sh >= tiny;
  sh = big ? sh : 0.0f;
  float ch = big ? t5 : 1.0f;
  float t1 = sh * sh, t2 = ch * ch;
  const float r = rsqrtf(t1 + t2);
  sh = r * sh;
  ch = r * ch;
  if (t2 <= gamma * t1) { sh = sin_pi8; ch = cos_pi8; }
  t1 = sh * sh;
  t2 = ch * ch;
  const float c = t2 - t1;
  float s = ch * sh;
  s = s + s;
  // conjugate the symmetric matrix (the (sh^2+ch^2) factors keep it consistent when the pair is unnormalised)
  const float n2 = t1 + t2;
  rr = rr * n2 * n2;
  a = a * n2;
  b = b * n2;
  const float ta = s * a, tb = s * b;
  a = c * a + tb;
  b = c * b - ta;
  const float s2 = s * s, c2 = c * c, cs = c * s;
  const float npp = pp * c2 + qq * s2, nqq = qq * c2 + pp * s2;
  const float twice = (off + off) * cs;
  off = off * (c2 - s2) - t5 * cs;
  pp = npp + twice;
  qq = nqq - twice;
  // accumulate the rotation as a quaternion (s, x, y, z)
  const float tx = sh * q[1], ty = sh * q[2], tz = sh * q[3];
  const float t[3] = {tx, ty, tz};
  sh = sh * q[0];
  q[0] = ch * q[0];
  q[1] = ch * q[1];
  q[2] = ch * q[2];
  q[3] = ch * q[3];
  constexpr int B = (AXIS + 1) % 3, C = (AXIS + 2) % 3;
  q[1 + AXIS] += sh;
  q[0] -= t[AXIS];
  q[1 + B] += t[C];
  q[1 + C] -= t[B];
}

ZPC_HD float rsqrt_refined(float x) {  // one Newton step, SVD.hpp:386-392
  const float r = rsqrtf(x);
  const float h = r * 0.5f;
  return (r + h) - x * (r * (r * h));
}

ZPC_HD void qr_givens(float piv, float low, float &c, float &s) {
  const float small = 1.e-12f;
  float sh = (low * low >= small) ? low : 0.0f;
  float ch = fmaxf(fmaxf(-piv, piv), small);
  const bool nonneg = piv >= 0.0f;
  float n2 = ch * ch + sh * sh;
  ch = ch + rsqrt_refined(n2) * n2;
  if (!nonneg) { const float t = ch; ch = sh; sh = t; }
  n2 = ch * ch + sh * sh;
  const float r = rsqrt_refined(n2);
  ch *= r;
  sh *= r;
  c = ch * ch - sh * sh;
  s = sh * ch;
  s = s + s;
}
ZPC_HD void rot_pair(float &x, float &y, float c, float s) {
  const float t1 = s * x, t2 = s * y;
  x = c * x + t2;
  y = c * y - t1;
}

// F, U, V column-major; S the three singular values (signed so that U,V are rotations)
ZPC_HD void svd3(const float (&F)[9], float (&U)[9], float (&S)[3], float (&V)[9]) {
  // row-major local copy a[r][c] = F[3c+r]
  float a00 = F[0], a01 = F[3], a02 = F[6], a10 = F[1], a11 = F[4], a12 = F[7], a20 = F[2], a21 = F[5], a22 = F[8];
  float s11 = a00 * a00 + a10 * a10 + a20 * a20;
  float s21 = a01 * a00 + a11 * a10 + a21 * a20;
  float s31 = a02 * a00 + a12 * a10 + a22 * a20;
  float s22 = a01 * a01 + a11 * a11 + a21 * a21;
  float s32 = a02 * a01 + a12 * a11 + a22 * a21;
  float s33 = a02 * a02 + a12 * a12 + a22 * a22;
  float q[4] = {1.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int sweep = 0; sweep < 4; ++sweep) {
    jacobi_step<2>(s11, s22, s21, s33, s31, s32, q);
    jacobi_step<0>(s22, s33, s32, s11, s21, s31, q);
    jacobi_step<1>(s33, s11, s31, s22, s32, s21, q);
  }
  {
    const float r = rsqrt_refined(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= r; q[1] *= r; q[2] *= r; q[3] *= r;
  }
  float v00, v01, v02, v10, v11, v12, v20, v21, v22;
  {
    const float xx = q[1] * q[1], yy = q[2] * q[2], zz = q[3] * q[3], ww = q[0] * q[0];
    v00 = ww + xx - yy - zz;
    v11 = ww - xx + yy - zz;
    v22 = ww - xx - yy + zz;
    const float x2 = q[1] + q[1], y2 = q[2] + q[2], z2 = q[3] + q[3];
    const float wx = q[0] * x2, wy = q[0] * y2, wz = q[0] * z2;
    const float xy = q[2] * x2, yz = q[3] * y2, zx = q[1] * z2;
    v01 = xy - wz; v12 = yz - wx; v20 = zx - wy;
    v10 = xy + wz; v21 = yz + wx; v02 = zx + wy;
  }
  // B = A V
  float b00 = a00 * v00 + a01 * v10 + a02 * v20, b01 = a00 * v01 + a01 * v11 + a02 * v21, b02 = a00 * v02 + a01 * v12 + a02 * v22;
  float b10 = a10 * v00 + a11 * v10 + a12 * v20, b11 = a10 * v01 + a11 * v11 + a12 * v21, b12 = a10 * v02 + a11 * v12 + a12 * v22;
  float b20 = a20 * v00 + a21 * v10 + a22 * v20, b21 = a20 * v01 + a21 * v11 + a22 * v21, b22 = a20 * v02 + a21 * v12 + a22 * v22;
  float r0 = b00 * b00 + b10 * b10 + b20 * b20, r1 = b01 * b01 + b11 * b11 + b21 * b21, r2 = b02 * b02 + b12 * b12 + b22 * b22;
#define ZPC_SWAPNEG(cond, xa, ya, za, xb, yb, zb, va, vb, vc, wa, wb, wc, ra, rb, NEG_FIRST)          \
  if (cond) {                                                                                           \
    float t;                                                                                            \
    t = xa; xa = xb; xb = t; t = ya; ya = yb; yb = t; t = za; za = zb; zb = t;                          \
    t = va; va = wa; wa = t; t = vb; vb = wb; wb = t; t = vc; vc = wc; wc = t;                          \
    t = ra; ra = rb; rb = t;                                                                            \
    if (NEG_FIRST) { xa = -xa; ya = -ya; za = -za; va = -va; vb = -vb; vc = -vc; }                      \
    else { xb = -xb; yb = -yb; zb = -zb; wa = -wa; wb = -wb; wc = -wc; }                                \
  }
  // swap (1,2) negate col 2; swap (1,3) negate col 1; swap (2,3) negate col 3  (SVD.hpp:516-676)
  ZPC_SWAPNEG(r0 < r1, b00, b10, b20, b01, b11, b21, v00, v10, v20, v01, v11, v21, r0, r1, false)
  ZPC_SWAPNEG(r0 < r2, b00, b10, b20, b02, b12, b22, v00, v10, v20, v02, v12, v22, r0, r2, true)
  ZPC_SWAPNEG(r1 < r2, b01, b11, b21, b02, b12, b22, v01, v11, v21, v02, v12, v22, r1, r2, false)
#undef ZPC_SWAPNEG
  float u00 = 1.f, u01 = 0.f, u02 = 0.f, u10 = 0.f, u11 = 1.f, u12 = 0.f, u20 = 0.f, u21 = 0.f, u22 = 1.f;
  float c, s;
  qr_givens(b00, b10, c, s);  // rows (0,1)
  rot_pair(b00, b10, c, s); rot_pair(b01, b11, c, s); rot_pair(b02, b12, c, s);
  rot_pair(u00, u01, c, s); rot_pair(u10, u11, c, s); rot_pair(u20, u21, c, s);
  qr_givens(b00, b20, c, s);  // rows (0,2)
  rot_pair(b00, b20, c, s); rot_pair(b01, b21, c, s); rot_pair(b02, b22, c, s);
  rot_pair(u00, u02, c, s); rot_pair(u10, u12, c, s); rot_pair(u20, u22, c, s);
  qr_givens(b11, b21, c, s);  // rows (1,2)
  rot_pair(b10, b20, c, s); rot_pair(b11, b21, c, s); rot_pair(b12, b22, c, s);
  rot_pair(u01, u02, c, s); rot_pair(u11, u12, c, s); rot_pair(u21, u22, c, s);
  U[0] = u00; U[1] = u10; U[2] = u20; U[3] = u01; U[4] = u11; U[5] = u21; U[6] = u02; U[7] = u12; U[8] = u22;
  V[0] = v00; V[1] = v10; V[2] = v20; V[3] = v01; V[4] = v11; V[5] = v21; V[6] = v02; V[7] = v12; V[8] = v22;
  S[0] = b00; S[1] = b11; S[2] = b22;
}

// lame_parameters<float> (physics/ConstitutiveModel.hpp:34-38): mu goes through double because of the
// 0.5 literal, lambda is all-float.  Evaluated once on the host and passed to the kernels.
inline void lame_host(float E, float nu, float &mu, float &lam) {
  mu = (float)(0.5 * E / (1 + nu));
  lam = E * nu / ((1 + nu) * (1 - 2 * nu));
}

// PF = P(F) F^T * volume for the fixed-corotated model
ZPC_HD void stress_fcr(float volume, float mu, float lam, const float (&F)[9], float (&PF)[9]) {
  float U[9], S[3], V[9];
  svd3(F, U, S, V);
  const float J = S[0] * S[1] * S[2];
  const float smu = 2.f * mu, sl = lam * (J - 1.f);
  const float Ph[3] = {smu * (S[0] - 1.f) + sl * (S[1] * S[2]), smu * (S[1] - 1.f) + sl * (S[0] * S[2]),
                       smu * (S[2] - 1.f) + sl * (S[0] * S[1])};
  float P[9];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      P[3 * c + r] = Ph[0] * U[r] * V[c] + Ph[1] * U[3 + r] * V[3 + c] + Ph[2] * U[6 + r] * V[6 + c];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) PF[3 * c + r] = (P[r] * F[c] + P[3 + r] * F[3 + c] + P[6 + r] * F[6 + c]) * volume;
}

// x / d in three dependent operations and no branch, given inv = 1.0f / d (IEEE, once per kernel): q = RN(x inv), r = x - q d (exact
// in one fma), RN(q + r inv) — the correction step of a Newton division (Markstein).  Bit-identical to the IEEE quotient for operands
// in the normal range: 5.2e8 random positions x 13 cell sizes without a mismatch on the host, and tests/test_hostmath_transfer.py
// replays that against `/`; nvcc's own x / d is the same arithmetic plus a range check behind a branch, which splits the record
// code into basic blocks the scheduler cannot interleave.
ZPC_HD float div_exact(float x, float d, float inv) {
  const float q = x * inv;
  return fmaf(fmaf(-q, d, x), inv, q);
}

// ---- the same stress with the arithmetic of the binned fast path (round 2) -----------------------------------------------------
// stress_fcr above follows the reference operation by operation (~1 050 instructions per particle, 36 of the 87 warp-instructions
// the binned P2G spent per particle).  The fixed-corotated model only needs  P F^T = U Phat V^T F^T = sum_k Phat_k u_k (F v_k)^T,
// so everything after the Jacobi sweeps is restated in its cheapest algebraically equal form:
//   * the twelve Jacobi rotations keep the reference's rule for the angle (approximate Givens, the pi/8 fallback, the 1e-20 guard:
//     SVD.hpp:27-32, 63-160 — the eigenvector basis the reference stops at after four sweeps is part of its result), but the
//     rotation is applied in normalised form: the (sh^2 + ch^2) factors that keep an UNnormalised pair consistent are 1 +- 2 ulp
//     here, and the third diagonal entry is not touched at all;
//   * B = F V, columns ordered by norm with the reference's sign rule (SVD.hpp:516-676);
//   * QR of B by Gram-Schmidt instead of three Givens rotations applied to B and to an identity matrix: the factorisation is unique
//     (diagonal of R >= 0 for the first two columns, U a rotation), so U and the singular values agree to rounding;
//   * P F^T as three outer products u_k (Phat_k b_k)^T instead of P = U Phat V^T followed by P F^T.
// Every difference to stress_fcr is a rounding-level perturbation of an intermediate (no step is skipped), i.e. of the same kind as
// the difference between the reference's own host and device builds; tests/test_oracle_plastic.py holds it to that distance.
ZPC_HD float rsqrt_fast(float x) {  // x is never subnormal where this is used (>= 1e-20 or a sum with 1)
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
template <int AXIS>
ZPC_HD void jacobi_lean(float &pp, float &qq, float &off, float &a, float &b, float (&q)[4]) {
  const float tiny = 1.e-20f, gamma = 5.8284273147583007813f;
  const float sin_pi8 = bits_to_float(1053028117u), cos_pi8 = bits_to_float(1064076127u);
  float sh = off * 0.5f;
  const float t5 = pp - qq;
  const bool big = sh * sh >= tiny;
  sh = big ? sh : 0.0f;
  float ch = big ? t5 : 1.0f;
  float t1 = sh * sh, t2 = ch * ch;
  const float r = rsqrt_fast(t1 + t2);
  sh = r * sh;
  ch = r * ch;
  if (t2 <= gamma * t1) { sh = sin_pi8; ch = cos_pi8; }
  t1 = sh * sh;
  t2 = ch * ch;
  const float c = t2 - t1;
  float s = ch * sh;
  s = s + s;
  const float sa = s * a, sb = s * b;
  a = fmaf(c, a, sb);
  b = fmaf(c, b, -sa);
  const float s2 = s * s, c2 = c * c, cs = c * s;
  const float twice = (off + off) * cs;
  const float npp = fmaf(pp, c2, fmaf(qq, s2, twice)), nqq = fmaf(qq, c2, fmaf(pp, s2, -twice));
  off = fmaf(off, c2 - s2, -(t5 * cs));
  pp = npp;
  qq = nqq;
  const float tx = sh * q[1], ty = sh * q[2], tz = sh * q[3];
  const float t[3] = {tx, ty, tz};
  sh = sh * q[0];
  q[0] = ch * q[0];
  q[1] = ch * q[1];
  q[2] = ch * q[2];
  q[3] = ch * q[3];
  constexpr int B = (AXIS + 1) % 3, C = (AXIS + 2) % 3;
  q[1 + AXIS] += sh;
  q[0] -= t[AXIS];
  q[1 + B] += t[C];
  q[1 + C] -= t[B];
}
// scale = the factor the caller wants on P F^T (volume, or volume * -dt * D_inv for the P2G record): folded into Phat
// EARLY (binned P2G, sweep variant 8): the reference always runs four cyclic sweeps; once every off-diagonal entry of every lane of
// the warp is below 2^-22 of the trace the remaining rotations have angles at rounding level — they move P F^T by less than the
// distance between two builds of the reference's own arithmetic (tests/test_oracle_sensitivity.py) — and the warp skips them
// together (a warp-uniform branch: no divergence).  The host pass (tests/hostmath) applies the same rule per particle.
template <bool EARLY = false>
ZPC_HD void stress_fcr_lean(float scale, float mu, float lam, const float (&F)[9], float (&PF)[9]) {
  const float a00 = F[0], a01 = F[3], a02 = F[6], a10 = F[1], a11 = F[4], a12 = F[7], a20 = F[2], a21 = F[5], a22 = F[8];
  float s11 = a00 * a00 + a10 * a10 + a20 * a20;
  float s21 = a01 * a00 + a11 * a10 + a21 * a20;
  float s31 = a02 * a00 + a12 * a10 + a22 * a20;
  float s22 = a01 * a01 + a11 * a11 + a21 * a21;
  float s32 = a02 * a01 + a12 * a11 + a22 * a21;
  float s33 = a02 * a02 + a12 * a12 + a22 * a22;
  float q[4] = {1.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int sweep = 0; sweep < 4; ++sweep) {
    if constexpr (EARLY) {
      if (sweep >= 2) {
        const float offmax = fmaxf(fmaxf(fabsf(s21), fabsf(s31)), fabsf(s32));
        const bool more = offmax > 2.384185791015625e-7f * (s11 + s22 + s33);
#ifdef __CUDA_ARCH__
        if (!__any_sync(__activemask(), more)) break;   // the caller's lanes (a partial last warp included)
#else
        if (!more) break;
#endif
      }
    }
    jacobi_lean<2>(s11, s22, s21, s31, s32, q);
    jacobi_lean<0>(s22, s33, s32, s21, s31, q);
    jacobi_lean<1>(s33, s11, s31, s32, s21, q);
  }
  {
    const float r = rsqrt_refined(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= r; q[1] *= r; q[2] *= r; q[3] *= r;
  }
  float v00, v01, v02, v10, v11, v12, v20, v21, v22;
  {
    const float xx = q[1] * q[1], yy = q[2] * q[2], zz = q[3] * q[3], ww = q[0] * q[0];
    v00 = ww + xx - yy - zz;
    v11 = ww - xx + yy - zz;
    v22 = ww - xx - yy + zz;
    const float x2 = q[1] + q[1], y2 = q[2] + q[2], z2 = q[3] + q[3];
    const float wx = q[0] * x2, wy = q[0] * y2, wz = q[0] * z2;
    const float xy = q[2] * x2, yz = q[3] * y2, zx = q[1] * z2;
    v01 = xy - wz; v12 = yz - wx; v20 = zx - wy;
    v10 = xy + wz; v21 = yz + wx; v02 = zx + wy;
  }
  // B = F V: column k = F v_k
  float b0[3] = {a00 * v00 + a01 * v10 + a02 * v20, a10 * v00 + a11 * v10 + a12 * v20, a20 * v00 + a21 * v10 + a22 * v20};
  float b1[3] = {a00 * v01 + a01 * v11 + a02 * v21, a10 * v01 + a11 * v11 + a12 * v21, a20 * v01 + a21 * v11 + a22 * v21};
  float b2[3] = {a00 * v02 + a01 * v12 + a02 * v22, a10 * v02 + a11 * v12 + a12 * v22, a20 * v02 + a21 * v12 + a22 * v22};
  float r0 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2], r1 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2],
        r2 = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
  // swap (1,2) negate col 2; swap (1,3) negate col 1; swap (2,3) negate col 3  (SVD.hpp:516-676): only B is needed afterwards
#define ZPC_SWAPB(cond, ba, bb, ra, rb, NEG_FIRST)                                                     \
  {                                                                                                     \
    const bool sw = (cond);                                                                             \
    _Pragma("unroll") for (int d = 0; d < 3; ++d) {                                                     \
      const float xa = ba[d], xb = bb[d];                                                               \
      ba[d] = sw ? (NEG_FIRST ? -xb : xb) : xa;                                                         \
      bb[d] = sw ? (NEG_FIRST ? xa : -xa) : xb;                                                         \
    }                                                                                                   \
    const float ta = ra;                                                                                \
    ra = sw ? rb : ra;                                                                                  \
    rb = sw ? ta : rb;                                                                                  \
  }
  ZPC_SWAPB(r0 < r1, b0, b1, r0, r1, false)
  ZPC_SWAPB(r0 < r2, b0, b2, r0, r2, true)
  ZPC_SWAPB(r1 < r2, b1, b2, r1, r2, false)
#undef ZPC_SWAPB
  // Gram-Schmidt QR: u0 = b0 / |b0| ; u1 = (b1 - (u0.b1) u0) / |.| ; u2 = u0 x u1 ; sigma = (|b0|, |b1'|, u2.b2)
  // floor of the squared column norms: rsqrt_refined's Newton step forms r * (r * h) ~ 0.5 x^-3/2, which overflows fp32 below
  // x ~ 1e-26 (an exactly rank-deficient or zero F then gave NaN where the reference's guarded Givens steps stay finite)
  const float tiny2 = 1.e-24f;
  const float i0 = rsqrt_refined(fmaxf(r0, tiny2));
  const float u0[3] = {b0[0] * i0, b0[1] * i0, b0[2] * i0};
  const float sg0 = r0 * i0;
  const float d01 = u0[0] * b1[0] + u0[1] * b1[1] + u0[2] * b1[2];
  const float c1[3] = {fmaf(-d01, u0[0], b1[0]), fmaf(-d01, u0[1], b1[1]), fmaf(-d01, u0[2], b1[2])};
  const float n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  const float i1 = rsqrt_refined(fmaxf(n1, tiny2));
  const float u1[3] = {c1[0] * i1, c1[1] * i1, c1[2] * i1};
  const float sg1 = n1 * i1;
  const float u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
  const float sg2 = u2[0] * b2[0] + u2[1] * b2[1] + u2[2] * b2[2];
  const float J = sg0 * sg1 * sg2;
  const float smu = 2.f * mu, sl = lam * (J - 1.f);
  const float P0 = (smu * (sg0 - 1.f) + sl * (sg1 * sg2)) * scale, P1 = (smu * (sg1 - 1.f) + sl * (sg0 * sg2)) * scale,
              P2 = (smu * (sg2 - 1.f) + sl * (sg0 * sg1)) * scale;
  const float w0[3] = {P0 * u0[0], P0 * u0[1], P0 * u0[2]}, w1[3] = {P1 * u1[0], P1 * u1[1], P1 * u1[2]},
              w2[3] = {P2 * u2[0], P2 * u2[1], P2 * u2[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) PF[3 * c + r] = fmaf(w0[r], b0[c], fmaf(w1[r], b1[c], w2[r] * b2[c]));
}

// math::sqrtNewtonRaphson<float> (math/MathUtils.h:239-251): Newton iteration from 1 until the step is below
// max(n * 1e-6, 128 eps).  Restated loop for loop: its result is only ~1e-6 accurate and the yield test depends on it.
ZPC_HD float sqrt_newton_raphson(float n) {
  const float eps = 128.f * 1.1920928955078125e-7f;
  if (n < -eps) return bits_to_float(0x7fc00000u);
  if (n < eps) return 0.f;
  float xn = 1.f;
  float xnp1 = 0.5f * (xn + n / xn);
  const float tol = fmaxf(n * 1e-6f, eps);
  for (; fabsf(xnp1 - xn) > tol; xnp1 = 0.5f * (xn + n / xn)) xn = xnp1;
  return xnp1;
}

// compute_stress_vonmisesfixedcorotated (physics/ConstitutiveModel_Vol_dP.hpp:49-110): fixed-corotated trial stress in
// principal space, radial return onto the von Mises cylinder, projected singular values, then the fixed-corotated
// P F^T with the projected F (the projection is not written back to the particle: P2G.hpp:85-91 works on a copy)
ZPC_HD void stress_vonmises(float volume, float mu, float lam, float yield_stress, const float (&Fin)[9],
                                                float (&PF)[9]) {
  float F[9], U[9], S[3], V[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  svd3(F, U, S, V);
  float Sc[3], tau[3], st[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) Sc[d] = 1e-4f > S[d] ? 1e-4f : S[d];
  float J = Sc[0] * Sc[1] * Sc[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) tau[d] = 2 * mu * (Sc[d] - 1) * Sc[d] + lam * (J - 1) * J;
  const float trace_tau = (tau[0] + tau[1]) + tau[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) st[d] = tau[d] - (trace_tau / 3.f);
  const float s_norm = sqrt_newton_raphson((st[0] * st[0] + st[1] * st[1]) + st[2] * st[2]);
  const float scaled_tauy = sqrt_newton_raphson(2.f / (6.f - 3)) * yield_stress;
  if (s_norm - scaled_tauy > 0) {
    const float alpha = scaled_tauy / s_norm;
    J = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float tau_new = alpha * st[d] + (trace_tau / 3.f);
      const float b2m4ac = mu * mu - 2 * mu * (lam * (J - 1) * J - tau_new);
      const float sq = b2m4ac < 0 ? 0.f : sqrt_newton_raphson(b2m4ac);
      S[d] = (mu + sq) / (2 * mu);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) F[3 * c + r] = U[r] * S[0] * V[c] + U[3 + r] * S[1] * V[3 + c] + U[6 + r] * S[2] * V[6 + c];
  }
  J = S[0] * S[1] * S[2];
  const float smu = 2.f * mu, sl = lam * (J - 1.f);
  const float Ph[3] = {smu * (S[0] - 1.f) + sl * (S[1] * S[2]), smu * (S[1] - 1.f) + sl * (S[0] * S[2]),
                       smu * (S[2] - 1.f) + sl * (S[0] * S[1])};
  float P[9];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      P[3 * c + r] = Ph[0] * U[r] * V[c] + Ph[1] * U[3 + r] * V[3 + c] + Ph[2] * U[6 + r] * V[6 + c];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) PF[3 * c + r] = (P[r] * F[c] + P[3 + r] * F[3 + c] + P[6 + r] * F[6 + c]) * volume;
}

// EquationOfStateConfig branch of P2GTransfer / G2P2GTransfer (P2G.hpp:66-83): weakly compressible fluid, pressure from J with the
// exponent fixed to 7 ("from Bow"), viscous part from the symmetrised C; contrib before the -dt * D_inv scaling
ZPC_HD void eos_contrib(const float (&C)[9], float J, float volume, float bulk, float viscosity, float (&contrib)[9]) {
  const float vol = volume * J;
  const float J2 = J * J, J4 = J2 * J2;
  const float pressure = bulk * (1.f / (J * J2 * J4) - 1.f);
  contrib[0] = ((C[0] + C[0]) * viscosity - pressure) * vol;
  contrib[1] = (C[1] + C[3]) * viscosity * vol;
  contrib[2] = (C[2] + C[6]) * viscosity * vol;
  contrib[3] = (C[3] + C[1]) * viscosity * vol;
  contrib[4] = ((C[4] + C[4]) * viscosity - pressure) * vol;
  contrib[5] = (C[5] + C[7]) * viscosity * vol;
  contrib[6] = (C[6] + C[2]) * viscosity * vol;
  contrib[7] = (C[7] + C[5]) * viscosity * vol;
  contrib[8] = ((C[8] + C[8]) * viscosity - pressure) * vol;
}

// parameters of the plastic models as the kernels take them: Drucker-Prager {cohesion, beta, yieldSurface, -}, flag = volumeCorrection;
// NACC {bulk, xi, beta, Msqr}, flag = hardeningOn
struct PlasticPrm {
  float a, b, c, d;
  int flag;
};

// matmul_mat_diag_matT_3D (math/matrix/MatrixUtils.h:26-47): out = A diag(d) B^T, column-major
ZPC_HD void mat_diag_matT(float (&out)[9], const float (&a)[9], const float (&d)[3], const float (&b)[9]) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) out[3 * c + r] = a[r] * d[0] * b[c] + a[3 + r] * d[1] * b[3 + c] + a[6 + r] * d[2] * b[6 + c];
}

// NACCConfig::bulk() / Msqr() (physics/ConstitutiveModel.hpp:756-776); fa goes to sin as is (radians), like the reference.
// Evaluated once on the host and passed to the kernel.
inline float nacc_bulk_host(float E, float nu) { return 2.f / 3.f * (E / (2 * (1 + nu))) + (E * nu / ((1 + nu) * (1 - 2 * nu))); }
inline float nacc_msqr_host(float fa, int dim) {
  const float sin_phi = sinf(fa);
  const float mohr = sqrtf(2.f / 3.f) * 2.f * sin_phi / (3.f - sin_phi);
  const float M = mohr * dim / sqrtf(2.f / (6.f - dim));
  return M * M;
}

// compute_stress_sand (physics/ConstitutiveModel_Vol_dP.hpp:242-326): Drucker-Prager return mapping in Hencky strain.
// The projected F stays local (P2G.hpp:85 works on a copy); logJp is read and written back (P2G.hpp:93,101).
// math::sqrt (math/MathUtils.h:288-323, a software root within 1 ulp) is sqrtf here.
ZPC_HD void stress_sand(float volume, float mu, float lam, float cohesion, float beta, float yieldSurface, bool volCorrection,
                        float &logJp, const float (&Fin)[9], float (&PF)[9]) {
  float F[9], U[9], S[3], V[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  svd3(F, U, S, V);
  const float scaled_mu = 2.f * mu;
  float epsilon[3], New_S[3] = {0.f, 0.f, 0.f}, New_F[9], epsilon_hat[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float abs_S = fabsf(S[i]);
    abs_S = (double)abs_S > 1e-4 ? abs_S : (float)1e-4;
    epsilon[i] = logf(abs_S) - cohesion;
  }
  const float sum_epsilon = epsilon[0] + epsilon[1] + epsilon[2];
  const float trace_epsilon = sum_epsilon + logJp;
#pragma unroll
  for (int i = 0; i < 3; ++i) epsilon_hat[i] = epsilon[i] - (trace_epsilon / 3.f);
  const float epsilon_hat_norm = sqrtf(epsilon_hat[0] * epsilon_hat[0] + epsilon_hat[1] * epsilon_hat[1] + epsilon_hat[2] * epsilon_hat[2]);
  if (trace_epsilon >= 0.f) {  // case II: the cone tip
    New_S[0] = New_S[1] = New_S[2] = expf(cohesion);
    mat_diag_matT(New_F, U, New_S, V);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (volCorrection) logJp = beta * sum_epsilon + logJp;
  } else if (mu != 0.f) {
    logJp = 0.f;
    const float delta_gamma = epsilon_hat_norm + (3.f * lam + scaled_mu) / scaled_mu * trace_epsilon * yieldSurface;
    float H[3];
    if (delta_gamma <= 0.f) {  // case I: inside the cone
#pragma unroll
      for (int i = 0; i < 3; ++i) H[i] = epsilon[i] + cohesion;
    } else {  // case III: onto the cone surface
#pragma unroll
      for (int i = 0; i < 3; ++i) H[i] = epsilon[i] - (delta_gamma / epsilon_hat_norm) * epsilon_hat[i] + cohesion;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) New_S[i] = expf(H[i]);
    mat_diag_matT(New_F, U, New_S, V);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
  }
  const float New_S_log[3] = {logf(New_S[0]), logf(New_S[1]), logf(New_S[2])};
  const float trace_log_S = New_S_log[0] + New_S_log[1] + New_S_log[2];
  float P_hat[3], P[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) P_hat[i] = (scaled_mu * New_S_log[i] + lam * trace_log_S) / New_S[i];
  mat_diag_matT(P, U, P_hat, V);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) PF[3 * c + r] = (P[r] * F[c] + P[3 + r] * F[3 + c] + P[6 + r] * F[6 + c]) * volume;
}

// compute_stress_nacc (physics/ConstitutiveModel_Vol_dP.hpp:116-240): non-associated Cam-Clay with the hardening solve of
// the "#if 1" branch; p0 = bm*1e-5 + sin(xi*max(-logJp,0)) evaluated in double like the reference (:123, sin — not sinh).
ZPC_HD void stress_nacc(float volume, float mu, float bm, float xi, float beta, float Msqr, bool hardeningOn, float &logJp,
                        const float (&Fin)[9], float (&PF)[9]) {
  float F[9], U[9], S[3], V[9], New_F[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  svd3(F, U, S, V);
  const float p0 = (float)((double)(bm * (float)0.00001) + sin((double)(xi * (-logJp > 0 ? -logJp : 0.f))));
  const float p_min = -beta * p0;
  const float Je_trial = S[0] * S[1] * S[2];
  const float B_hat_trial[3] = {S[0] * S[0], S[1] * S[1], S[2] * S[2]};
  const float trace_B_hat_trial_divdim = (B_hat_trial[0] + B_hat_trial[1] + B_hat_trial[2]) / 3.f;
  const float J_power_neg_2_d_mulmu = mu * powf(Je_trial, -2.f / 3.f);
  const float s_hat_trial[3] = {J_power_neg_2_d_mulmu * (B_hat_trial[0] - trace_B_hat_trial_divdim),
                                J_power_neg_2_d_mulmu * (B_hat_trial[1] - trace_B_hat_trial_divdim),
                                J_power_neg_2_d_mulmu * (B_hat_trial[2] - trace_B_hat_trial_divdim)};
  const float psi_kappa_partial_J = bm * 0.5f * (Je_trial - 1.f / Je_trial);
  const float p_trial = -psi_kappa_partial_J * Je_trial;
  const float y_s_half_coeff = 3.f / 2.f * (1 + 2.f * beta);
  const float y_p_half = (Msqr * (p_trial - p_min) * (p_trial - p0));
  const float s_hat_trial_sqrnorm = s_hat_trial[0] * s_hat_trial[0] + s_hat_trial[1] * s_hat_trial[1] + s_hat_trial[2] * s_hat_trial[2];
  const float y = (y_s_half_coeff * s_hat_trial_sqrnorm) + y_p_half;
  if (p_trial > p0 || p_trial < p_min) {  // cases 1 and 2: project to the max / min tip of the yield surface
    const float Je_new = sqrtf(-2.f * (p_trial > p0 ? p0 : p_min) / bm + 1.f);
    S[0] = S[1] = S[2] = powf(Je_new, 1.f / 3.f);
    mat_diag_matT(New_F, U, S, V);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (hardeningOn) logJp += logf(Je_trial / Je_new);
  } else if ((double)y >= 1e-4) {  // case 3, outside the yield surface: project onto it
    const float B_s_coeff = powf(Je_trial, 2.f / 3.f) / mu * sqrtf(-y_p_half / y_s_half_coeff) / sqrtf(s_hat_trial_sqrnorm);
#pragma unroll
    for (int i = 0; i < 3; ++i) S[i] = sqrtf(s_hat_trial[i] * B_s_coeff + trace_B_hat_trial_divdim);
    mat_diag_matT(New_F, U, S, V);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (hardeningOn && (double)p0 > 1e-4 && (double)p_trial < (double)p0 - 1e-4 && (double)p_trial > 1e-4 + (double)p_min) {
      const float p_center = (1.f - beta) * p0 / 2;
      const float q_trial = sqrtf(3.f / 2.f * s_hat_trial_sqrnorm);
      float direction[2] = {p_center - p_trial, -q_trial};
      const float direction_norm = sqrtf(direction[0] * direction[0] + direction[1] * direction[1]);
      direction[0] /= direction_norm;
      direction[1] /= direction_norm;
      const float C = Msqr * (p_center - p_min) * (p_center - p0);
      const float B = Msqr * direction[0] * (2 * p_center - p0 - p_min);
      const float A = Msqr * direction[0] * direction[0] + (1 + 2 * beta) * direction[1] * direction[1];
      const float disc = sqrtf(B * B - 4 * A * C);
      const float l1 = (-B + disc) / (2 * A), l2 = (-B - disc) / (2 * A);
      const float p1 = p_center + l1 * direction[0], p2 = p_center + l2 * direction[0];
      const float p_fake = (p_trial - p_center) * (p1 - p_center) > 0 ? p1 : p2;
      const float tmp_Je_sqr = (-2 * p_fake / bm + 1);
      const float Je_new_fake = sqrtf(fabsf(tmp_Je_sqr));
      if ((double)Je_new_fake > 1e-4) logJp += logf(Je_trial / Je_new_fake);
    }
  }
  const float J = S[0] * S[1] * S[2];
  float b[9];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) b[3 * c + r] = F[r] * F[c] + F[3 + r] * F[3 + c] + F[6 + r] * F[6 + c];
  const float tr3 = (b[0] + b[4] + b[8]) / 3.f;
  const float dev_b_coeff = mu * powf(J, -2.f / 3.f);
  const float i_coeff = bm * .5f * (J * J - 1.f);
#pragma unroll
  for (int i = 0; i < 9; ++i) PF[i] = (i % 4 == 0) ? (dev_b_coeff * (b[i] - tr3) + i_coeff) * volume : (dev_b_coeff * b[i]) * volume;
}

// Individually rounded fp32 operations: the compiler may not contract them into FMAs on the device; the host build
// (tests/hostmath, -ffp-contract=off) uses the plain operators, which is what the reference's host path executes.
#ifdef __CUDA_ARCH__
ZPC_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
ZPC_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
ZPC_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
ZPC_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
ZPC_HD float rn_sqrt(float a) { return __fsqrt_rn(a); }
#else
ZPC_HD float rn_add(float a, float b) { return a + b; }
ZPC_HD float rn_sub(float a, float b) { return a - b; }
ZPC_HD float rn_mul(float a, float b) { return a * b; }
ZPC_HD float rn_div(float a, float b) { return a / b; }
ZPC_HD float rn_sqrt(float a) { return sqrtf(a); }
#endif

// AnalyticLevelSet<Cuboid>::do_getSignedDistance (geometry/AnalyticLevelSet.h:89-96), box [mn, mx].  Every operation is
// the reference's, individually rounded (no FMA contraction): the normal below is a central difference with eps = 1e-6 in
// float — it amplifies a one-ulp difference of the distance to 1e-2 of a normal component, so only a bit-identical
// distance reproduces the reference's normals.
ZPC_HD float cuboid_sdf(float x0, float x1, float x2, const float *mn, const float *mx) {
  const float x[3] = {x0, x1, x2};
  float point[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float center = rn_div(rn_add(mn[i], mx[i]), 2.f);
    const float a = rn_sub(x[i], center);
    point[i] = rn_sub(a > 0.f ? a : -a, rn_div(rn_sub(mx[i], mn[i]), 2.f));
  }
  float mxp = point[0];
  if (point[1] > mxp) mxp = point[1];
  if (point[2] > mxp) mxp = point[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) if (point[i] < 0.f) point[i] = 0.f;
  const float l2 = rn_add(rn_add(rn_add(0.f, rn_mul(point[0], point[0])), rn_mul(point[1], point[1])), rn_mul(point[2], point[2]));
  return rn_add(mxp < 0.f ? mxp : 0.f, rn_sqrt(l2));
}
// ::do_getNormal (:98-110)
ZPC_HD void cuboid_normal(float x0, float x1, float x2, const float *mn, const float *mx, float &n0, float &n1, float &n2) {
  const float eps = (float)1e-6, two_eps = rn_add(eps, eps);
  const float g0 = rn_div(rn_sub(cuboid_sdf(rn_add(x0, eps), x1, x2, mn, mx), cuboid_sdf(rn_sub(x0, eps), x1, x2, mn, mx)), two_eps);
  const float g1 = rn_div(rn_sub(cuboid_sdf(x0, rn_add(x1, eps), x2, mn, mx), cuboid_sdf(x0, rn_sub(x1, eps), x2, mn, mx)), two_eps);
  const float g2 = rn_div(rn_sub(cuboid_sdf(x0, x1, rn_add(x2, eps), mn, mx), cuboid_sdf(x0, x1, rn_sub(x2, eps), mn, mx)), two_eps);
  const float len = rn_sqrt(rn_add(rn_add(rn_add(0.f, rn_mul(g0, g0)), rn_mul(g1, g1)), rn_mul(g2, g2)));
  n0 = rn_div(g0, len); n1 = rn_div(g1, len); n2 = rn_div(g2, len);
}

// Collider::resolveCollision over AnalyticLevelSet<Plane | Sphere | Cuboid> with its rigid motion (geometry/Collider.h:16-24, 98-127,
// geometry/AnalyticLevelSet.h:11-43,130-157): projects the velocity of a node at (px,py,pz) that lies inside the collider
ZPC_HD void collide(const zpc_collider &col, float px, float py, float pz, float &vx, float &vy, float &vz) {
  // material-space position X = R^T (x - b) / s (Collider.h:106-108); products and sums in the reference's order, no contraction
  // where a sign decides
  const float xb0 = px - col.b[0], xb1 = py - col.b[1], xb2 = pz - col.b[2];
  const float inv_s = 1.f / col.s;
  const float *R = col.R;
  const float X0 = rn_mul(rn_add(rn_add(rn_mul(R[0], xb0), rn_mul(R[3], xb1)), rn_mul(R[6], xb2)), inv_s);
  const float X1 = rn_mul(rn_add(rn_add(rn_mul(R[1], xb0), rn_mul(R[4], xb1)), rn_mul(R[7], xb2)), inv_s);
  const float X2 = rn_mul(rn_add(rn_add(rn_mul(R[2], xb0), rn_mul(R[5], xb1)), rn_mul(R[8], xb2)), inv_s);
  const float d0 = X0 - col.origin[0], d1 = X1 - col.origin[1], d2 = X2 - col.origin[2];
  float m0, m1, m2, dist;  // normal in material space
  if (col.geometry == ZPC_GEOM_PLANE) {
    m0 = col.normal[0]; m1 = col.normal[1]; m2 = col.normal[2];
    dist = rn_add(rn_add(rn_mul(m0, d0), rn_mul(m1, d1)), rn_mul(m2, d2));
  } else if (col.geometry == ZPC_GEOM_CUBOID) {  // origin = box min, normal = box max (material space)
    dist = cuboid_sdf(X0, X1, X2, col.origin, col.normal);
    m0 = m1 = m2 = 0.f;
    if (dist < 0.f && col.type != ZPC_COLLIDER_STICKY) cuboid_normal(X0, X1, X2, col.origin, col.normal, m0, m1, m2);
  } else {
    const float l2 = rn_add(rn_add(rn_mul(d0, d0), rn_mul(d1, d1)), rn_mul(d2, d2));
    const float len = sqrtf(l2);
    dist = len - col.normal[0];
    const bool tiny = l2 < 1e-7f;
    m0 = tiny ? 0.f : d0 / len; m1 = tiny ? 0.f : d1 / len; m2 = tiny ? 0.f : d2 / len;
  }
  if (dist < 0.f) {
    // v_object = omega x (x - b) + (ds/dt / s) (x - b) + db/dt (the analytic level sets have no material velocity), :110-111
    const float k = col.dsdt * inv_s;
    const float o0 = (col.omega[1] * xb2 - col.omega[2] * xb1) + k * xb0 + col.dbdt[0];
    const float o1 = (col.omega[2] * xb0 - col.omega[0] * xb2) + k * xb1 + col.dbdt[1];
    const float o2 = (col.omega[0] * xb1 - col.omega[1] * xb0) + k * xb2 + col.dbdt[2];
    if (col.type == ZPC_COLLIDER_STICKY) {
      vx = o0; vy = o1; vz = o2;
    } else {
      vx -= o0; vy -= o1; vz -= o2;
      const float n0 = R[0] * m0 + R[1] * m1 + R[2] * m2, n1 = R[3] * m0 + R[4] * m1 + R[5] * m2, n2 = R[6] * m0 + R[7] * m1 + R[8] * m2;
      const float proj = rn_add(rn_add(rn_mul(n0, vx), rn_mul(n1, vy)), rn_mul(n2, vz));
      if (col.type == ZPC_COLLIDER_SLIP || proj < 0.f) { vx -= proj * n0; vy -= proj * n1; vz -= proj * n2; }
      vx += o0; vy += o1; vz += o2;
    }
  }
}

// LocalArena::init (simulation/Utils.hpp:51-70): base node, in-cell offset (scaled by dx), 3x3 weights
struct Arena {
  int corner[3];
  float local[3];   // (X - corner) * dx
  float w[3][3];
};
// EXACT_DIV_BY_FMA: X = pos / dx through div_exact (the correctly rounded quotient from 1/dx and two FMAs, no slow-path branch): same
// bits as the division the reference writes, 40 fewer instructions per particle than the three IEEE divisions (binned G2P)
template <bool EXACT_DIV_BY_FMA = false>
ZPC_HD void arena_init(Arena &a, float dx, const float (&pos)[3], float dx_inv = 0.f) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float X = EXACT_DIV_BY_FMA ? div_exact(pos[d], dx, dx_inv) : pos[d] / dx;  // reference divides (Utils.hpp:56), it does not multiply by 1/dx
    const int cn = (int)floorf(X - 0.5f);
    const float lp = X - (float)cn;
    const float d0 = lp - floorf(lp - 0.5f);
    a.corner[d] = cn;
    a.w[d][0] = 0.5f * (1.5f - d0) * (1.5f - d0);
    const float d1 = d0 - 1.0f;
    a.w[d][1] = 0.75f - d1 * d1;
    const float zz = 0.5f + d1;
    a.w[d][2] = 0.5f * zz * zz;
    a.local[d] = lp * dx;
  }
}

// home block of a particle as ComputeSparsity assigns it (SparsityOp.hpp:68-79): floor_div(floor(x/dx+.5)-2, 4)
ZPC_HD int sparsity_coord(float x, float dxinv) { return (int)floorf(x * dxinv + 0.5f) - 2; }
ZPC_HD int floor_div4(int c) { return c >> 2; }  // arithmetic shift == floor division by 4

// ---- legacy hash table lookups (container/HashTable.hpp) -----------------------------------------
// do_hash (:496-500) + 64-bit hash_combine (math/Hash.hpp:17-27), then "(h % size + size) % size" (:358)
ZPC_HD int hash_slot0(int kx, int ky, int kz, int table_size) {
  unsigned long long seed = (unsigned long long)(long long)kx;
  seed ^= ((unsigned long long)(long long)ky + 0x9e3779b97f4a7c15ULL + (seed << 12) + (seed >> 4));
  seed ^= ((unsigned long long)(long long)kz + 0x9e3779b97f4a7c15ULL + (seed << 12) + (seed >> 4));
  const int e = (int)(unsigned)seed;
  return (e % table_size + table_size) % table_size;
}
// HashTableView::query (:447-456) along insert's probe sequence (:386)
ZPC_HD int table_query(int kx, int ky, int kz, int table_size, const int *__restrict__ keys,
                                           const int *__restrict__ indices) {
  int slot = hash_slot0(kx, ky, kz, table_size);
  while (true) {
    const int ix = indices[slot];
    if (ix == -1) return -1;
    const int *k = keys + 3 * (size_t)slot;
    if (k[0] == kx && k[1] == ky && k[2] == kz) return ix;
    slot = (slot + 127) % table_size;
  }
}

// ---- bht<i32,3,int,16> lookups (container/Bht.hpp) -------------------------------------------------
// universal_hash over a vec3i key (py_interop/HashUtils.hpp:22-43): sub(k) = ((hashx ^ k) + hashy) % 4294967291 in
// 32-bit unsigned arithmetic, combined with the 32-bit hash_combine
ZPC_HD unsigned bht_hash(unsigned hx, unsigned hy, int kx, int ky, int kz) {
  const unsigned P = 4294967291u;
  unsigned h = ((hx ^ (unsigned)kx) + hy) % P;
  h ^= (((hx ^ (unsigned)ky) + hy) % P) + 0x9e3779b9u + (h << 6) + (h >> 2);
  h ^= (((hx ^ (unsigned)kz) + hy) % P) + 0x9e3779b9u + (h << 6) + (h >> 2);
  return h;
}
// BHTView::query (Bht.hpp:666-700): the 16 slots of the hf0 bucket, then hf1's, then hf2's
ZPC_HD int bht_query(int kx, int ky, int kz, const zpc_bht_view &tb) {
  if (tb.numBuckets == 0) return -1;
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const unsigned b = bht_hash(tb.hf[2 * it], tb.hf[2 * it + 1], kx, ky, kz) % tb.numBuckets * 16u;
    const int4 *k = reinterpret_cast<const int4 *>(tb.keys) + b;
#pragma unroll 4
    for (int s = 0; s < 16; ++s) {
      const int4 c = k[s];
      if (c.x == kx && c.y == ky && c.z == kz) return tb.indices[b + s];
    }
  }
  return -1;
}

}  // namespace zpcm

/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See oracle/oracle.h for the rules.
 *
 * CPU restatement (plain C, fp32, no FMA contraction: build with -ffp-contract=off) of the
 * reference's MPM particle<->sparse-grid path.  Paths cited are relative to
 * /root/reference/include/zensim/.  Operation ORDER follows the reference expression by expression
 * so that, on the host, results are bit-identical to the reference's seq_exec path
 * (tests/test_oracle_vs_ref.py checks exactly that).
 */
#include "oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* container/HashTable.hpp                                                                    */
/* ------------------------------------------------------------------------------------------ */
int zo_next_2pow(int n) { /* math/bit/Bits.h:177-184: 1 << bit_length(n-1) */
  int p = 1;
  if (n <= 0) return 1;
  while (p < n) p <<= 1;
  return p;
}
int zo_table_size_for(int expected) { /* HashTable.hpp:87-90, reserve_ratio_v = 16 (:70) */
  return expected == 0 ? 0 : zo_next_2pow(expected) * 16;
}
/* do_hash (HashTable.hpp:496-500) + math/Hash.hpp:17-27 (64-bit hash_combine), then the
 * "(h % size + size) % size" of insert/query (:358, :449).  key[d] is int -> sign-extended. */
int zo_hash_slot0(const int key[3], int table_size) {
  uint64_t seed = (uint64_t)(int64_t)key[0];
  for (int d = 1; d < 3; ++d)
    seed ^= ((uint64_t)(int64_t)key[d] + 0x9e3779b97f4a7c15ULL + (seed << 12) + (seed >> 4));
  int e = (int)(uint32_t)seed; /* static_cast<value_t>(size_t) */
  return (e % table_size + table_size) % table_size;
}
void zo_table_clear(int table_size, int *keys, int *indices, int *status, int *cnt) {
  for (int e = 0; e < table_size; ++e) { /* SparsityOp.hpp:46-52 */
    keys[3 * e] = keys[3 * e + 1] = keys[3 * e + 2] = INT_MAX;
    indices[e] = -1;
    if (status) status[e] = -1;
  }
  *cnt = 0;
}
static int key_eq(const int *a, const int *b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
int zo_table_insert(const int key[3], int table_size, int *keys, int *indices, int *active_keys,
                    int *cnt) { /* HashTable.hpp:376-398, single-threaded semantics */
  static const int sentinel[3] = {INT_MAX, INT_MAX, INT_MAX};
  int slot = zo_hash_slot0(key, table_size);
  for (;;) {
    int *stored = keys + 3 * slot;
    if (key_eq(stored, sentinel)) {
      memcpy(stored, key, 3 * sizeof(int));
      int no = (*cnt)++;
      indices[slot] = no;
      memcpy(active_keys + 3 * no, key, 3 * sizeof(int));
      return no;
    }
    if (key_eq(stored, key)) return -1;
    slot = (slot + 127) % table_size; /* :386 */
  }
}
int zo_table_query(const int key[3], int table_size, const int *keys, const int *indices) {
  int slot = zo_hash_slot0(key, table_size); /* HashTable.hpp:447-456 */
  for (;;) {
    if (key_eq(keys + 3 * slot, key)) return indices[slot];
    if (indices[slot] == -1) return -1;
    slot += 127;
    if (slot > table_size) slot = slot % table_size; /* sic: '>' not '>=' (:454) */
    if (slot == table_size) return -1; /* the reference reads out of bounds here; never hit in tests */
  }
}

/* ------------------------------------------------------------------------------------------ */
/* simulation/sparsity/SparsityOp.hpp                                                         */
/* ------------------------------------------------------------------------------------------ */
void zo_block_of_particle(const float x[3], float dx, int blockid[3]) { /* :58-79 */
  const float dxinv = (float)1.0 / dx; /* :66 */
  for (int d = 0; d < 3; ++d) {
    int coord = (int)floorf(x[d] * dxinv + 0.5f) + (-2); /* :73-74, offset=-2 displacement=.5 */
    int b = coord + (coord < 0 ? -4 + 1 : 0);            /* :76 */
    blockid[d] = b / 4;                                  /* :77 (C truncation on the shifted value) */
  }
}
int zo_partition_build(int n, const float *x, float dx, int table_size, int *keys, int *indices,
                       int *status, int *active_keys, int *cnt) {
  zo_table_clear(table_size, keys, indices, status, cnt);
  for (int p = 0; p < n; ++p) {
    int b[3];
    zo_block_of_particle(x + 3 * p, dx, b);
    zo_table_insert(b, table_size, keys, indices, active_keys, cnt);
  }
  const int first = *cnt; /* EnlargeSparsity{lo=0, hi=2} over the first table.size() entries, :88-112 */
  for (int i = 0; i < first; ++i) {
    int base[3] = {active_keys[3 * i], active_keys[3 * i + 1], active_keys[3 * i + 2]};
    for (int ox = 0; ox < 2; ++ox)
      for (int oy = 0; oy < 2; ++oy)
        for (int oz = 0; oz < 2; ++oz) {
          int k[3] = {base[0] + ox, base[1] + oy, base[2] + oz};
          zo_table_insert(k, table_size, keys, indices, active_keys, cnt);
        }
  }
  return *cnt;
}

/* index_buckets_for_particles (simulation/particle/Query.tpp:9-58) on the serial policy: a table of the occupied CELLS
 * (ComputeSparsity with blockLen 1, offset 0 and the given displacement, SparsityOp.hpp:58-86), particles counted per cell
 * (SpatiallyCount, :115-151), exclusive scan, particle ids scattered (SpatiallyDistribute, :154-195).  counts / offsets
 * have table.size() + 1 entries.  Serially, the ids of a bucket come out in ascending order. */
static void cell_of_particle(const float x[3], float dxinv, float displacement, int cell[3]) {
  for (int d = 0; d < 3; ++d) cell[d] = (int)floorf(x[d] * dxinv + displacement) + 0;   /* lower_trunc(pos * dxinv + displacement) + offset */
}
int zo_index_buckets(int n, const float *x, float dx, float displacement, int table_size, int *keys, int *indices_tab,
                     int *status, int *active_keys, int *cnt, int *counts, int *offsets, int *indices) {
  const float dxinv = (float)1.0 / dx;
  zo_table_clear(table_size, keys, indices_tab, status, cnt);
  for (int p = 0; p < n; ++p) {
    int c[3];
    cell_of_particle(x + 3 * p, dxinv, displacement, c);
    zo_table_insert(c, table_size, keys, indices_tab, active_keys, cnt);
  }
  const int numCells = *cnt + 1;
  int *tmp = (int *)calloc((size_t)numCells, sizeof(int));
  for (int i = 0; i < numCells; ++i) counts[i] = 0;
  for (int p = 0; p < n; ++p) {
    int c[3];
    cell_of_particle(x + 3 * p, dxinv, displacement, c);
    counts[zo_table_query(c, table_size, keys, indices_tab)] += 1;
  }
  { int run = 0; for (int i = 0; i < numCells; ++i) { offsets[i] = run; run += counts[i]; } }
  for (int p = 0; p < n; ++p) {
    int c[3];
    cell_of_particle(x + 3 * p, dxinv, displacement, c);
    const int cellno = zo_table_query(c, table_size, keys, indices_tab);
    indices[offsets[cellno] + tmp[cellno]++] = p;
  }
  free(tmp);
  return *cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* math/matrix/SVD.hpp:15-1026  — McAdams/Selle/Tamstorf/Teran/Sifakis 3x3 SVD, 4 Jacobi sweeps  */
/* ------------------------------------------------------------------------------------------ */
static float rsqrt_host(float v) { return 1.0f / sqrtf(v); } /* ZpcMathUtils.hpp:842-846 */

/* rsqrt with one Newton-Raphson refinement as written at SVD.hpp:386-392, 701-708 */
static float rsqrt_refined(float t2) {
  float t1 = rsqrt_host(t2);
  float t4 = t1 * 0.5f;
  float t3 = t1 * t4;
  t3 = t1 * t3;
  t3 = t2 * t3;
  t1 = t1 + t4;
  t1 = t1 - t3;
  return t1;
}

/* One cyclic Jacobi step on the symmetric matrix (SVD.hpp:96-178 and its two rotated copies).
 * pp,qq = the two diagonal entries, off = the entry coupling them, rr = the third diagonal entry,
 * a,b = the two remaining off-diagonals in the order the reference updates them.
 * q = quaternion (s,x,y,z); axis = which vector component takes "+sh*qs". */
static void jacobi_step(float *pp, float *qq, float *off, float *rr, float *a, float *b, float q[4],
                        int axis) {
  const float tiny = 1.e-20f, gamma = 5.8284273147583007813f;
  union { float f; uint32_t u; } sp8 = {.u = 1053028117u}, cp8 = {.u = 1064076127u};
  float sh = *off * 0.5f;
  float t5 = *pp - *qq;
  float t2 = sh * sh;
  int big = t2 >= tiny;
  sh = big ? sh : 0.0f;
  float ch = big ? t5 : 1.0f;
  float t1 = sh * sh;
  t2 = ch * ch;
  float t3 = t1 + t2;
  float t4 = rsqrt_host(t3);
  sh = t4 * sh;
  ch = t4 * ch;
  t1 = gamma * t1;
  if (t2 <= t1) { sh = sp8.f; ch = cp8.f; }
  t1 = sh * sh;
  t2 = ch * ch;
  float c = t2 - t1;
  float s = ch * sh;
  s = s + s;
  /* conjugation */
  t3 = t1 + t2;
  *rr = *rr * t3; *a = *a * t3; *b = *b * t3; *rr = *rr * t3;
  t1 = s * *a; t2 = s * *b;
  *a = c * *a; *b = c * *b;
  *a = t2 + *a; *b = *b - t1;
  t2 = s * s;
  t1 = *qq * t2; t3 = *pp * t2;
  t4 = c * c;
  *pp = *pp * t4; *qq = *qq * t4;
  *pp = *pp + t1; *qq = *qq + t3;
  t4 = t4 - t2;
  t2 = *off + *off;
  *off = *off * t4;
  t4 = c * s;
  t2 = t2 * t4; t5 = t5 * t4;
  *pp = *pp + t2; *off = *off - t5; *qq = *qq - t2;
  /* quaternion accumulation */
  float t[3] = {sh * q[1], sh * q[2], sh * q[3]};
  sh = sh * q[0];
  for (int i = 0; i < 4; ++i) q[i] = ch * q[i];
  const int bx = (axis + 1) % 3, cx = (axis + 2) % 3;
  q[1 + axis] = q[1 + axis] + sh;
  q[0] = q[0] - t[axis];
  q[1 + bx] = q[1 + bx] + t[cx];
  q[1 + cx] = q[1 + cx] - t[bx];
}

/* Givens quaternion zeroing `low` against pivot `piv` (SVD.hpp:688-738 and two copies) */
static void qr_givens(float piv, float low, float *c, float *s) {
  const float small = 1.e-12f;
  float sh = low * low;
  sh = (sh >= small) ? low : 0.0f;
  float ch = 0.0f - piv;
  ch = piv > ch ? piv : ch;       /* math::max(ch, piv) = y > x ? y : x, ZpcMathUtils.hpp:299 */
  ch = small > ch ? small : ch;
  int nonneg = piv >= 0.0f;
  float t1 = ch * ch, t2 = sh * sh;
  t2 = t1 + t2;
  t1 = rsqrt_refined(t2);
  t1 = t1 * t2;
  ch = ch + t1;
  if (!nonneg) { float tmp = ch; ch = sh; sh = tmp; }
  t1 = ch * ch; t2 = sh * sh;
  t2 = t1 + t2;
  t1 = rsqrt_refined(t2);
  ch = ch * t1; sh = sh * t1;
  *c = ch * ch; *s = sh * sh;
  *c = *c - *s;
  *s = sh * ch;
  *s = *s + *s;
}
static void rot_pair(float *x, float *y, float c, float s) { /* x' = c x + s y ; y' = c y - s x */
  float t1 = s * *x, t2 = s * *y;
  *x = c * *x; *y = c * *y;
  *x = *x + t2; *y = *y - t1;
}

/* F,U,V column-major 9-vectors (F[3*col+row]) as passed by ConstitutiveModel_Vol_dP.hpp:14-16 */
void zo_svd3(const float F[9], float U[9], float S[3], float V[9]) {
#define A_(r, c) a[(r)*3 + (c)] /* row-major local copies, 0-based */
  float a[9], v[9], u[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A_(r, c) = F[3 * c + r];
  /* normal equations S = A^T A, lower triangle (SVD.hpp:54-88) */
  float s11 = (A_(1,0)*A_(1,0) + A_(0,0)*A_(0,0)); s11 = A_(2,0)*A_(2,0) + s11;
  float s21 = (A_(1,1)*A_(1,0) + A_(0,1)*A_(0,0)); s21 = A_(2,1)*A_(2,0) + s21;
  float s31 = (A_(1,2)*A_(1,0) + A_(0,2)*A_(0,0)); s31 = A_(2,2)*A_(2,0) + s31;
  float s22 = (A_(1,1)*A_(1,1) + A_(0,1)*A_(0,1)); s22 = A_(2,1)*A_(2,1) + s22;
  float s32 = (A_(1,2)*A_(1,1) + A_(0,2)*A_(0,1)); s32 = A_(2,2)*A_(2,1) + s32;
  float s33 = (A_(1,2)*A_(1,2) + A_(0,2)*A_(0,2)); s33 = A_(2,2)*A_(2,2) + s33;
  float q[4] = {1.f, 0.f, 0.f, 0.f};
  for (int sweep = 0; sweep < 4; ++sweep) { /* :96 */
    jacobi_step(&s11, &s22, &s21, &s33, &s31, &s32, q, 2); /* (1,2), :97-178 */
    jacobi_step(&s22, &s33, &s32, &s11, &s21, &s31, q, 0); /* (2,3), :183-262 */
    jacobi_step(&s33, &s11, &s31, &s22, &s32, &s21, q, 1); /* (3,1), :269-371 */
  }
  { /* normalise quaternion, :378-397 */
    float t2 = q[0] * q[0];
    t2 = q[1] * q[1] + t2; t2 = q[2] * q[2] + t2; t2 = q[3] * q[3] + t2;
    float t1 = rsqrt_refined(t2);
    for (int i = 0; i < 4; ++i) q[i] = q[i] * t1;
  }
  { /* quaternion -> V, :403-430 */
    float t1 = q[1]*q[1], t2 = q[2]*q[2], t3 = q[3]*q[3];
    float v11 = q[0]*q[0];
    float v22 = v11 - t1;
    float v33 = v22 - t2; v33 = v33 + t3;
    v22 = v22 + t2; v22 = v22 - t3;
    v11 = v11 + t1; v11 = v11 - t2; v11 = v11 - t3;
    t1 = q[1] + q[1]; t2 = q[2] + q[2]; t3 = q[3] + q[3];
    float v32 = q[0]*t1, v13 = q[0]*t2, v21 = q[0]*t3;
    t1 = q[2]*t1; t2 = q[3]*t2; t3 = q[1]*t3;
    float v12 = t1 - v21, v23 = t2 - v32, v31 = t3 - v13;
    v21 = t1 + v21; v32 = t2 + v32; v13 = t3 + v13;
    v[0]=v11; v[1]=v12; v[2]=v13; v[3]=v21; v[4]=v22; v[5]=v23; v[6]=v31; v[7]=v32; v[8]=v33;
  }
#define V_(r, c) v[(r)*3 + (c)]
  for (int r = 0; r < 3; ++r) { /* A <- A V, :436-488 */
    float x = A_(r,0), y = A_(r,1), z = A_(r,2);
    for (int c = 0; c < 3; ++c) {
      float acc = V_(0,c) * x;
      acc = acc + V_(1,c) * y;
      acc = acc + V_(2,c) * z;
      A_(r,c) = acc;
    }
  }
  /* sort columns by norm, :494-676 */
  float rho[3];
  for (int c = 0; c < 3; ++c) {
    float t = A_(0,c)*A_(0,c);
    t = t + A_(1,c)*A_(1,c);
    t = t + A_(2,c)*A_(2,c);
    rho[c] = t;
  }
  static const int swp[3][3] = {{0, 1, 1}, {0, 2, 0}, {1, 2, 2}}; /* (i, j, column to negate) */
  for (int k = 0; k < 3; ++k) {
    int i = swp[k][0], j = swp[k][1], neg = swp[k][2];
    if (rho[i] < rho[j]) {
      for (int r = 0; r < 3; ++r) {
        float t = A_(r,i); A_(r,i) = A_(r,j); A_(r,j) = t;
        t = V_(r,i); V_(r,i) = V_(r,j); V_(r,j) = t;
      }
      float t = rho[i]; rho[i] = rho[j]; rho[j] = t;
      for (int r = 0; r < 3; ++r) { A_(r,neg) = A_(r,neg) * -1.0f; V_(r,neg) = V_(r,neg) * -1.0f; }
    }
  }
  /* QR of A V by three Givens rotations, :682-1000 */
  for (int i = 0; i < 9; ++i) u[i] = (i % 4 == 0) ? 1.f : 0.f;
#define U_(r, c) u[(r)*3 + (c)]
  static const int piv[3][3] = {{0, 0, 1}, {0, 0, 2}, {1, 1, 2}}; /* pivot (r,c) ... and lower row */
  for (int k = 0; k < 3; ++k) {
    int p = piv[k][0], lo = piv[k][2];
    float c, s;
    qr_givens(A_(p, piv[k][1]), A_(lo, piv[k][1]), &c, &s);
    for (int col = 0; col < 3; ++col) rot_pair(&A_(p,col), &A_(lo,col), c, s);
    for (int row = 0; row < 3; ++row) rot_pair(&U_(row,p), &U_(row,lo), c, s);
  }
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { U[3*c + r] = U_(r,c); V[3*c + r] = V_(r,c); }
  S[0] = A_(0,0); S[1] = A_(1,1); S[2] = A_(2,2);
#undef A_
#undef V_
#undef U_
}

void zo_lame(float E, float nu, float *mu, float *lam) { /* ConstitutiveModel.hpp:34-38, T=float */
  *mu = (float)(0.5 * E / (1 + nu));               /* 0.5 is a double literal; (1+nu) is float */
  *lam = (float)(E * nu / ((1 + nu) * (1 - 2 * nu)));
}

void zo_stress_fixedcorotated(float volume, float mu, float lam, const float F[9], float PF[9]) {
  float U[9], S[3], V[9], P[9], Ph[3]; /* ConstitutiveModel_Vol_dP.hpp:10-47 */
  zo_svd3(F, U, S, V);
  float J = S[0] * S[1] * S[2];
  float smu = 2.f * mu, sl = lam * (J - 1.f);
  Ph[0] = smu * (S[0] - 1.f) + sl * (S[1] * S[2]);
  Ph[1] = smu * (S[1] - 1.f) + sl * (S[0] * S[2]);
  Ph[2] = smu * (S[2] - 1.f) + sl * (S[0] * S[1]);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) /* P[3c+r] = sum_k Ph[k] U[3k+r] V[3k+c], :26-34 */
      P[3*c + r] = (Ph[0]*U[r]*V[c] + Ph[1]*U[3 + r]*V[3 + c]) + Ph[2]*U[6 + r]*V[6 + c];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) /* PF' : PF[3c+r] = sum_k P[3k+r] F[3k+c], :37-45 */
      PF[3*c + r] = ((P[r]*F[c] + P[3 + r]*F[3 + c]) + P[6 + r]*F[6 + c]) * volume;
}

/* math::sqrtNewtonRaphson<float>, math/MathUtils.h:239-251: Newton iteration from 1 until |x_{n+1} - x_n| <= max(n * 1e-6, 128 eps) */
static float sqrt_newton_raphson(float n) {
  const float eps = 128.f * 1.1920928955078125e-7f, relTol = 1e-6f;
  if (n < -eps) return NAN;
  if (n < eps) return 0.f;
  float xn = 1.f;
  float xnp1 = 0.5f * (xn + n / xn);
  const float tol = n * relTol > eps ? n * relTol : eps;
  for (; fabsf(xnp1 - xn) > tol; xnp1 = 0.5f * (xn + n / xn)) xn = xnp1;
  return xnp1;
}

/* compute_stress_vonmisesfixedcorotated, physics/ConstitutiveModel_Vol_dP.hpp:49-110: fixed-corotated trial stress
 * in principal space, radial return onto the von Mises cylinder, projected singular values, then the fixed-corotated
 * P F^T with the projected F (the projection stays local: P2G.hpp:85-91 passes a copy of the particle's F) */
void zo_stress_vonmises(float volume, float mu, float lam, float yield_stress, const float Fin[9], float PF[9]) {
  float F[9], U[9], S[3], V[9], Sc[3], tau[3], s_trial[3], P[9], Ph[3];
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  zo_svd3(F, U, S, V);
  for (int d = 0; d < 3; ++d) Sc[d] = 1e-4f > S[d] ? 1e-4f : S[d];            /* :60 */
  float J = Sc[0] * Sc[1] * Sc[2];                                            /* vec::prod: left fold */
  for (int d = 0; d < 3; ++d) tau[d] = 2 * mu * (Sc[d] - 1) * Sc[d] + lam * (J - 1) * J; /* :63-64 */
  const float trace_tau = (tau[0] + tau[1]) + tau[2];
  for (int d = 0; d < 3; ++d) s_trial[d] = tau[d] - (trace_tau / 3.f);
  const float s_norm = sqrt_newton_raphson((s_trial[0] * s_trial[0] + s_trial[1] * s_trial[1]) + s_trial[2] * s_trial[2]);
  const float scaled_tauy = sqrt_newton_raphson(2.f / (6.f - 3)) * yield_stress;  /* :68 */
  if (s_norm - scaled_tauy > 0) {
    const float alpha = scaled_tauy / s_norm;
    J = 1.f;
    for (int d = 0; d < 3; ++d) {
      const float tau_new = alpha * s_trial[d] + (trace_tau / 3.f);
      const float b2m4ac = mu * mu - 2 * mu * (lam * (J - 1) * J - tau_new);
      const float sq = b2m4ac < 0 ? 0 : sqrt_newton_raphson(b2m4ac);
      S[d] = (mu + sq) / (2 * mu);
    }
    /* F = U diag(S) V^T, math/matrix/MatrixUtils.h:26-49 */
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r)
        F[3 * c + r] = (U[r] * S[0] * V[c] + U[3 + r] * S[1] * V[3 + c]) + U[6 + r] * S[2] * V[6 + c];
  }
  J = S[0] * S[1] * S[2];
  const float smu = 2.f * mu, sl = lam * (J - 1.f);
  Ph[0] = smu * (S[0] - 1.f) + sl * (S[1] * S[2]);
  Ph[1] = smu * (S[1] - 1.f) + sl * (S[0] * S[2]);
  Ph[2] = smu * (S[2] - 1.f) + sl * (S[0] * S[1]);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      P[3*c + r] = (Ph[0]*U[r]*V[c] + Ph[1]*U[3 + r]*V[3 + c]) + Ph[2]*U[6 + r]*V[6 + c];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      PF[3*c + r] = ((P[r]*F[c] + P[3 + r]*F[3 + c]) + P[6 + r]*F[6 + c]) * volume;
}

/* math::q_rsqrt(float) + math::sqrt(float), math/MathUtils.h:258-269, 288-323: software square root (bit trick, Newton
 * steps, one correction); restated operation for operation — without FMA contraction it is not always the correctly
 * rounded root, and the plastic models below go through it */
static float zo_q_rsqrt(float number) {
  uint32_t i;
  float x2 = number * 0.5f, y = number;
  memcpy(&i, &y, 4);
  i = 0x5f375a86 - (i >> 1);
  memcpy(&y, &i, 4);
  y = y * (1.5f - (x2 * y * y));
  y = y * (1.5f - (x2 * y * y));
  return y;
}
float zo_math_sqrt(float arg) {
  const float scale_in = 0x1.0p+26f, scale_out = 0x1.0p-13f;
  if (arg < 0.0f) { uint32_t u = 0x7fffffffu; float f; memcpy(&f, &u, 4); return f; }
  if ((arg == 0.0f) || !(fabsf(arg) < INFINITY)) return arg + arg;
  arg = arg * scale_in;
  float rsq = zo_q_rsqrt(arg);
  rsq = ((-0.5f * arg * rsq) * rsq + 0.5f) * rsq + rsq;
  rsq = ((-0.5f * arg * rsq) * rsq + 0.5f) * rsq + rsq;
  float sqt = rsq * arg;
  float err = sqt * -sqt + arg;
  sqt = (0.5f * rsq * err + sqt);
  sqt = sqt * scale_out;
  return sqt;
}

/* matmul_mat_diag_matT_3D, math/matrix/MatrixUtils.h:26-47 */
static void mat_diag_matT(float out[9], const float a[9], const float d[3], const float b[9]) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      out[3 * c + r] = (a[r] * d[0] * b[c] + a[3 + r] * d[1] * b[3 + c]) + a[6 + r] * d[2] * b[6 + c];
}
/* PF = P F^T * volume, ConstitutiveModel_Vol_dP.hpp:37-45, 317-325 */
static void p_ft_vol(float PF[9], const float P[9], const float F[9], float volume) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      PF[3 * c + r] = ((P[r] * F[c] + P[3 + r] * F[3 + c]) + P[6 + r] * F[6 + c]) * volume;
}

/* NACCConfig::bulk(), Msqr() — physics/ConstitutiveModel.hpp:756-776 (fa is passed to sin as is: radians) */
float zo_nacc_bulk(float E, float nu) { return 2.f / 3.f * (E / (2 * (1 + nu))) + (E * nu / ((1 + nu) * (1 - 2 * nu))); }
float zo_nacc_msqr(float fa, int dim) {
  float sin_phi = sinf(fa);
  float mohr = sqrtf(2.f / 3.f) * 2.f * sin_phi / (3.f - sin_phi);
  float M = mohr * dim / sqrtf(2.f / (6.f - dim));
  return M * M;
}

/* compute_stress_sand, physics/ConstitutiveModel_Vol_dP.hpp:242-326: Drucker-Prager plasticity in log-strain space
 * (Klar et al. 2016) on top of the St. Venant-Kirchhoff-with-Hencky-strain energy; the projected F stays local
 * (P2G.hpp:85 passes a copy), logJp is written back (P2G.hpp:101) */
void zo_stress_sand(float volume, float mu, float lam, float cohesion, float beta, float yieldSurface, int volCorrection,
                    float *logJp_io, const float Fin[9], float PF[9]) {
  float F[9], U[9], S[3], V[9];
  float logJp = *logJp_io;
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  zo_svd3(F, U, S, V);
  const float scaled_mu = 2.f * mu;
  float epsilon[3], New_S[3] = {0.f, 0.f, 0.f}, New_F[9], epsilon_hat[3];
  for (int i = 0; i < 3; ++i) {
    float abs_S = S[i] > 0 ? S[i] : -S[i];
    abs_S = (double)abs_S > 1e-4 ? abs_S : (float)1e-4;      /* :255 — the comparison is in double */
    epsilon[i] = logf(abs_S) - cohesion;
  }
  const float sum_epsilon = epsilon[0] + epsilon[1] + epsilon[2];
  const float trace_epsilon = sum_epsilon + logJp;
  for (int i = 0; i < 3; ++i) epsilon_hat[i] = epsilon[i] - (trace_epsilon / 3.f);
  const float epsilon_hat_norm
      = zo_math_sqrt(epsilon_hat[0] * epsilon_hat[0] + epsilon_hat[1] * epsilon_hat[1] + epsilon_hat[2] * epsilon_hat[2]);
  if (trace_epsilon >= 0.f) {                                  /* case II: the cone tip, :270-278 */
    New_S[0] = New_S[1] = New_S[2] = expf(cohesion);
    mat_diag_matT(New_F, U, New_S, V);
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (volCorrection) logJp = beta * sum_epsilon + logJp;
  } else if (mu != 0) {                                        /* :279-297 */
    logJp = 0;
    const float delta_gamma = epsilon_hat_norm + (3.f * lam + scaled_mu) / scaled_mu * trace_epsilon * yieldSurface;
    float H[3];
    if (delta_gamma <= 0) {                                    /* case I: inside the cone */
      for (int i = 0; i < 3; ++i) H[i] = epsilon[i] + cohesion;
    } else {                                                   /* case III: onto the cone surface */
      for (int i = 0; i < 3; ++i) H[i] = epsilon[i] - (delta_gamma / epsilon_hat_norm) * epsilon_hat[i] + cohesion;
    }
    for (int i = 0; i < 3; ++i) New_S[i] = expf(H[i]);
    mat_diag_matT(New_F, U, New_S, V);
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
  }
  const float New_S_log[3] = {logf(New_S[0]), logf(New_S[1]), logf(New_S[2])};
  const float trace_log_S = New_S_log[0] + New_S_log[1] + New_S_log[2];
  float P_hat[3], P[9];
  for (int i = 0; i < 3; ++i) P_hat[i] = (scaled_mu * New_S_log[i] + lam * trace_log_S) / New_S[i];
  mat_diag_matT(P, U, P_hat, V);
  p_ft_vol(PF, P, F, volume);
  *logJp_io = logJp;
}

/* compute_stress_nacc, physics/ConstitutiveModel_Vol_dP.hpp:116-240: non-associated Cam-Clay (Wolper et al. 2019) with
 * the hardening solve "in 19 Josh Fracture paper" (the #if 1 branch).  p0 uses sin, not sinh, as the reference does. */
void zo_stress_nacc(float volume, float mu, float lam, float bm, float xi, float beta, float Msqr, int hardeningOn,
                    float *logJp_io, const float Fin[9], float PF[9]) {
  (void)lam;
  float F[9], U[9], S[3], V[9], New_F[9];
  float logJp = *logJp_io;
  for (int d = 0; d < 9; ++d) F[d] = Fin[d];
  zo_svd3(F, U, S, V);
  const float p0 = (float)((double)(bm * (float)0.00001) + sin((double)(xi * (-logJp > 0 ? -logJp : 0))));  /* :123 */
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
  const float s_hat_trial_sqrnorm
      = s_hat_trial[0] * s_hat_trial[0] + s_hat_trial[1] * s_hat_trial[1] + s_hat_trial[2] * s_hat_trial[2];
  const float y = (y_s_half_coeff * s_hat_trial_sqrnorm) + y_p_half;
  if (p_trial > p0) {                                          /* case 1: max tip of the yield surface, :148-156 */
    const float Je_new = zo_math_sqrt(-2.f * p0 / bm + 1.f);
    S[0] = S[1] = S[2] = powf(Je_new, 1.f / 3.f);
    mat_diag_matT(New_F, U, S, V);
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (hardeningOn) logJp += logf(Je_trial / Je_new);
  } else if (p_trial < p_min) {                                /* case 2: min tip, :159-167 */
    const float Je_new = zo_math_sqrt(-2.f * p_min / bm + 1.f);
    S[0] = S[1] = S[2] = powf(Je_new, 1.f / 3.f);
    mat_diag_matT(New_F, U, S, V);
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (hardeningOn) logJp += logf(Je_trial / Je_new);
  } else if ((double)y >= 1e-4) {                              /* case 3, outside the yield surface, :170-218 */
    const float B_s_coeff = powf(Je_trial, 2.f / 3.f) / mu * zo_math_sqrt(-y_p_half / y_s_half_coeff)
                            / zo_math_sqrt(s_hat_trial_sqrnorm);
    for (int i = 0; i < 3; ++i) S[i] = zo_math_sqrt(s_hat_trial[i] * B_s_coeff + trace_B_hat_trial_divdim);
    mat_diag_matT(New_F, U, S, V);
    for (int i = 0; i < 9; ++i) F[i] = New_F[i];
    if (hardeningOn && (double)p0 > 1e-4 && (double)p_trial < (double)p0 - 1e-4 && (double)p_trial > 1e-4 + (double)p_min) {
      const float p_center = (1.f - beta) * p0 / 2;
      const float q_trial = zo_math_sqrt(3.f / 2.f * s_hat_trial_sqrnorm);
      float direction[2] = {p_center - p_trial, -q_trial};
      const float direction_norm = zo_math_sqrt(direction[0] * direction[0] + direction[1] * direction[1]);
      direction[0] /= direction_norm;
      direction[1] /= direction_norm;
      const float C = Msqr * (p_center - p_min) * (p_center - p0);
      const float B = Msqr * direction[0] * (2 * p_center - p0 - p_min);
      const float A = Msqr * direction[0] * direction[0] + (1 + 2 * beta) * direction[1] * direction[1];
      const float l1 = (-B + zo_math_sqrt(B * B - 4 * A * C)) / (2 * A);
      const float l2 = (-B - zo_math_sqrt(B * B - 4 * A * C)) / (2 * A);
      const float p1 = p_center + l1 * direction[0];
      const float p2 = p_center + l2 * direction[0];
      const float p_fake = (p_trial - p_center) * (p1 - p_center) > 0 ? p1 : p2;
      const float tmp_Je_sqr = (-2 * p_fake / bm + 1);
      const float Je_new_fake = zo_math_sqrt(tmp_Je_sqr > 0 ? tmp_Je_sqr : -tmp_Je_sqr);
      if ((double)Je_new_fake > 1e-4) logJp += logf(Je_trial / Je_new_fake);
    }
  }
  /* elasticity, :224-239: J from the (renewed) S, b = F F^T, deviatoric part */
  const float J = S[0] * S[1] * S[2];
  float b[9], b_dev[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) b[3 * c + r] = (F[r] * F[c] + F[3 + r] * F[3 + c]) + F[6 + r] * F[6 + c]; /* MatrixUtils.h:244-256 */
  const float tr3 = ((b[0] + b[4]) + b[8]) / 3.f;                                                          /* :258-271 */
  for (int i = 0; i < 9; ++i) b_dev[i] = (i % 4 == 0) ? b[i] - tr3 : b[i];
  const float dev_b_coeff = mu * powf(J, -2.f / 3.f);
  const float i_coeff = bm * .5f * (J * J - 1.f);
  for (int i = 0; i < 9; ++i) PF[i] = (i % 4 == 0) ? (dev_b_coeff * b_dev[i] + i_coeff) * volume : (dev_b_coeff * b_dev[i]) * volume;
  *logJp_io = logJp;
}

/* ------------------------------------------------------------------------------------------ */
/* simulation/Utils.hpp:30-174 LocalArena (collocated, quadratic) + InterpolationKernel.hpp   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int corner[3]; float local[3]; float w[3][3]; float dx; } zo_arena;
static void arena_init(zo_arena *a, float dx, const float pos[3]) {
  a->dx = dx;
  for (int d = 0; d < 3; ++d) {
    float X = pos[d] / dx;                                  /* Utils.hpp:56 */
    a->corner[d] = (int)floorf(X - 0.5f);                   /* base_node<1>, InterpolationKernel.hpp:46-55 */
    float lp = X - (float)a->corner[d];                     /* Utils.hpp:60 */
    float d0 = lp - (float)(int)floorf(lp - 0.5f);          /* InterpolationKernel.hpp:106 */
    a->w[d][0] = 0.5f * (1.5f - d0) * (1.5f - d0);          /* :107 */
    float d1 = d0 - 1.0f;
    a->w[d][1] = 0.75f - d1 * d1;
    float zz = 0.5f + d1;
    a->w[d][2] = 0.5f * zz * zz;
    a->local[d] = lp * dx;                                  /* Utils.hpp:67 */
  }
}
static int floor_div4(int c, int *loc) { /* Utils.hpp:20-28: loc = c & 3 ; block = (c - loc)/4 */
  *loc = c & 3;
  return (c - *loc) / 4;
}

void zo_clean_grid(int nblocks, float *grid) { /* GridOp.hpp:54-69 */
  memset(grid, 0, sizeof(float) * 7 * 64 * (size_t)nblocks);
}

/* P2GTransfer for the F-based models (P2G.hpp:84-102): model 0 = FixedCorotatedConfig, 1 = VonMisesFixedCorotatedConfig
 * (prm = {yieldStress}), 2 = DruckerPragerConfig (prm = {cohesion, beta, yieldSurface, volumeCorrection}),
 * 3 = NACCConfig (prm = {xi, beta, hardeningOn, fa, dim}); the plastic models read and write logJp (:93, :101) */
static void p2g_elastic(int model, const float *prm, float *logJp, int n, const float *x, const float *v, const float *m, const float *C,
                        const float *F, float dx, float dt, float E, float nu, float volume, int table_size,
                        const int *keys, const int *indices, float *grid) {
  const float dx_inv = (float)1.0 / dx;        /* P2G.hpp:43 */
  const float D_inv = 4.f * dx_inv * dx_inv;   /* :51 */
  float mu, lam;
  zo_lame(E, nu, &mu, &lam);                   /* :84 */
  for (int p = 0; p < n; ++p) {
    float contrib[9];
    const float *Cp = C + 9 * p, *vel = v + 3 * p;
    const float mass = m[p];
    if (model == 0) zo_stress_fixedcorotated(volume, mu, lam, F + 9 * p, contrib);   /* :87 */
    else if (model == 1) zo_stress_vonmises(volume, mu, lam, prm[0], F + 9 * p, contrib); /* :89-90 */
    else if (model == 2) zo_stress_sand(volume, mu, lam, prm[0], prm[1], prm[2], prm[3] != 0.f, logJp + p, F + 9 * p, contrib); /* :94-96 */
    else zo_stress_nacc(volume, mu, lam, zo_nacc_bulk(E, nu), prm[0], prm[1], zo_nacc_msqr(prm[3], (int)prm[4]), prm[2] != 0.f,
                        logJp + p, F + 9 * p, contrib);                                  /* :97-100 */
    for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;        /* :104 */
    zo_arena ar;
    arena_init(&ar, dx, x + 3 * p);                                           /* :107 */
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) { /* :108 */
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];                          /* Utils.hpp:163-165 */
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];       /* Utils.hpp:77-82 */
      int bno = zo_table_query(blk, table_size, keys, indices);
      float *tile = grid + (size_t)bno * 7 * 64;
      int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];                      /* Structure.hpp:851-859 */
      tile[cell] += mass * W;                                                 /* P2G.hpp:114 */
      for (int d = 0; d < 3; ++d) {
        tile[(1 + d) * 64 + cell] += W * mass * (vel[d] + (Cp[d] * xixp[0] + Cp[3 + d] * xixp[1] + Cp[6 + d] * xixp[2])); /* :117-119 */
        tile[(4 + d) * 64 + cell] += (contrib[d] * xixp[0] + contrib[3 + d] * xixp[1] + contrib[6 + d] * xixp[2]) * W;     /* :121-123 */
      }
    }
  }
}

void zo_p2g_fcr(int n, const float *x, const float *v, const float *m, const float *C,
                const float *F, float dx, float dt, float E, float nu, float volume, int table_size,
                const int *keys, const int *indices, float *grid) {
  p2g_elastic(0, NULL, NULL, n, x, v, m, C, F, dx, dt, E, nu, volume, table_size, keys, indices, grid);
}
void zo_p2g_vonmises(int n, const float *x, const float *v, const float *m, const float *C, const float *F, float dx,
                     float dt, float E, float nu, float yield_stress, float volume, int table_size, const int *keys,
                     const int *indices, float *grid) {
  p2g_elastic(1, &yield_stress, NULL, n, x, v, m, C, F, dx, dt, E, nu, volume, table_size, keys, indices, grid);
}
void zo_p2g_sand(int n, const float *x, const float *v, const float *m, const float *C, const float *F, float *logJp,
                 float dx, float dt, float E, float nu, float cohesion, float beta, float yieldSurface, int volCorrection,
                 float volume, int table_size, const int *keys, const int *indices, float *grid) {
  const float prm[4] = {cohesion, beta, yieldSurface, volCorrection ? 1.f : 0.f};
  p2g_elastic(2, prm, logJp, n, x, v, m, C, F, dx, dt, E, nu, volume, table_size, keys, indices, grid);
}
void zo_p2g_nacc(int n, const float *x, const float *v, const float *m, const float *C, const float *F, float *logJp,
                 float dx, float dt, float E, float nu, float fa, float xi, float beta, int hardeningOn, int dim,
                 float volume, int table_size, const int *keys, const int *indices, float *grid) {
  const float prm[5] = {xi, beta, hardeningOn ? 1.f : 0.f, fa, (float)dim};
  p2g_elastic(3, prm, logJp, n, x, v, m, C, F, dx, dt, E, nu, volume, table_size, keys, indices, grid);
}

/* EquationOfStateConfig branch of P2GTransfer, P2G.hpp:66-87 (weakly compressible fluid: J instead of F) */
void zo_p2g_eos(int n, const float *x, const float *v, const float *m, const float *C, const float *Jp,
                float dx, float dt, float bulk, float viscosity, float volume, int table_size,
                const int *keys, const int *indices, float *grid) {
  const float dx_inv = (float)1.0 / dx;
  const float D_inv = 4.f * dx_inv * dx_inv;
  for (int p = 0; p < n; ++p) {
    float contrib[9];
    const float *Cp = C + 9 * p, *vel = v + 3 * p;
    const float mass = m[p];
    float J = Jp[p];
    float vol = volume * J;                      /* :68 */
    float pressure = bulk;
    {
      float J2 = J * J;
      float J4 = J2 * J2;
      pressure = pressure * (1 / (J * J2 * J4) - 1); /* :74 */
    }
    contrib[0] = ((Cp[0] + Cp[0]) * viscosity - pressure) * vol;
    contrib[1] = (Cp[1] + Cp[3]) * viscosity * vol;
    contrib[2] = (Cp[2] + Cp[6]) * viscosity * vol;
    contrib[3] = (Cp[3] + Cp[1]) * viscosity * vol;
    contrib[4] = ((Cp[4] + Cp[4]) * viscosity - pressure) * vol;
    contrib[5] = (Cp[5] + Cp[7]) * viscosity * vol;
    contrib[6] = (Cp[6] + Cp[2]) * viscosity * vol;
    contrib[7] = (Cp[7] + Cp[5]) * viscosity * vol;
    contrib[8] = ((Cp[8] + Cp[8]) * viscosity - pressure) * vol;
    for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * -dt * D_inv;        /* :104 */
    zo_arena ar;
    arena_init(&ar, dx, x + 3 * p);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];
      int bno = zo_table_query(blk, table_size, keys, indices);
      float *tile = grid + (size_t)bno * 7 * 64;
      int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];
      tile[cell] += mass * W;
      for (int d = 0; d < 3; ++d) {
        tile[(1 + d) * 64 + cell] += W * mass * (vel[d] + (Cp[d] * xixp[0] + Cp[3 + d] * xixp[1] + Cp[6 + d] * xixp[2]));
        tile[(4 + d) * 64 + cell] += (contrib[d] * xixp[0] + contrib[3 + d] * xixp[1] + contrib[6 + d] * xixp[2]) * W;
      }
    }
  }
}

void zo_grid_update(int nblocks, float *grid, float dt, const float extf[3], int mode,
                    float *max_vel_sqr) {
  float mx = *max_vel_sqr;
  for (int b = 0; b < nblocks; ++b) {
    float *tile = grid + (size_t)b * 7 * 64;
    for (int c = 0; c < 64; ++c) {
      if (mode == 1) for (int d = 0; d < 3; ++d) tile[(1 + d) * 64 + c] += tile[(4 + d) * 64 + c];
      float mass = tile[c];                       /* GridOp.hpp:94 */
      if (mass != 0.f) {
        mass = 1.f / mass;                        /* :96 */
        float nrm = 0.f;
        for (int d = 0; d < 3; ++d) {
          float vd = tile[(1 + d) * 64 + c] * mass + extf[d] * dt;  /* :97 */
          tile[(1 + d) * 64 + c] = vd;
          nrm += vd * vd;                         /* l2NormSqr: sequential sum from 0 */
        }
        if (nrm > mx) mx = nrm;                   /* atomic_max, :104 */
      }
    }
  }
  *max_vel_sqr = mx;
}

/* GridMomentumToVelocity (GridOp.hpp:184-214): v = mv * (1/m) on cells with m != 0; max |v|^2.  nch channels per block,
 * mass in channel mChn, momentum in mvChn..mvChn+2. */
void zo_grid_momentum_to_velocity(int nblocks, int nch, float *grid, int mChn, int mvChn, float *max_vel_sqr) {
  float mx = *max_vel_sqr;
  for (int b = 0; b < nblocks; ++b) {
    float *tile = grid + (size_t)b * nch * 64;
    for (int c = 0; c < 64; ++c) {
      float mass = tile[mChn * 64 + c];             /* :200 */
      if (mass != 0.f) {
        mass = 1.f / mass;                          /* :202 */
        float nrm = 0.f;
        for (int d = 0; d < 3; ++d) {
          float vd = tile[(mvChn + d) * 64 + c] * mass;   /* :203 */
          tile[(mvChn + d) * 64 + c] = vd;
          nrm += vd * vd;
        }
        if (nrm > mx) mx = nrm;                     /* :208-209 */
      }
    }
  }
  *max_vel_sqr = mx;
}

/* GridAngularMomentum (GridOp.hpp:216-262): out6[0..2] += x cross mv, out6[3..5] += mv, per cell with m != 0;
 * x = (blockkey*4 + cell coord)*dx and the cross product in float (Vec cross, VecInterface.hpp:945-956), the sums in double.
 * The reference adds with atomics in launch order; this loop adds in (block, cell) order. */
void zo_grid_angular_momentum(int nblocks, int nch, const int *active_keys, const float *grid, float dx, int mChn, int mvChn,
                              double *out6) {
  for (int b = 0; b < nblocks; ++b) {
    const float *tile = grid + (size_t)b * nch * 64;
    for (int c = 0; c < 64; ++c) {
      if (tile[mChn * 64 + c] == 0.f) continue;     /* :235 */
      const int cc[3] = {(c >> 4) & 3, (c >> 2) & 3, c & 3};
      float x[3], mv[3];
      for (int d = 0; d < 3; ++d) {
        x[d] = ((float)active_keys[3 * b + d] * 4.f + (float)cc[d]) * dx;   /* :237-239 */
        mv[d] = tile[(mvChn + d) * 64 + c];
      }
      const float r0 = x[1] * mv[2] - x[2] * mv[1], r1 = x[2] * mv[0] - x[0] * mv[2], r2 = x[0] * mv[1] - x[1] * mv[0];
      out6[0] += (double)r0; out6[1] += (double)r1; out6[2] += (double)r2;
      for (int d = 0; d < 3; ++d) out6[3 + d] += (double)mv[d];
    }
  }
}

void zo_g2p(int n, float *x, float *v, float *C, float *F, float dx, float dt, int table_size,
            const int *keys, const int *indices, const float *grid) {
  const float dx_inv = (float)1 / dx;            /* G2P.hpp:45 */
  const float D_inv = 4.f * dx_inv * dx_inv;     /* :47 */
  for (int p = 0; p < n; ++p) {
    float pos[3] = {x[3*p], x[3*p+1], x[3*p+2]}, vel[3] = {0, 0, 0}, Cn[9] = {0};
    zo_arena ar;
    arena_init(&ar, dx, pos);                    /* :56 */
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3], vi[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];
      int bno = zo_table_query(blk, table_size, keys, indices);
      const float *tile = grid + (size_t)bno * 7 * 64;
      int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];
      for (int d = 0; d < 3; ++d) { vi[d] = tile[(1 + d) * 64 + cell]; vel[d] += vi[d] * W; } /* :63-64 */
      for (int d = 0; d < 9; ++d) Cn[d] += W * vi[d % 3] * xixp[d / 3] * D_inv;               /* :65 */
    }
    for (int d = 0; d < 3; ++d) pos[d] += vel[d] * dt;                                        /* :67 */
    float tmp[9], Fo[9], Fn[9];
    memcpy(Fo, F + 9 * p, sizeof Fo);
    for (int d = 0; d < 9; ++d) tmp[d] = Cn[d] * dt + ((d & 0x3) ? 0.f : 1.f);                /* :76 */
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) /* MatrixUtils.h:136-146, column-major a*b */
      Fn[3*c + r] = (tmp[r] * Fo[3*c] + tmp[3 + r] * Fo[3*c + 1]) + tmp[6 + r] * Fo[3*c + 2];
    memcpy(F + 9 * p, Fn, sizeof Fn);
    memcpy(x + 3 * p, pos, sizeof pos);
    memcpy(v + 3 * p, vel, sizeof vel);
    memcpy(C + 9 * p, Cn, sizeof Cn);
  }
}

/* G2P2GTransfer::operator() (simulation/transfer/G2P2G.hpp:49-141): the matrix-free force evaluation of the implicit solver —
 * gather C from a grid DOF vector gridv (3 floats per node, node = blockno * 64 + cellid, :62-76), F_trial = (I + dt C) F
 * (:100-103, not stored), stress of the trial state (:104-121), scatter W * (contrib * D_inv) * xixp into the DOF vector gridr
 * (:123-138).  The reference's dof_view types (types/View.h) do not compile under gcc 13 here, so the HOST build of the reference
 * cannot run this functor; it is pinned on the GPU instead (round 2): oracle/ref_driver_cuda.cu instantiates the unmodified
 * G2P2GTransfer with a plain three-floats-per-node DOF view and tests/test_gpu_models.py compares this restatement with its output
 * (fixed-corotated 2.6e-5, von Mises 2.4e-5 of max |r|: the distance between the reference's host and device arithmetic).
 * model 0 fixed-corotated, 1 von Mises {yield}, 2 Drucker-Prager {cohesion, beta, yieldSurface, volumeCorrection},
 * 3 NACC {xi, beta, hardeningOn, fa, dim}, 4 equation of state {bulk, viscosity} (Jp instead of F). */
void zo_g2p2g(int model, const float *prm, int n, const float *x, const float *F, const float *Jp, const float *logJp,
              float dx, float dt, float E, float nu, float volume, int table_size, const int *keys, const int *indices,
              const float *gridv, float *gridr) {
  const float dx_inv = (float)1 / dx;
  const float D_inv = 4.f * dx_inv * dx_inv;
  float mu, lam;
  zo_lame(E, nu, &mu, &lam);
  for (int p = 0; p < n; ++p) {
    float pos[3] = {x[3*p], x[3*p+1], x[3*p+2]}, Cn[9] = {0}, contrib[9];
    zo_arena ar;
    arena_init(&ar, dx, pos);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];
      const int bno = zo_table_query(blk, table_size, keys, indices);
      const int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];
      const float *vi = gridv + ((size_t)bno * 64 + cell) * 3;
      for (int d = 0; d < 9; ++d) Cn[d] += W * vi[d % 3] * xixp[d / 3] * D_inv;               /* :75-76 */
    }
    if (model == 4) {                                                                          /* :80-98 */
      float J = Jp[p];
      J = (1 + (Cn[0] + Cn[4] + Cn[8]) * dt) * J;
      const float vol = volume * J;
      float pressure = prm[0];
      { float J2 = J * J; float J4 = J2 * J2; pressure = pressure * (1 / (J * J2 * J4) - 1); }
      const float visc = prm[1];
      contrib[0] = ((Cn[0] + Cn[0]) * visc - pressure) * vol;
      contrib[1] = (Cn[1] + Cn[3]) * visc * vol;
      contrib[2] = (Cn[2] + Cn[6]) * visc * vol;
      contrib[3] = (Cn[3] + Cn[1]) * visc * vol;
      contrib[4] = ((Cn[4] + Cn[4]) * visc - pressure) * vol;
      contrib[5] = (Cn[5] + Cn[7]) * visc * vol;
      contrib[6] = (Cn[6] + Cn[2]) * visc * vol;
      contrib[7] = (Cn[7] + Cn[5]) * visc * vol;
      contrib[8] = ((Cn[8] + Cn[8]) * visc - pressure) * vol;
    } else {
      float tmp[9], Fn[9];
      const float *Fo = F + 9 * p;
      for (int d = 0; d < 9; ++d) tmp[d] = Cn[d] * dt + ((d & 0x3) ? 0.f : 1.f);
      for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        Fn[3*c + r] = (tmp[r] * Fo[3*c] + tmp[3 + r] * Fo[3*c + 1]) + tmp[6 + r] * Fo[3*c + 2];
      float lj = logJp ? logJp[p] : 0.f;                                                       /* a local copy: never written back */
      if (model == 0) zo_stress_fixedcorotated(volume, mu, lam, Fn, contrib);
      else if (model == 1) zo_stress_vonmises(volume, mu, lam, prm[0], Fn, contrib);
      else if (model == 2) zo_stress_sand(volume, mu, lam, prm[0], prm[1], prm[2], prm[3] != 0.f, &lj, Fn, contrib);
      else zo_stress_nacc(volume, mu, lam, zo_nacc_bulk(E, nu), prm[0], prm[1], zo_nacc_msqr(prm[3], (int)prm[4]), prm[2] != 0.f, &lj, Fn, contrib);
    }
    for (int d = 0; d < 9; ++d) contrib[d] = contrib[d] * D_inv;                               /* :122 */
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];
      const int bno = zo_table_query(blk, table_size, keys, indices);
      const int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];
      float *r = gridr + ((size_t)bno * 64 + cell) * 3;
      for (int d = 0; d < 3; ++d)
        r[d] += W * (contrib[d] * xixp[0] + contrib[3 + d] * xixp[1] + contrib[6 + d] * xixp[2]);   /* :134-137 */
    }
  }
}

/* EquationOfStateConfig branch of G2PTransfer, G2P.hpp:69-73: J <- (1 + tr(C) dt) J, F untouched */
void zo_g2p_eos(int n, float *x, float *v, float *C, float *Jp, float dx, float dt, int table_size,
                const int *keys, const int *indices, const float *grid) {
  const float dx_inv = (float)1 / dx;
  const float D_inv = 4.f * dx_inv * dx_inv;
  for (int p = 0; p < n; ++p) {
    float pos[3] = {x[3*p], x[3*p+1], x[3*p+2]}, vel[3] = {0, 0, 0}, Cn[9] = {0};
    zo_arena ar;
    arena_init(&ar, dx, pos);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) {
      int o[3] = {i, j, k}, loc[3], blk[3];
      float xixp[3], vi[3];
      for (int d = 0; d < 3; ++d) {
        blk[d] = floor_div4(ar.corner[d] + o[d], &loc[d]);
        xixp[d] = (float)o[d] * ar.dx - ar.local[d];
      }
      float W = 1.f; W *= ar.w[0][i]; W *= ar.w[1][j]; W *= ar.w[2][k];
      int bno = zo_table_query(blk, table_size, keys, indices);
      const float *tile = grid + (size_t)bno * 7 * 64;
      int cell = (loc[0] << 4) | (loc[1] << 2) | loc[2];
      for (int d = 0; d < 3; ++d) { vi[d] = tile[(1 + d) * 64 + cell]; vel[d] += vi[d] * W; }
      for (int d = 0; d < 9; ++d) Cn[d] += W * vi[d % 3] * xixp[d / 3] * D_inv;
    }
    for (int d = 0; d < 3; ++d) pos[d] += vel[d] * dt;
    Jp[p] = (1 + (Cn[0] + Cn[4] + Cn[8]) * dt) * Jp[p];   /* :71 */
    memcpy(x + 3 * p, pos, sizeof pos);
    memcpy(v + 3 * p, vel, sizeof vel);
    memcpy(C + 9 * p, Cn, sizeof Cn);
  }
}

/* ApplyBoundaryConditionOnGridBlocks (GridOp.hpp:112-164) with a STATIC analytic collider (Collider.h:98-127 with the
 * default R = I, s = 1, b = dbdt = omega = 0, so v_object = 0): geom 0 = Plane{origin p0, normal p1}
 * (AnalyticLevelSet.h:11-43), geom 1 = Sphere{centre p0, radius p1[0]} (:130-157); type = collider_e
 * {0 Sticky, 1 Slip, 2 Separate} (Collider.h:8). */
/* AnalyticLevelSet<Cuboid>::do_getSignedDistance (AnalyticLevelSet.h:89-96): box [mn, mx] */
static float cuboid_sdf(const float x[3], const float mn[3], const float mx[3]) {
  float point[3];
  for (int i = 0; i < 3; ++i) {
    const float center = (mn[i] + mx[i]) / 2;
    const float a = x[i] - center;
    point[i] = (a > 0 ? a : -a) - (mx[i] - mn[i]) / 2;
  }
  float max = point[0];
  for (int i = 1; i < 3; ++i) if (point[i] > max) max = point[i];
  for (int i = 0; i < 3; ++i) if (point[i] < 0) point[i] = 0;
  float l2 = 0.f;
  for (int i = 0; i < 3; ++i) l2 += point[i] * point[i];
  return (max < 0 ? max : 0) + sqrtf(l2);
}
/* ::do_getNormal (:98-110): central differences of the signed distance with eps = 1e-6 IN FLOAT, then normalized() */
static void cuboid_normal(const float x[3], const float mn[3], const float mx[3], float nm[3]) {
  const float eps = (float)1e-6;
  float diff[3];
  for (int i = 0; i < 3; ++i) {
    float v1[3] = {x[0], x[1], x[2]}, v2[3] = {x[0], x[1], x[2]};
    v1[i] = x[i] + eps;
    v2[i] = x[i] - eps;
    diff[i] = (cuboid_sdf(v1, mn, mx) - cuboid_sdf(v2, mn, mx)) / (eps + eps);
  }
  float l2 = 0.f;
  for (int i = 0; i < 3; ++i) l2 += diff[i] * diff[i];
  const float len = sqrtf(l2);
  for (int i = 0; i < 3; ++i) nm[i] = diff[i] / len;
}

void zo_cuboid(int n, const float *x, const float mn[3], const float mx[3], float *sdf, float *normal) {
  for (int p = 0; p < n; ++p) {
    sdf[p] = cuboid_sdf(x + 3 * p, mn, mx);
    cuboid_normal(x + 3 * p, mn, mx, normal + 3 * p);
  }
}

/* motion (may be NULL = the default rigid motion): b[3], dbdt[3], R[9] row-major, omega[3], s, dsdt — Collider.h:16-24,136-143;
 * geom 2 = Cuboid{min p0, max p1} (AnalyticLevelSet.h:55-126) */
void zo_apply_boundary_moving(int nblocks, const int *active_keys, float *grid, float dx, int geom, int type,
                              const float p0[3], const float p1[3], const float *motion) {
  float bb[3] = {0, 0, 0}, dbdt[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, om[3] = {0, 0, 0}, sc = 1.f, dsdt = 0.f;
  if (motion) {
    for (int k = 0; k < 3; ++k) { bb[k] = motion[k]; dbdt[k] = motion[3 + k]; om[k] = motion[15 + k]; }
    for (int k = 0; k < 9; ++k) R[k] = motion[6 + k];
    sc = motion[18]; dsdt = motion[19];
  }
  for (int b = 0; b < nblocks; ++b) {
    float *tile = grid + (size_t)b * 7 * 64;
    for (int c = 0; c < 64; ++c) {
      if (!(tile[c] > 0)) continue;                                    /* GridOp.hpp:141 */
      const int cc[3] = {(c >> 4) & 3, (c >> 2) & 3, c & 3};           /* cellid_to_coord, Structure.hpp:836-847 */
      float pos[3], vel[3], xmb[3], X[3], nm[3], n[3], d[3], vobj[3];
      for (int k = 0; k < 3; ++k) {
        pos[k] = ((float)active_keys[3 * b + k] * 4.f + (float)cc[k]) * dx;   /* :143-144 */
        vel[k] = tile[(1 + k) * 64 + c];
        xmb[k] = pos[k] - bb[k];                                        /* Collider.h:106 */
      }
      const float one_over_s = 1 / sc;
      for (int i = 0; i < 3; ++i) {                                     /* X = R^T (x - b) * (1/s), :108 */
        float sum = 0.f;
        for (int j = 0; j < 3; ++j) sum += R[3 * j + i] * xmb[j];
        X[i] = sum * one_over_s;
      }
      for (int k = 0; k < 3; ++k) d[k] = X[k] - p0[k];
      float dist;
      if (geom == 0) {
        dist = 0.f;
        for (int k = 0; k < 3; ++k) dist += p1[k] * d[k];              /* _normal.dot(x - _origin) */
        for (int k = 0; k < 3; ++k) nm[k] = p1[k];
      } else if (geom == 2) {
        dist = cuboid_sdf(X, p0, p1);
        if (dist < 0.f && type != 0) cuboid_normal(X, p0, p1, nm);     /* only evaluated where it is used (:116) */
        else nm[0] = nm[1] = nm[2] = 0.f;
      } else {
        float l2 = 0.f;
        for (int k = 0; k < 3; ++k) l2 += d[k] * d[k];
        const float len = sqrtf(l2);
        dist = len - p1[0];
        for (int k = 0; k < 3; ++k) nm[k] = l2 < 1e-7f ? 0.f : d[k] / len;   /* AnalyticLevelSet.h:148-152 */
      }
      if (dist < 0.f) {                                                 /* Collider.h:109 (erosion 0) */
        /* v_object = omega x (x-b) + (dsdt/s)(x-b) + R s V_material(= 0) + dbdt, :110-111 */
        const float cr[3] = {om[1] * xmb[2] - om[2] * xmb[1], om[2] * xmb[0] - om[0] * xmb[2], om[0] * xmb[1] - om[1] * xmb[0]};
        for (int i = 0; i < 3; ++i) {
          float rs = 0.f;
          for (int j = 0; j < 3; ++j) rs += (R[3 * i + j] * sc) * 0.f;
          vobj[i] = ((cr[i] + (dsdt * one_over_s) * xmb[i]) + rs) + dbdt[i];
        }
        if (type == 0) {
          for (int k = 0; k < 3; ++k) vel[k] = vobj[k];                 /* Sticky */
        } else {
          for (int k = 0; k < 3; ++k) vel[k] -= vobj[k];
          for (int i = 0; i < 3; ++i) {                                 /* n = R * normal(X), :116 */
            float sum = 0.f;
            for (int j = 0; j < 3; ++j) sum += R[3 * i + j] * nm[j];
            n[i] = sum;
          }
          float proj = 0.f;
          for (int k = 0; k < 3; ++k) proj += n[k] * vel[k];
          if ((type == 2 && proj < 0.f) || type == 1)
            for (int k = 0; k < 3; ++k) vel[k] -= proj * n[k];
          for (int k = 0; k < 3; ++k) vel[k] += vobj[k];
        }
      }
      for (int k = 0; k < 3; ++k) tile[(1 + k) * 64 + c] = vel[k];      /* block.set(1, cellid, vel) */
    }
  }
}
void zo_apply_boundary(int nblocks, const int *active_keys, float *grid, float dx, int geom, int type,
                       const float p0[3], const float p1[3]) {
  zo_apply_boundary_moving(nblocks, active_keys, grid, dx, geom, type, p0, p1, 0);
}

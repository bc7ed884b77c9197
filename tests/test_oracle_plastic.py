"""DruckerPragerConfig / NACCConfig: P2G.hpp:92-102 -> compute_stress_sand / compute_stress_nacc
(physics/ConstitutiveModel_Vol_dP.hpp:116-326).  CPU only:
  * the C restatement (oracle/mpm_oracle.c) is bit-exact against the reference itself (oracle/_ref) in every branch of
    both return mappings, for the stress functions and for the whole P2G functor including the logJp write-back;
  * the reference-generated golden vectors pin the oracle where /root/reference is absent;
  * the product's device math (zpc_b200/csrc/mpm_math.cuh) compiled for the host (tests/hostmath) tracks the oracle as
    closely as the oracle tracks itself across builds (FMA contraction on / off) — the bound the GPU parity rule uses.
"""
import ast
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests.golden.make_golden import NACC, SAND, nacc_margin, sand_margin
from zpc_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
E, NU = synth.MODEL["E"], synth.MODEL["nu"]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _isochoric(rs, n, amp):
    """deformation gradients with det F = 1 up to rounding (p_trial inside NACC's O(1) Pa wide yield surface)"""
    out = np.empty((n, 9), np.float32)
    for i in range(n):
        q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
        a, b = np.exp(rs.uniform(-amp, amp, 2))
        Fm = q @ np.diag([a, b, 1.0 / (a * b)]) @ q.T
        out[i] = Fm.T.reshape(9)          # column-major 9-vector
    return out


def test_math_sqrt_restatement_is_bit_exact(oracle, ref):
    rs = np.random.RandomState(1)
    xs = np.abs(rs.standard_normal(5000)) * 10.0 ** rs.randint(-30, 30, 5000)
    xs = np.concatenate([xs, [0.0, 1.0, 2.0, 4.0, 1e-40, 3.4e38, -1.0]]).astype(np.float32)
    for x in xs:
        a, b = oracle.math_sqrt(float(x)), ref.math_sqrt(float(x))
        assert (np.isnan(a) and np.isnan(b)) or a == b, x
    assert oracle.nacc_consts(E, NU, 45.0) == ref.nacc_consts(E, NU, 45.0)
    assert oracle.nacc_consts(NACC["E"], NACC["nu"], NACC["fa"]) == ref.nacc_consts(NACC["E"], NACC["nu"], NACC["fa"])


def test_sand_stress_bit_exact_vs_reference_in_every_branch(oracle, ref):
    rs = np.random.RandomState(5)
    seen = {"tip": 0, "inside": 0, "surface": 0}
    for sand in (SAND, dict(SAND, cohesion=0.02, beta=0.5, volumeCorrection=False)):
        for eps in (1e-3, 1e-2, 0.05, 0.2):
            for _ in range(150):
                F = (np.eye(3) + eps * rs.standard_normal((3, 3))).astype(np.float32).reshape(9)
                lj = float(np.float32(rs.uniform(-0.05, 0.02)))
                a, la = ref.stress_sand(1e-6, E, NU, sand, lj, F)
                b, lb = oracle.stress_sand(1e-6, E, NU, sand, lj, F)
                assert np.array_equal(_bits(a), _bits(b)) and _bits(la) == _bits(lb), (sand, F, lj)
                _, fr = sand_margin(dict(F=F[None], logJp=np.float32([lj])), E, NU, sand)
                seen["tip"] += fr[0] > 0; seen["inside"] += fr[1] > 0; seen["surface"] += fr[2] > 0
    assert min(seen.values()) > 100, seen


def test_nacc_stress_bit_exact_vs_reference_in_every_branch(oracle, ref):
    rs = np.random.RandomState(6)
    seen = np.zeros(5)
    for nacc in (NACC, dict(NACC, hardeningOn=False, beta=0.3, xi=1.1), dict(NACC, E=E, nu=NU)):
        Fs = np.concatenate([(np.eye(3).reshape(9)[None] + eps * rs.standard_normal((120, 9))).astype(np.float32)
                             for eps in (1e-3, 1e-2, 0.05)] + [_isochoric(rs, 200, 0.05), _isochoric(rs, 200, 0.3)])
        ljs = rs.uniform(-1.8, 0.2, Fs.shape[0]).astype(np.float32)
        for F, lj in zip(Fs, ljs):
            a, la = ref.stress_nacc(1e-6, nacc["E"], nacc["nu"], nacc, float(lj), F)
            b, lb = oracle.stress_nacc(1e-6, nacc["E"], nacc["nu"], nacc, float(lj), F)
            assert np.array_equal(_bits(a), _bits(b)) and _bits(la) == _bits(lb), (nacc, F, lj)
        _, fr = nacc_margin(dict(F=Fs, logJp=ljs), nacc)
        seen += np.array(fr) * Fs.shape[0]
    assert seen.min() > 30, seen          # max tip, min tip, projection onto the surface, hardening solve, inside


@pytest.mark.parametrize("model", ["sand", "nacc"])
def test_plastic_p2g_bit_exact_vs_reference(oracle, ref, model):
    """the whole functor on the reference's seq_exec policy: grid and the logJp it writes back (P2G.hpp:101)"""
    P = synth.elastic_cube(6, 16, jitter_F=0.06 if model == "sand" else 0.03, jitter_C=0.5, shuffle_seed=3)
    n, dx = P["x"].shape[0], P["dx"]
    lo, hi = (-0.06, 0.03) if model == "sand" else (-1.8, 0.2)
    P["logJp"] = np.random.RandomState(9).uniform(lo, hi, n).astype(np.float32)
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.set_logJp(P["logJp"])
    h.partition()
    tab = h.table()
    h.clean_grid()
    if model == "sand":
        h.p2g_sand(synth.DT, E, NU, SAND, P["volume"])
    else:
        h.p2g_nacc(synth.DT, NACC["E"], NACC["nu"], NACC, P["volume"])
    g_ref, lj_ref = h.grid(), h.get_logJp()
    h.close()
    Po = dict(P, logJp=P["logJp"].copy())
    if model == "sand":
        g = oracle.p2g_sand(Po, tab, dx, synth.DT, E, NU, SAND, P["volume"])
    else:
        g = oracle.p2g_nacc(Po, tab, dx, synth.DT, NACC["E"], NACC["nu"], NACC, P["volume"])
    assert np.array_equal(g, g_ref)
    assert np.array_equal(_bits(Po["logJp"]), _bits(lj_ref))
    assert not np.array_equal(Po["logJp"], P["logJp"])


@pytest.mark.parametrize("model", ["sand", "nacc"])
def test_plastic_golden_pins_the_oracle(oracle, model):
    """reference-generated vectors (tests/golden/make_golden.py), by block key; needs no reference at run time"""
    z = np.load(os.path.join(G, "mpm_cube6_%s.npz" % model))
    P = synth.elastic_cube(int(z["s"]), int(z["G"]), **dict(ast.literal_eval(str(z["kw"]))))
    n, dx = P["x"].shape[0], P["dx"]
    P["logJp"] = z["logJp_in"].copy()
    assert z["branch_fractions"][:3].min() > 0.05        # every branch of the return mapping is populated
    tab = oracle.partition_build(P["x"], dx, oracle.table_size_for(max(n // 8, 1)))
    if model == "sand":
        g = oracle.p2g_sand(P, tab, dx, synth.DT, E, NU, SAND, P["volume"])
    else:
        g = oracle.p2g_nacc(P, tab, dx, synth.DT, NACC["E"], NACC["nu"], NACC, P["volume"])
    ko, kr = tab["active_keys"], z["active_keys"]
    o, r = np.lexsort((ko[:, 2], ko[:, 1], ko[:, 0])), np.lexsort((kr[:, 2], kr[:, 1], kr[:, 0]))
    assert np.array_equal(ko[o], kr[r])
    assert np.array_equal(g[o], z["grid_p2g"][r])
    assert np.array_equal(_bits(P["logJp"]), _bits(z["logJp"]))
    mx = oracle.grid_update(g, synth.DT, (0.0, synth.GRAVITY, 0.0), 1)
    assert mx == float(z["max_vel_sqr"])
    oracle.g2p(P, tab, g, dx, synth.DT)
    for k in "xvCF":
        assert np.array_equal(P[k], z[k]), k


# ---- the product's device math, compiled for the host ----------------------------------------------------------------
@pytest.fixture(scope="module")
def hostmath():
    from tests.hostmath import build_hostmath
    return C.CDLL(build_hostmath())


@pytest.fixture(scope="module")
def oracle_fma():
    """the oracle built WITH FMA contraction: how far the reference's arithmetic moves between two builds of one source"""
    from oracle.pyoracle import Oracle
    d = tempfile.mkdtemp()
    so = os.path.join(d, "liboracle_fma.so")
    try:
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=fast", "-mfma", "-o", so] +
                              [os.path.join(ROOT, "oracle", f) for f in ("mpm_oracle.c", "prims_oracle.c", "sparse_oracle.c")] +
                              ["-lm", "-fopenmp"])
        o2 = Oracle.__new__(Oracle)
        o2.lib = C.CDLL(so)
        o2.stress_fixedcorotated(1.0, E, NU, np.eye(3, dtype=np.float32).reshape(9))
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("no FMA build of the oracle on this host")
    return o2


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _oracle_stress(o, model, F, lj, prm):
    n = F.shape[0]
    out, l = np.empty((n, 9), np.float32), lj.copy()
    for i in range(n):
        if model == 0:
            out[i] = o.stress_fixedcorotated(1.0, prm["E"], prm["nu"], F[i])
        elif model == 1:
            out[i] = o.stress_vonmises(1.0, prm["E"], prm["nu"], prm["ys"], F[i])
        elif model == 2:
            out[i], l[i] = o.stress_sand(1.0, prm["E"], prm["nu"], prm, float(lj[i]), F[i])
        else:
            out[i], l[i] = o.stress_nacc(1.0, prm["E"], prm["nu"], prm, float(lj[i]), F[i])
    return out, l


@pytest.mark.parametrize("model", [0, 10, 11, 1, 2, 3])
def test_device_math_on_the_host_tracks_the_oracle(oracle, oracle_fma, hostmath, model):
    """zpcm::stress_* (what the kernels run) vs the oracle on random F.  The device code orders the SVD's arithmetic
    differently, so it is held to the distance between two builds of the reference's own arithmetic (+ an fp32 floor)
    — measured here, per regime, against the stress scale of the sample."""
    rs = np.random.RandomState(40 + model)
    prm = {0: dict(E=E, nu=NU), 10: dict(E=E, nu=NU), 11: dict(E=E, nu=NU), 1: dict(E=E, nu=NU, ys=300.0), 2: dict(SAND, E=E, nu=NU), 3: dict(NACC)}[model]
    mu, lam = oracle.lame(prm["E"], prm["nu"])
    if model == 3:
        bm, msqr = oracle.nacc_consts(prm["E"], prm["nu"], prm["fa"])
        b, m = C.c_float(), C.c_float()
        hostmath.hm_nacc_consts(C.c_float(prm["E"]), C.c_float(prm["nu"]), C.c_float(prm["fa"]), C.c_int(3), C.byref(b), C.byref(m))
        assert (b.value, m.value) == (bm, msqr)
    vec = {0: [0.0], 10: [0.0], 11: [0.0], 1: [300.0], 2: [SAND["cohesion"], SAND["beta"], SAND["yieldSurface"], 1.0],
           3: [0, NACC["xi"], NACC["beta"], 0, 1.0]}[model]
    if model == 3:
        vec[0], vec[3] = bm, msqr
    vec = np.array(vec, np.float32)
    for eps in (1e-2, 0.05, 0.2):
        n = 1500
        F = (np.eye(3).reshape(9)[None] + eps * rs.standard_normal((n, 9))).astype(np.float32)
        if model == 3:
            F[: n // 2] = _isochoric(rs, n // 2, eps)
        lj = rs.uniform(-1.8, 0.2, n).astype(np.float32) if model == 3 else rs.uniform(-0.05, 0.02, n).astype(np.float32)
        got, l = np.empty((n, 9), np.float32), lj.copy()
        hostmath.hm_stress(C.c_int(model), C.c_int(n), C.c_float(1.0), C.c_float(mu), C.c_float(lam), _p(vec), _p(l), _p(F), _p(got))
        want, lw = _oracle_stress(oracle, 0 if model >= 10 else model, F, lj, prm)       # 10 = stress_fcr_lean: the fixed-corotated stress as the binned P2G evaluates it, 11 = with the converged-sweep skip
        other, lo = _oracle_stress(oracle_fma, 0 if model >= 10 else model, F, lj, prm)
        # particles whose branch differs between the two reference builds sit on a branch boundary: not comparable
        stable = np.abs(lo - lw) <= 1e-5 * np.maximum(np.abs(lw), 1.0)
        assert stable.mean() > 0.98
        scale = float(np.abs(want[stable]).max())
        dev = np.abs(got - want)[stable].max(1) / scale
        ref_dev = np.abs(other - want)[stable].max(1) / scale
        # the bulk of the distribution (99th percentile) within 2x, the heavy tail (particles next to a yield surface,
        # where the projection is ill-conditioned) within 10x of the reference's own build-to-build distance
        assert np.percentile(dev, 99) <= 2.0 * np.percentile(ref_dev, 99) + 2e-6, (model, eps, np.percentile(dev, 99), np.percentile(ref_dev, 99))
        assert dev.max() <= 10.0 * ref_dev.max() + 2e-6, (model, eps, dev.max(), ref_dev.max())
        assert np.abs(l - lw)[stable].max() <= 1e-5


def _rot(rs):
    q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _hm_fcr(hostmath, model, mu, lam, F):
    n = F.shape[0]
    got, vec, l = np.empty((n, 9), np.float32), np.zeros(1, np.float32), np.zeros(n, np.float32)
    hostmath.hm_stress(C.c_int(model), C.c_int(n), C.c_float(1.0), C.c_float(mu), C.c_float(lam), _p(vec), _p(l), _p(F), _p(got))
    return got


def test_fast_path_stress_far_from_the_identity(oracle, hostmath):
    """stress_fcr_lean (what the binned P2G evaluates; 11 = with the converged-sweep skip) against the oracle where the golden clouds
    do not go: large rotations, stretches of 0.5 .. 2, a general U S V^T, inverted elements.  Held to the restated full SVD path
    (model 0 = the reference's algorithm operation by operation) on the same inputs: the bulk within 2x of its distance to the
    oracle.  Inverted elements whose two smaller |sigma| nearly coincide are ill-conditioned in the model itself (which singular
    value carries the sign) — there the reference's own arithmetic is 1e-3 from the exact answer — so only the bulk is compared."""
    rs = np.random.RandomState(5)
    mu, lam = oracle.lame(E, NU)
    n = 1500
    regimes = {
        "rotation x small stretch": [_rot(rs) @ (np.eye(3) + 0.03 * rs.standard_normal((3, 3))) for _ in range(n)],
        "rotation x stretch": [(lambda q: q @ np.diag(np.exp(rs.uniform(-0.7, 0.7, 3))) @ q.T)(_rot(rs)) for _ in range(n)],
        "general U S V^T": [_rot(rs) @ np.diag(np.exp(rs.uniform(-0.7, 0.7, 3))) @ _rot(rs) for _ in range(n)],
        "inverted": [_rot(rs) @ np.diag(np.exp(rs.uniform(-0.3, 0.3, 3)) * np.array([1, 1, -1.0])) @ _rot(rs) for _ in range(n)],
        "two equal sigmas": [_rot(rs) @ np.diag([1.2, 1.2, 0.7]) @ _rot(rs) for _ in range(n)],
    }
    for name, Fs in regimes.items():
        F = np.array([f.T.reshape(9) for f in Fs], np.float32)
        want = np.array([oracle.stress_fixedcorotated(1.0, E, NU, F[i]) for i in range(n)])
        scale = float(np.abs(want).max())
        full = np.abs(_hm_fcr(hostmath, 0, mu, lam, F) - want).max(1) / scale
        for model in (10, 11):
            dev = np.abs(_hm_fcr(hostmath, model, mu, lam, F) - want).max(1) / scale
            assert np.percentile(dev, 99) <= 2.0 * np.percentile(full, 99) + 2e-6, (name, model, np.percentile(dev, 99), np.percentile(full, 99))
            if name != "inverted":
                assert dev.max() <= 3.0 * full.max() + 5e-6, (name, model, dev.max(), full.max())


def test_fast_path_stress_stays_finite_on_degenerate_F(oracle, hostmath):
    """exactly rank-deficient, zero and vanishing F (a collapsed particle): the reference's guarded Givens steps return finite numbers,
    and so must the Gram-Schmidt form of the fast path — a NaN in one record would spread over the grid.  (The floor under the squared
    column norms was 1e-30 at first: rsqrt_refined's Newton step overflows below ~1e-26 and turned these inputs into NaN.)"""
    rs = np.random.RandomState(3)
    mu, lam = oracle.lame(E, NU)
    Fs = [_rot(rs) @ np.diag([1.0, 0.8, 0.0]) @ _rot(rs) for _ in range(100)] + [_rot(rs) @ np.diag([1.0, 0.0, 0.0]) @ _rot(rs) for _ in range(100)]
    Fs += [np.diag(d) for d in ([1.0, 0.8, 0.0], [0.0, 1.0, 0.8], [1.0, 0.0, 0.8], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0])]
    Fs += [1e-20 * np.eye(3), 1e-12 * _rot(rs), 1e-6 * _rot(rs)]
    F = np.array([np.asarray(f, np.float64).T.reshape(9) for f in Fs], np.float32)
    want = np.array([oracle.stress_fixedcorotated(1.0, E, NU, F[i]) for i in range(F.shape[0])])
    assert np.isfinite(want).all()
    for model in (0, 10, 11):
        got = _hm_fcr(hostmath, model, mu, lam, F)
        assert np.isfinite(got).all(), (model, np.nonzero(~np.isfinite(got).all(1))[0])
        assert np.abs(got - want).max() <= 1e-4 * 2.0 * mu, (model, np.abs(got - want).max())


def test_device_stencil_on_the_host_is_bit_exact(oracle, hostmath):
    """zpcm::arena_init == LocalArena (simulation/Utils.hpp:51-70): base node, local offset, 3x3 weights — same
    expressions, no re-association, so the host build must agree bit for bit with an independent numpy restatement"""
    rs = np.random.RandomState(2)
    n, dx = 4000, np.float32(1.0 / 64)
    x = rs.uniform(-0.3, 1.3, (n, 3)).astype(np.float32)
    corner, local, w = np.empty((n, 3), np.int32), np.empty((n, 3), np.float32), np.empty((n, 9), np.float32)
    hostmath.hm_arena(C.c_int(n), C.c_float(dx), _p(x), _p(corner), _p(local), _p(w))
    X = (x / dx).astype(np.float32)
    cn = np.floor(X - np.float32(0.5)).astype(np.int32)
    lp = (X - cn.astype(np.float32)).astype(np.float32)
    d0 = (lp - np.floor(lp - np.float32(0.5))).astype(np.float32)
    w0 = (np.float32(0.5) * (np.float32(1.5) - d0) * (np.float32(1.5) - d0)).astype(np.float32)
    d1 = (d0 - np.float32(1.0)).astype(np.float32)
    w1 = (np.float32(0.75) - d1 * d1).astype(np.float32)
    zz = (np.float32(0.5) + d1).astype(np.float32)
    w2 = (np.float32(0.5) * zz * zz).astype(np.float32)
    assert np.array_equal(corner, cn)
    assert np.array_equal(_bits(local), _bits((lp * dx).astype(np.float32)))
    assert np.array_equal(_bits(w.reshape(n, 3, 3)), _bits(np.stack([w0, w1, w2], axis=2)))


@pytest.mark.parametrize("dx", [1.0 / 64, 1.0 / 256, 0.01, 0.0371, 1.0 / 3])
def test_exact_division_stencil_is_the_division_stencil(hostmath, dx):
    """arena_init<true> (the binned G2P: x / dx as 1/dx and two FMAs, no slow-path branch) against arena_init (the IEEE division the
    reference writes): corner, local offset and the nine weights bit for bit — positions near the origin, far from it (6 000 cells
    out, where an ulp is 5e-4 of a cell), negative, and exactly on cell boundaries, for power-of-two and other cell sizes"""
    rs = np.random.RandomState(11)
    dx = np.float32(dx)
    parts = [rs.uniform(-0.3, 1.3, (200000, 3)), rs.uniform(-200.0, 200.0, (200000, 3)), rs.uniform(5999, 6001, (100000, 3)) * float(dx),
             (rs.randint(-5000, 5000, (100000, 3)) + rs.choice([0.0, 0.5], (100000, 3))) * float(dx)]
    x = np.concatenate(parts).astype(np.float32)
    n = x.shape[0]
    out = []
    for fn in (hostmath.hm_arena, hostmath.hm_arena_fastdiv):
        corner, local, w = np.empty((n, 3), np.int32), np.empty((n, 3), np.float32), np.empty((n, 9), np.float32)
        fn(C.c_int(n), C.c_float(dx), _p(x), _p(corner), _p(local), _p(w))
        out.append((corner, local, w))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(_bits(out[0][1]), _bits(out[1][1]))
    assert np.array_equal(_bits(out[0][2]), _bits(out[1][2]))


"""Generates tests/golden/*.npz by executing the REFERENCE itself (oracle/_ref/libzpcref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_golden.py) and, through it, the CUDA path on the GPU box, where
neither /root/reference nor (necessarily) oracle/_ref exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import Ref  # noqa: E402
from zpc_b200 import synth  # noqa: E402


def mpm_case(ref, name, s, G, mode, **kw):
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    nb = h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    g1 = h.grid()
    mx = h.grid_update(synth.DT, synth.GRAVITY, mode)
    g2 = h.grid()
    h.g2p(synth.DT)
    out = h.get_particles()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, mode=mode, kw=repr(sorted(kw.items())),
                        nblocks=nb, active_keys=tab["active_keys"], table_keys=tab["keys"],
                        table_indices=tab["indices"], grid_p2g=g1, grid_upd=g2, max_vel_sqr=mx,
                        x=out["x"], v=out["v"], C=out["C"], F=out["F"])
    print(name, "n", n, "blocks", nb, "maxvel2", mx)


BOUNDARY_CASES = [(0, 0, (0.0, 0.33, 0.0), (0.0, 1.0, 0.0)), (0, 1, (0.3, 0.3, 0.3), (0.6, 0.8, 0.0)),
                  (0, 2, (0.0, 0.33, 0.0), (0.0, 1.0, 0.0)), (1, 0, (0.33, 0.3, 0.33), (0.08, 0.0, 0.0)),
                  (1, 1, (0.33, 0.3, 0.33), (0.08, 0.0, 0.0)), (1, 2, (0.33, 0.3, 0.33), (0.08, 0.0, 0.0))]


def boundary_case(ref, name, s, G, **kw):
    """ApplyBoundaryConditionOnGridBlocks after P2G + update, every (geometry, collider type) of BOUNDARY_CASES"""
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    out = {}
    for i, (geom, ctype, p0, p1) in enumerate(BOUNDARY_CASES):
        h = ref.mpm(n, dx, 0)
        h.set_particles(P)
        h.partition()
        tab = h.table()
        h.clean_grid()
        h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
        h.grid_update(synth.DT, synth.GRAVITY, 1)
        h.apply_boundary(geom, ctype, p0, p1)
        out["grid_%d" % i] = h.grid()
        out["active_keys"] = tab["active_keys"]
        h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), **out)
    print(name, "n", n, len(BOUNDARY_CASES), "colliders")


def boundary_moving_case(ref, name, s, G, **kw):
    """ApplyBoundaryConditionOnGridBlocks with moving colliders (translation, rotation, angular velocity, scaling:
    geometry/Collider.h:16-24,98-127), every case of tests/parity.py MOVING_COLLIDERS"""
    from tests.parity import MOVING_COLLIDERS, motion_vec
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    out = {}
    for i, (geom, ctype, p0, p1, motion) in enumerate(MOVING_COLLIDERS):
        h = ref.mpm(n, dx, 0)
        h.set_particles(P)
        h.partition()
        tab = h.table()
        h.clean_grid()
        h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
        h.grid_update(synth.DT, synth.GRAVITY, 1)
        h.apply_boundary(geom, ctype, p0, p1, motion_vec(motion))
        out["grid_%d" % i] = h.grid()
        out["active_keys"] = tab["active_keys"]
        h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), **out)
    print(name, "n", n, len(MOVING_COLLIDERS), "moving colliders")


def boundary_cuboid_case(ref, name, s, G, **kw):
    """ApplyBoundaryConditionOnGridBlocks over AnalyticLevelSet<Cuboid> colliders, static and moving (tests/parity.py
    CUBOID_COLLIDERS): signed distance of a box, normal by central differences (AnalyticLevelSet.h:89-110)"""
    from tests.parity import CUBOID_COLLIDERS, motion_vec
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    out = {}
    for i, (geom, ctype, p0, p1, motion) in enumerate(CUBOID_COLLIDERS):
        h = ref.mpm(n, dx, 0)
        h.set_particles(P)
        h.partition()
        tab = h.table()
        h.clean_grid()
        h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
        h.grid_update(synth.DT, synth.GRAVITY, 1)
        h.apply_boundary(geom, ctype, p0, p1, motion_vec(motion))
        out["grid_%d" % i] = h.grid()
        out["active_keys"] = tab["active_keys"]
        h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), **out)
    print(name, "n", n, len(CUBOID_COLLIDERS), "cuboid colliders")


def grid_momentum_case(ref, name, s, G, **kw):
    """GridMomentumToVelocity and GridAngularMomentum (GridOp.hpp:184-262) under the sequential policy on the grid P2G leaves"""
    P = synth.elastic_cube(s, G, **kw)
    P["v"] = (P["v"] + np.float32([0.3, -0.2, 0.1])).astype(np.float32)
    n, dx = P["x"].shape[0], P["dx"]
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    grid = h.grid()
    sum6 = h.angular_momentum()
    mx = h.momentum_to_velocity()
    vel = h.grid()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), v_shift=np.float32([0.3, -0.2, 0.1]),
                        active_keys=tab["active_keys"], grid=grid, sum6=sum6, max_vel_sqr=np.float32(mx), vel=vel)
    print(name, "n", n, "blocks", grid.shape[0], "sum6", sum6)


def vonmises_margin(P, E, nu, ys):
    """smallest relative distance of any particle's trial deviatoric stress norm from the yield radius"""
    mu, lam = 0.5 * E / (1 + nu), E * nu / ((1 + nu) * (1 - 2 * nu))
    F = P["F"].reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)      # column-major 9-vectors
    sig = np.linalg.svd(F, compute_uv=False)
    J = sig.prod(1, keepdims=True)
    tau = 2 * mu * (sig - 1) * sig + lam * (J - 1) * J
    sn = np.linalg.norm(tau - tau.mean(1, keepdims=True), axis=1)
    r = np.sqrt(2.0 / 3.0) * ys
    return float(np.abs(sn - r).min() / r), float((sn > r).mean())


def vonmises_case(ref, name, s, G, ys, **kw):
    """VonMisesFixedCorotatedConfig{E, nu, yieldStress=ys} substep (P2G.hpp:89-90, ConstitutiveModel_Vol_dP.hpp:49-110)"""
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    margin, frac = vonmises_margin(P, synth.MODEL["E"], synth.MODEL["nu"], ys)
    assert margin > 1e-3, margin      # no particle sits on the yield surface: host / device rounding cannot flip a branch
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    nb = h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g_vonmises(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], ys, P["volume"])
    g1 = h.grid()
    mx = h.grid_update(synth.DT, synth.GRAVITY, 1)
    h.g2p(synth.DT)
    out = h.get_particles()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, ys=ys, kw=repr(sorted(kw.items())), nblocks=nb,
                        active_keys=tab["active_keys"], grid_p2g=g1, max_vel_sqr=mx, yielded_fraction=frac,
                        x=out["x"], v=out["v"], C=out["C"], F=out["F"])
    print(name, "n", n, "blocks", nb, "yielded fraction %.2f, margin %.2e" % (frac, margin))


SAND = dict(cohesion=0.0, beta=1.0, yieldSurface=0.816496580927726 * 2.0 * 0.5 / (3.0 - 0.5), volumeCorrection=True)
# NACC in a regime where all three branches of the return mapping occur and the stress is well conditioned: with the
# reference's p0 = bulk*1e-5 + sin(xi*max(-logJp,0)) (ConstitutiveModel_Vol_dP.hpp:123) the yield surface is O(1) Pa
# wide, so a soft material (E = 10) and logJp in [-1.8, 0.2] put particles inside it, beyond either tip and on it.
NACC = dict(fa=45.0, xi=0.8, beta=0.5, hardeningOn=True, E=10.0, nu=0.4)


def sand_margin(P, E, nu, sand):
    """branch statistics of compute_stress_sand in float64: fractions of (tip, inside, cone surface) and the smallest
    distance of any particle from a branch boundary"""
    mu, lam = 0.5 * E / (1 + nu), E * nu / ((1 + nu) * (1 - 2 * nu))
    F = P["F"].reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)
    sig = np.linalg.svd(F, compute_uv=False)
    eps = np.log(np.maximum(np.abs(sig), 1e-4)) - sand["cohesion"]
    tr = eps.sum(1) + P["logJp"].astype(np.float64)
    eh = eps - tr[:, None] / 3
    dg = np.linalg.norm(eh, axis=1) + (3 * lam + 2 * mu) / (2 * mu) * tr * sand["yieldSurface"]
    tip, inside = tr >= 0, (tr < 0) & (dg <= 0)
    margin = min(float(np.abs(tr).min()), float(np.abs(dg[~tip]).min(initial=np.inf)))
    return margin, (float(tip.mean()), float(inside.mean()), float((~tip & ~inside).mean()))


def sand_case(ref, name, s, G, **kw):
    """DruckerPragerConfig substep (P2G.hpp:92-96, ConstitutiveModel_Vol_dP.hpp:242-326): logJp read and written back"""
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    P["logJp"] = np.random.RandomState(44).uniform(-0.06, 0.03, n).astype(np.float32)
    margin, fr = sand_margin(P, synth.MODEL["E"], synth.MODEL["nu"], SAND)
    assert margin > 2e-5 and min(fr) > 0.05, (margin, fr)   # every branch is taken, nobody sits on a branch boundary
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.set_logJp(P["logJp"])
    nb = h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g_sand(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], SAND, P["volume"])
    g1 = h.grid()
    lj = h.get_logJp()
    mx = h.grid_update(synth.DT, synth.GRAVITY, 1)
    h.g2p(synth.DT)
    out = h.get_particles()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), nblocks=nb,
                        active_keys=tab["active_keys"], logJp_in=P["logJp"], logJp=lj, grid_p2g=g1, max_vel_sqr=mx,
                        branch_fractions=np.array(fr), x=out["x"], v=out["v"], C=out["C"], F=out["F"])
    print(name, "n", n, "blocks", nb, "tip/inside/surface %.2f/%.2f/%.2f, margin %.2e" % (fr + (margin,)))


def nacc_margin(P, nacc):
    """branch statistics of compute_stress_nacc in float64: (max tip, min tip, projected onto the surface, hardening
    solve taken, inside) and the smallest relative distance from the p_trial branch boundaries"""
    E, nu = nacc["E"], nacc["nu"]
    mu = 0.5 * E / (1 + nu)
    bm = 2.0 / 3.0 * (E / (2 * (1 + nu))) + E * nu / ((1 + nu) * (1 - 2 * nu))
    sin_phi = np.sin(np.float32(nacc["fa"]))
    M = np.sqrt(2.0 / 3.0) * 2 * sin_phi / (3 - sin_phi) * 3 / np.sqrt(2.0 / 3.0)
    F = P["F"].reshape(-1, 3, 3).transpose(0, 2, 1).astype(np.float64)
    sig = np.linalg.svd(F, compute_uv=False)
    lj = P["logJp"].astype(np.float64)
    p0 = bm * 1e-5 + np.sin(nacc["xi"] * np.maximum(-lj, 0))
    pmin = -nacc["beta"] * p0
    Je = sig.prod(1)
    B = sig ** 2
    s_hat = mu * Je[:, None] ** (-2.0 / 3.0) * (B - B.mean(1, keepdims=True))
    pt = -bm * 0.5 * (Je - 1 / Je) * Je
    y = 1.5 * (1 + 2 * nacc["beta"]) * (s_hat ** 2).sum(1) + M * M * (pt - pmin) * (pt - p0)
    c1, c2 = pt > p0, pt < pmin
    c3 = ~c1 & ~c2 & (y >= 1e-4)
    hard = c3 & (p0 > 1e-4) & (pt < p0 - 1e-4) & (pt > 1e-4 + pmin)
    margin = min(float(np.abs(pt - p0).min()), float(np.abs(pt - pmin).min()), float(np.abs(y - 1e-4)[~c1 & ~c2].min(initial=np.inf)))
    return margin, tuple(float(m.mean()) for m in (c1, c2, c3, hard, ~c1 & ~c2 & ~c3))


def nacc_case(ref, name, s, G, **kw):
    """NACCConfig substep (P2G.hpp:97-100, ConstitutiveModel_Vol_dP.hpp:116-240): logJp read and written back"""
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    P["logJp"] = np.random.RandomState(37).uniform(-1.8, 0.2, n).astype(np.float32)
    margin, fr = nacc_margin(P, NACC)
    assert margin > 1e-5 and min(fr[:4]) > 0.03, (margin, fr)
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.set_logJp(P["logJp"])
    nb = h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g_nacc(synth.DT, NACC["E"], NACC["nu"], NACC, P["volume"])
    g1 = h.grid()
    lj = h.get_logJp()
    mx = h.grid_update(synth.DT, synth.GRAVITY, 1)
    h.g2p(synth.DT)
    out = h.get_particles()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), nblocks=nb,
                        active_keys=tab["active_keys"], logJp_in=P["logJp"], logJp=lj, grid_p2g=g1, max_vel_sqr=mx,
                        branch_fractions=np.array(fr), x=out["x"], v=out["v"], C=out["C"], F=out["F"])
    print(name, "n", n, "blocks", nb, "maxtip/mintip/surface/hardening/inside %.2f/%.2f/%.2f/%.2f/%.2f, margin %.2e" % (fr + (margin,)))


def eos_case(ref, name, s, G, **kw):
    """EquationOfStateConfig{bulk=4e4, gamma=7.15, viscosity=0.01} substep (P2G.hpp:66-87, G2P.hpp:69-73)"""
    P = synth.elastic_cube(s, G, **kw)
    n, dx = P["x"].shape[0], P["dx"]
    J = (1.0 + np.random.RandomState(21).uniform(-0.05, 0.05, n)).astype(np.float32)
    h = ref.mpm(n, dx, 0)
    h.set_particles(P)
    h.set_J(J)
    nb = h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g_eos(synth.DT, 4.0e4, 7.15, 0.01, P["volume"])
    g1 = h.grid()
    mx = h.grid_update(synth.DT, synth.GRAVITY, 1)
    h.g2p_eos(synth.DT)
    out = h.get_particles()
    Jo = h.get_J()
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=s, G=G, kw=repr(sorted(kw.items())), nblocks=nb,
                        active_keys=tab["active_keys"], J_in=J, grid_p2g=g1, max_vel_sqr=mx, x=out["x"], v=out["v"],
                        C=out["C"], J=Jo)
    print(name, "n", n, "blocks", nb)


def prims_case(ref):
    rs = np.random.RandomState(12345)
    out = {}
    for n in (0, 1, 7, 1000, 4096, 10007):
        ku = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
        ki = ku.view(np.int32)
        k64 = rs.randint(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rs.randint(0, 2, size=n).astype(np.uint64)
        v = np.arange(n, dtype=np.int32)
        for kind, k in (("u32", ku), ("i32", ki), ("u64", k64)):
            ko, vo = ref.radix_sort_pair(kind, k, v)
            out["sortpair_%s_%d_k" % (kind, n)] = ko
            out["sortpair_%s_%d_v" % (kind, n)] = vo
        ko, vo = ref.radix_sort_pair("u32", ku, v, 4, 20)
        out["sortpair_u32_%d_bits4_20_k" % n] = ko
        out["sortpair_u32_%d_bits4_20_v" % n] = vo
        a = rs.randint(-1000, 1000, size=n).astype(np.int32)
        out["in_i32_%d" % n] = a
        out["in_u32_%d" % n] = ku
        out["in_u64_%d" % n] = k64
        out["exscan_i32_%d" % n] = ref.scan("exclusive", "i32", a)
        out["inscan_i32_%d" % n] = ref.scan("inclusive", "i32", a)
        for op in ("sum", "min", "max"):
            out["reduce_%s_i32_%d" % (op, n)] = np.array([ref.reduce(op, "i32", a)])
    np.savez_compressed(os.path.join(HERE, "prims.npz"), **out)
    print("prims", len(out), "arrays")


def svd_case(ref):
    rs = np.random.RandomState(3)
    Fs = np.concatenate([
        (np.eye(3).reshape(1, 9) + rs.uniform(-0.3, 0.3, (200, 9))).astype(np.float32),
        rs.uniform(-2, 2, (50, 9)).astype(np.float32),
        np.eye(3, dtype=np.float32).reshape(1, 9),
        np.diag([2.0, 0.5, 1.0]).astype(np.float32).reshape(1, 9),
        np.zeros((1, 9), np.float32),
    ])
    U = np.empty_like(Fs); V = np.empty_like(Fs); S = np.empty((Fs.shape[0], 3), np.float32)
    PF = np.empty_like(Fs)
    for i, F in enumerate(Fs):
        U[i], S[i], V[i] = ref.svd3(F)
        PF[i] = ref.stress_fixedcorotated(1.0e-6, 5.0e4, 0.4, F)
    np.savez_compressed(os.path.join(HERE, "svd.npz"), F=Fs, U=U, S=S, V=V, PF=PF)
    print("svd", Fs.shape[0])


if __name__ == "__main__":
    r = Ref()
    mpm_case(r, "mpm_cube6_mode0", 6, 32, 0, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11)
    mpm_case(r, "mpm_cube6_mode1", 6, 32, 1, jitter_F=0.05, jitter_C=0.5, shuffle_seed=11)
    mpm_case(r, "mpm_cube8_rest", 8, 32, 1)
    mpm_case(r, "mpm_cube5_neg", 5, 16, 1, jitter_F=0.02, jitter_C=0.2, origin_cells=-9)
    eos_case(r, "mpm_cube6_eos", 6, 32, jitter_C=0.5, shuffle_seed=13)
    boundary_case(r, "mpm_cube7_boundary", 7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    boundary_moving_case(r, "mpm_cube7_boundary_moving", 7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    vonmises_case(r, "mpm_cube6_vonmises", 6, 32, 2946.0, jitter_F=0.05, jitter_C=0.5, shuffle_seed=17)   # 39 % of the particles yield
    boundary_cuboid_case(r, "mpm_cube7_boundary_cuboid", 7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    sand_case(r, "mpm_cube6_sand", 6, 32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=19)
    nacc_case(r, "mpm_cube6_nacc", 6, 32, jitter_F=0.03, jitter_C=0.5, shuffle_seed=23)
    grid_momentum_case(r, "mpm_cube6_grid_momentum", 6, 32, jitter_F=0.05, jitter_C=0.5, shuffle_seed=29)
    if "--all" in sys.argv:   # the primitive / SVD vectors use unseeded-order-independent inputs: regenerate on demand
        prims_case(r)
        svd_case(r)

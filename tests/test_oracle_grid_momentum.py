"""GridMomentumToVelocity and GridAngularMomentum (simulation/grid/GridOp.hpp:184-262), CPU only: the oracle against the
reference functors run through the reference's own execution policies (sequential: bit-exact; OpenMP: the double sums to
rounding), plus the properties the quantities must have (total momentum = sum of particle momenta)."""
import numpy as np

from zpc_b200 import synth


def _grid_after_p2g(h, P):
    h.set_particles(P)
    h.partition()
    tab = h.table()
    h.clean_grid()
    h.p2g(synth.DT, synth.MODEL["E"], synth.MODEL["nu"], P["volume"])
    return tab


def test_momentum_to_velocity_bit_exact_vs_reference(oracle, ref):
    P = synth.elastic_cube(7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    n, dx = P["x"].shape[0], P["dx"]
    for threads in (0, 3):
        h = ref.mpm(n, dx, threads)
        _grid_after_p2g(h, P)
        before = h.grid()
        mx_ref = h.momentum_to_velocity()
        after = h.grid()
        h.close()
        g = before.copy()
        mx = oracle.grid_momentum_to_velocity(g)
        assert np.array_equal(g.view(np.uint32), after.view(np.uint32))
        assert np.float32(mx) == np.float32(mx_ref) and mx > 0
        assert np.array_equal(g[:, 4:7], before[:, 4:7]), "rhs channels are not touched"
        assert np.array_equal(g[:, 0], before[:, 0])
        occupied = before[:, 0] != 0
        assert occupied.sum() > 500 and (~occupied).sum() > 0
        vel = g[:, 1:4].transpose(0, 2, 1)[occupied]
        assert np.float32((vel.astype(np.float32) ** 2).sum(1).max()) == np.float32(mx) or abs((vel ** 2).sum(1).max() - mx) <= 1e-6 * mx


def test_angular_momentum_vs_reference(oracle, ref):
    P = synth.elastic_cube(7, 32, jitter_C=0.6, jitter_F=0.03, shuffle_seed=4)
    P["v"] = (P["v"] + np.float32([0.3, -0.2, 0.1])).astype(np.float32)
    n, dx = P["x"].shape[0], P["dx"]
    for threads in (0, 3):
        h = ref.mpm(n, dx, threads)
        tab = _grid_after_p2g(h, P)
        grid = h.grid()                       # the OpenMP P2G adds in another order: compare on this run's own grid
        want = h.angular_momentum()
        h.close()
        got = oracle.grid_angular_momentum(grid, tab["active_keys"], dx)
        if threads == 0:
            assert np.array_equal(got, want), "sequential policy: same addition order, same doubles"
        else:
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), "OpenMP: atomics in another order"
    # linear momentum on the grid = particle momentum (the weights sum to one), in float accuracy
    pm = (P["m"][:, None].astype(np.float64) * P["v"]).sum(0)
    assert np.abs(got[3:] - pm).max() <= 2e-5 * np.abs(pm).max()
    assert np.abs(got[:3]).max() > 0
    # other channel choice: angular "momentum" of the rhs channels against the mass channel
    alt = oracle.grid_angular_momentum(grid, tab["active_keys"], dx, 0, 4)
    assert not np.array_equal(alt, got) and np.isfinite(alt).all()


def test_golden_pins_the_oracle(oracle):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mpm_cube6_grid_momentum.npz"))
    got = oracle.grid_angular_momentum(z["grid"], z["active_keys"], np.float32(1.0 / 32))
    assert np.array_equal(got, z["sum6"])
    g = z["grid"].copy()
    mx = oracle.grid_momentum_to_velocity(g)
    assert np.array_equal(g.view(np.uint32), z["vel"].view(np.uint32)) and np.float32(mx) == z["max_vel_sqr"]

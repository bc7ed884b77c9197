"""Parity rules shared by the tests.

Integers / indices: bit-exact.

fp32 grid and particle state: |a-b| <= rtol * max(|a|, |b|, floor), floor = the channel's natural scale
(its max-abs, or the magnitude of the summands where a quantity is a cancelling sum, e.g. C = 4/dx^2 * sum W v x
is ~|v| * 4/dx per term but ~0 in a uniform field).  A sum of ~200 signed fp32 terms cannot be reproduced more
tightly by ANY re-ordering — the reference's own CUDA path adds them with unordered float atomics.

rtol = 1e-5 (BASELINE.json north_star) everywhere except the three rhs grid channels, which carry the stress of
the reference's 4-sweep approximate SVD (math/matrix/SVD.hpp): that computation is not reproducible to 1e-5 even
between two builds of the SAME source — enabling FMA contraction in the host oracle moves rhs by 1.3e-5..1.8e-5
of channel scale (tests/test_oracle_sensitivity.py measures it), and the reference's device build differs from its
host build in exactly that way (::rsqrtf vs 1/sqrtf, nvcc FMA contraction).  Measured on a B200 at
8 M particles (profiles/r02_parity_vs_reference_c2.md): the reference's OWN CUDA and OpenMP paths differ by 2.8e-5 of
channel scale on rhs (1.4e-6 on m / mv); this library sits at 1.6e-5 (AoS) .. 2.1e-5 (binned) from the reference CUDA
path and 2.6e-5 from the OpenMP path — closer to each than they are to each other — and at <= 1.2e-6 on every other
channel of the grid and of the particles.  rhs is therefore held to RTOL_STRESS = 5e-5 (round 1: 1e-4), everything else
to 1e-5.

The stricter floor of 1e-3 * channel max-abs from the survey is applied where asked (strict_frac) as the
fraction of entries that must meet rtol under it.
"""
import numpy as np

RTOL = 1e-5
RTOL_STRESS = 5e-5
GRID_RTOL = [RTOL] * 4 + [RTOL_STRESS] * 3          # channels m, mv(3), rhs(3)


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return np.abs(a - b) / den


def check_channels(a, b, axis_channels, what, rtol=RTOL, floor=None, strict_frac=None):
    """a, b: arrays whose axis `axis_channels` enumerates channels; each channel gets its own scale.
    rtol: scalar or per-channel list.  floor: extra natural scale (scalar) added to the channel max-abs."""
    a = np.moveaxis(np.asarray(a), axis_channels, 0)
    b = np.moveaxis(np.asarray(b), axis_channels, 0)
    worst = 0.0
    for c in range(a.shape[0]):
        rt = rtol[c] if isinstance(rtol, (list, tuple)) else rtol
        scale = float(max(np.abs(a[c]).max(), np.abs(b[c]).max()))
        if floor is not None:
            scale = max(scale, float(floor))
        if scale == 0.0:
            continue
        e = rel_err(a[c], b[c], scale)
        worst = max(worst, float(e.max()) / rt)
        assert e.max() <= rt, "%s channel %d: max rel err %.3e > %.0e (scale %.3e)" % (what, c, e.max(), rt, scale)
        if strict_frac is not None and rt <= RTOL:   # the strict floor is only meaningful at the 1e-5 level
            es = rel_err(a[c], b[c], 1e-3 * scale)
            frac = float((es <= rt).mean())
            assert frac >= strict_frac, "%s channel %d: only %.4f of entries within strict tolerance" % (what, c, frac)
    return worst


def check_particles(got, want, dx, what, rtol=RTOL):
    """x, v, C, F after G2P.  Components of one vector / tensor share a physical scale, so the floor is the
    attribute's max-abs over all components (F ~ 1: an off-diagonal of 1e-9 is rounding noise of I + dt*C);
    C's floor is the size of its summands, 4/dx * max|v| (G2P.hpp:65)."""
    vmax = float(np.abs(want["v"]).max())
    check_channels(got["x"], want["x"], 1, what + " x", rtol, floor=float(np.abs(want["x"]).max()))
    check_channels(got["v"], want["v"], 1, what + " v", rtol, floor=vmax)
    check_channels(got["C"], want["C"], 1, what + " C", rtol, floor=4.0 / dx * vmax)
    check_channels(got["F"], want["F"], 1, what + " F", rtol, floor=float(np.abs(want["F"]).max()))


def grid_by_key(keys, grid):
    """dict-free canonical form: rows sorted by block key (removes block numbering)."""
    keys = np.asarray(keys)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order], np.asarray(grid)[order]


# ---- moving analytic colliders used by the CPU pin test, the golden generator and the GPU test --------------------
def _rot(axis, angle):
    a = np.asarray(axis, np.float64); a /= np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return (np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K).astype(np.float32)


MOVING_COLLIDERS = [  # (geometry, collider type, p0, p1, motion = b, dbdt, R, omega, s, dsdt)
    (0, 0, (0.0, 0.10, 0.0), (0.0, 1.0, 0.0), ((0.0, 0.2, 0.0), (0.3, 0.5, -0.2), np.eye(3, dtype=np.float32), (0, 0, 0), 1.0, 0.0)),
    (0, 2, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), ((0.3, 0.3, 0.3), (0.0, 1.0, 0.0), _rot((0, 0, 1), 0.4), (0.0, 0.0, 2.0), 1.0, 0.0)),
    (1, 1, (0.0, 0.0, 0.0), (0.05, 0.0, 0.0), ((0.33, 0.3, 0.33), (0.2, 0.0, 0.1), _rot((1, 2, 3), 1.1), (1.0, -2.0, 0.5), 1.5, 0.7)),
    (1, 0, (0.01, 0.0, 0.0), (0.06, 0.0, 0.0), ((0.3, 0.32, 0.3), (0.0, -0.4, 0.0), _rot((0, 1, 0), 2.0), (0.0, 3.0, 0.0), 0.9, -0.2)),
    (1, 2, (0.0, 0.0, 0.0), (0.08, 0.0, 0.0), ((0.33, 0.3, 0.33), (0.0, 0.0, 0.0), np.eye(3, dtype=np.float32), (0, 0, 0), 1.0, 0.0)),
]


# AnalyticLevelSet<Cuboid>{min, max} colliders (geometry/AnalyticLevelSet.h:55-126): static and moving, all three types
CUBOID_COLLIDERS = [  # (geometry = 2, collider type, box min, box max, motion or None)
    (2, 0, (0.2, 0.2, 0.2), (0.3, 0.3, 0.45), None),
    (2, 1, (0.2, 0.2, 0.2), (0.3, 0.3, 0.45), None),
    (2, 2, (0.25, 0.1, 0.2), (0.5, 0.3, 0.33), None),
    (2, 1, (-0.05, -0.04, -0.06), (0.05, 0.04, 0.06), ((0.3, 0.3, 0.3), (0.2, 0.0, 0.1), _rot((1, 2, 3), 0.7), (1.0, -2.0, 0.5), 1.2, 0.3)),
    (2, 2, (-0.05, -0.04, -0.06), (0.05, 0.04, 0.06), ((0.33, 0.27, 0.3), (0.0, 0.5, 0.0), _rot((0, 0, 1), 0.4), (0.0, 0.0, 2.0), 1.0, 0.0)),
]


def motion_vec(m):
    if m is None:
        return None
    b, dbdt, R, om, s, dsdt = m
    return np.concatenate([np.asarray(b, np.float32), np.asarray(dbdt, np.float32), np.asarray(R, np.float32).reshape(9),
                           np.asarray(om, np.float32), np.asarray([s, dsdt], np.float32)]).astype(np.float32)



"""Deterministic synthetic particle clouds — one generator for the CPU and GPU arms.

Elastic cube of ``s^3`` cells at 8 particles per cell in a ``G^3`` domain (dx = 1/G), cube origin 7
cells from the domain corner.  Particle of cell (i,j,k), sub-slot sigma in [0,8):
``x = (7 + {i,j,k} + (sigma_bit + u)/2) * dx`` with ``u`` a 24-bit uniform from ``mt19937(seed)`` drawn
in (x,y,z) order; v = (0,-1,0), m = rho*vol, rho = 1000, vol = dx^3/8, F = I, C = 0.
Host-side numpy only: this is input generation, not part of the timed path.
"""
import numpy as np

CONFIGS = {
    # name: (G, s)   -> N = 8 s^3
    "C1": (64, 50),    # 1.0 M particles, 64^3 domain (reference CPU case)
    "C2": (128, 100),  # 8.0 M particles, 128^3
    "C3": (256, 200),  # 64.0 M particles, 256^3
}
MODEL = dict(E=5.0e4, nu=0.4, rho=1000.0)
DT = 1.0e-4
GRAVITY = -9.8


def _mt19937_u24(n, seed):
    """n floats in [0,1): top 24 bits of consecutive std::mt19937(seed) outputs, times 2^-24."""
    rs = np.random.RandomState(seed)  # init_genrand(seed) == std::mt19937(seed)
    raw = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    return (raw >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _mt19937_u24_slice(first, count, seed, chunk=1 << 24):
    """elements [first, first+count) of the _mt19937_u24 stream without holding the whole stream"""
    rs = np.random.RandomState(seed)
    out = np.empty(count, np.float32)
    pos, filled = 0, 0
    end = first + count
    while pos < end:
        m = min(chunk, end - pos)
        raw = rs.randint(0, 2 ** 32, size=m, dtype=np.uint64)
        lo, hi = max(first, pos), min(end, pos + m)
        if hi > lo:
            seg = raw[lo - pos:hi - pos].astype(np.uint32)
            out[filled:filled + (hi - lo)] = (seg >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
            filled += hi - lo
        pos += m
    return out


def elastic_cube(s, G, seed=7, shuffle_seed=None, jitter_F=0.0, jitter_C=0.0, origin_cells=7, cell_range=None):
    """Returns dict of AoS float32 arrays x[N,3] v[N,3] m[N] C[N,9] F[N,9] plus dx, volume.
    cell_range=(c0,c1) keeps only the cells with global index (i*s*s + j*s + k) in [c0,c1): the same particles,
    bit for bit, as that slice of the full cube (used to shard the cloud over ranks)."""
    dx = np.float32(1.0 / G)
    c0, c1 = cell_range if cell_range is not None else (0, s ** 3)
    ncell = c1 - c0
    n = 8 * ncell
    cell = np.arange(c0, c1, dtype=np.int64)
    ci = np.stack([cell // (s * s), (cell // s) % s, cell % s], axis=1)  # (i,j,k), k fastest
    sig = np.arange(8, dtype=np.int64)
    sbit = np.stack([(sig >> 2) & 1, (sig >> 1) & 1, sig & 1], axis=1)
    u = _mt19937_u24_slice(c0 * 24, 3 * n, seed).reshape(n, 3)
    base = (ci[:, None, :] + origin_cells).astype(np.float32)          # [cells,1,3]
    sub = sbit[None, :, :].astype(np.float32)                          # [1,8,3]
    x = (base + (sub + u.reshape(ncell, 8, 3)) * np.float32(0.5)) * dx
    x = np.ascontiguousarray(x.reshape(n, 3), np.float32)
    vol = np.float32(dx * dx * dx / np.float32(8.0))
    P = dict(
        x=x,
        v=np.ascontiguousarray(np.tile(np.array([0.0, -1.0, 0.0], np.float32), (n, 1))),
        m=np.full(n, np.float32(MODEL["rho"]) * vol, np.float32),
        C=np.zeros((n, 9), np.float32),
        F=np.ascontiguousarray(np.tile(np.eye(3, dtype=np.float32).reshape(9), (n, 1))),
    )
    if jitter_F or jitter_C:
        rs = np.random.RandomState(seed + 1000)
        if jitter_F:
            P["F"] += (rs.uniform(-jitter_F, jitter_F, (n, 9))).astype(np.float32)
        if jitter_C:
            P["C"] += (rs.uniform(-jitter_C, jitter_C, (n, 9))).astype(np.float32)
            P["v"] += (rs.uniform(-0.5, 0.5, (n, 3))).astype(np.float32)
    if shuffle_seed is not None:
        perm = np.random.RandomState(shuffle_seed).permutation(n)
        for k in ("x", "v", "m", "C", "F"):
            P[k] = np.ascontiguousarray(P[k][perm])
    P["dx"] = float(dx)
    P["volume"] = float(vol)
    return P


def slab_cell_range(s, rank, world):
    """x-slab decomposition: rank r owns the cells with i in [r*s//world, (r+1)*s//world)"""
    i0, i1 = rank * s // world, (rank + 1) * s // world
    return i0 * s * s, i1 * s * s


def elastic_cube_slab(s, G, rank, world, **kw):
    return elastic_cube(s, G, cell_range=slab_cell_range(s, rank, world), **kw)


def config(name, **kw):
    G, s = CONFIGS[name]
    return elastic_cube(s, G, **kw)

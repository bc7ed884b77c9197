"""Parity rules shared by the tests.

Integers / indices: bit-exact.  fp32 grid and particle state: |a-b| <= 1e-5 * max(|a|, |b|, floor) with
floor = the channel's max-abs (a sum of ~200 signed fp32 terms cannot be reproduced more tightly by ANY
re-ordering, including the reference's own atomics).  The stricter floor of 1e-3 * channel max-abs from the
survey is reported as the fraction of entries that meet it.
"""
import numpy as np

RTOL = 1e-5


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return np.abs(a - b) / den


def check_channels(a, b, axis_channels, what, rtol=RTOL, strict_frac=0.99):
    """a, b: arrays whose axis `axis_channels` enumerates channels; each channel gets its own scale."""
    a = np.moveaxis(np.asarray(a), axis_channels, 0)
    b = np.moveaxis(np.asarray(b), axis_channels, 0)
    worst = 0.0
    for c in range(a.shape[0]):
        scale = float(max(np.abs(a[c]).max(), np.abs(b[c]).max()))
        if scale == 0.0:
            continue
        e = rel_err(a[c], b[c], scale)
        worst = max(worst, float(e.max()))
        assert e.max() <= rtol, "%s channel %d: max rel err %.3e (scale %.3e)" % (what, c, e.max(), scale)
        es = rel_err(a[c], b[c], 1e-3 * scale)
        frac = float((es <= rtol).mean())
        assert frac >= strict_frac, "%s channel %d: only %.4f of entries within strict tolerance" % (what, c, frac)
    return worst


def grid_by_key(keys, grid):
    """dict-free canonical form: rows sorted by block key (removes block numbering)."""
    keys = np.asarray(keys)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return keys[order], np.asarray(grid)[order]

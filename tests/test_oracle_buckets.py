"""index_buckets_for_particles (simulation/particle/Query.tpp:9-58, SparsityOp.hpp:115-195), CPU only: the oracle
is bit-exact against the reference's functor sequence on seq_exec (cell table, counts, offsets, ids) and equal bucket by
bucket on omp_exec (whose numbering and in-bucket order are racy)."""
import numpy as np
import pytest

from zpc_b200 import synth


def by_key(D):
    return {tuple(k): tuple(sorted(D["ids"][D["offsets"][i]: D["offsets"][i] + D["counts"][i]])) for i, k in enumerate(D["active_keys"])}


@pytest.mark.parametrize("disp", [0.5, 0.0])
@pytest.mark.parametrize("case", [dict(s=10, G=32, shuffle_seed=5), dict(s=6, G=16, origin_cells=-9, shuffle_seed=2), dict(s=1, G=8)])
def test_index_buckets_bit_exact_vs_reference(oracle, ref, case, disp):
    kw = dict(case)
    P = synth.elastic_cube(kw.pop("s"), kw.pop("G"), **kw)
    x, n, dx = P["x"], P["x"].shape[0], P["dx"]
    A = oracle.index_buckets(x, dx, disp, oracle.table_size_for(n))
    B = ref.index_buckets(x, dx, disp, 0)
    for k in ("active_keys", "counts", "offsets", "ids"):
        assert np.array_equal(A[k], B[k]), k
    assert by_key(A) == by_key(ref.index_buckets(x, dx, disp, 8))
    # structure: offsets = exclusive scan of counts, every particle once, ids ascending inside a bucket, right cell
    assert A["counts"][-1] == 0 and A["offsets"][-1] == n and np.array_equal(np.sort(A["ids"]), np.arange(n))
    assert np.array_equal(np.cumsum(A["counts"])[:-1], A["offsets"][1:])
    cells = np.floor(x / np.float32(dx) + np.float32(disp)).astype(np.int32)
    for b in range(0, A["nblocks"], max(A["nblocks"] // 50, 1)):
        ids = A["ids"][A["offsets"][b]: A["offsets"][b + 1]]
        assert (np.diff(ids) > 0).all() and (cells[ids] == A["active_keys"][b]).all()

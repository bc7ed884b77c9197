"""oracle/sparse_oracle.c (bht<i32,3,int,16>, SparseGrid<3,f32,8> accessors) pinned against the reference's own
containers (oracle/_ref/libzpcref.so = unmodified reference headers).  CPU only."""
import numpy as np
import pytest

from oracle.pyoracle import Ref


def _keys(rs, n, span=40):
    k = rs.randint(-span, span, (n * 2, 3)).astype(np.int32) * 8
    k = np.unique(k, axis=0)
    rs.shuffle(k)
    return np.ascontiguousarray(k[:n])


def test_bht_hash_parameters_and_table_size(oracle, ref):
    for expected in (1, 7, 100, 128, 129, 4096, 100000):
        t = Ref.Bht(ref, expected)
        i = t.info()
        assert oracle.bht_table_size(expected) == i["table_size"], expected
        assert i["num_buckets"] * 16 == i["table_size"]
        assert np.array_equal(oracle.bht_params(), i["hf"])          # std::mt19937(2), six draws
        t.close()


@pytest.mark.parametrize("n,expected", [(50, 100), (1000, 1000), (5000, 4096), (2000, 1024)])
def test_bht_serial_insert_matches_reference_slot_for_slot(oracle, ref, n, expected):
    """sequential inserts in the same order: identical keys / indices arrays, identical query answers"""
    rs = np.random.RandomState(n)
    keys = _keys(rs, n)
    dup = np.concatenate([keys, keys[: n // 3]])                      # re-inserting returns the sentinel
    t = Ref.Bht(ref, expected)
    r_out = t.insert(dup)
    o = oracle.bht_new(expected)
    o_out = oracle.bht_insert(o, dup)
    assert np.array_equal(r_out, o_out)
    a = t.arrays()
    n_in = int((r_out[:n] >= 0).sum())          # (2000 into 1024: three full buckets -> failure token, in both)
    assert n_in == n or n > expected
    assert a["cnt"] == int(o["cnt"][0]) == n_in
    assert np.array_equal(a["keys16"][:, :3], o["keys16"][:, :3])
    occ = a["keys16"][:, 0] != 0x3F3F3F3F
    assert np.array_equal(a["indices"][occ], o["indices"][occ])
    assert np.array_equal(a["active_keys"], o["active_keys"][:n_in])
    assert (a["status"] == -1).all() and (o["status"] == -1).all()   # locks released
    miss = _keys(np.random.RandomState(n + 1), 200, span=400) + 4     # not multiples of 8: never inserted
    q = np.concatenate([keys, miss])
    assert np.array_equal(t.query(q), oracle.bht_query(o, q))
    assert (oracle.bht_query(o, miss) == -1).all()
    t.close()


def test_oracle_built_table_is_readable_by_the_reference_query(oracle, ref):
    """a table produced elsewhere (here: the oracle's side-8 partition) loaded into the reference container resolves
    through the UNMODIFIED BHTView::query — the check the GPU-built tables go through as well"""
    from zpc_b200 import synth
    P = synth.elastic_cube(12, 32, shuffle_seed=4, origin_cells=-5)
    t = oracle.sg_partition_build(P["x"], P["dx"], 512)
    nb = t["nblocks"]
    ak = t["active_keys"][:nb]
    assert (ak % 8 == 0).all() and np.unique(ak, axis=0).shape[0] == nb
    r = Ref.Bht(ref, 512)
    r.load(t["keys16"], t["indices"], ak, nb)
    assert np.array_equal(r.query(ak), np.arange(nb))
    assert (r.query(ak + 1) == -1).all()
    # every particle's 3^3 stencil lies in active blocks
    X = P["x"] / P["dx"]
    base = np.floor(X - 0.5).astype(np.int32)
    for o in ((0, 0, 0), (2, 2, 2), (0, 2, 0)):
        c = base + np.array(o, np.int32)
        assert (r.query(c - (c & 7)) >= 0).all()
    r.close()


def test_sparsegrid_accessors_match_reference(oracle, ref):
    rs = np.random.RandomState(3)
    nb, nch = 40, 4
    keys = _keys(rs, nb, span=6)
    sg = Ref.SparseGrid(ref, nb, nch)
    sg.table.insert(keys)
    sg.scale(0.125)
    sg.translate([0.5, -1.0, 2.0])
    sg.set_background(-7.5)
    grid = rs.uniform(-1, 1, (nb, nch, 512)).astype(np.float32)
    sg.load_grid(grid)
    o = oracle.bht_new(nb)
    oracle.bht_insert(o, keys)
    # valueOr at cells of active blocks and at cells of absent blocks
    inside = keys[rs.randint(0, nb, 300)] + rs.randint(0, 8, (300, 3)).astype(np.int32)
    outside = rs.randint(-400, 400, (300, 3)).astype(np.int32)
    coords = np.concatenate([inside, outside])
    for chn in (0, 3):
        a = sg.value_or(chn, coords, 42.0)
        b = oracle.sg_value_or(o, grid, chn, coords, 42.0)
        assert np.array_equal(a, b)
        assert (a[:300] != 42.0).all()
    # iCoord / wCoord through the index-to-world transform
    bno = rs.randint(0, nb, 200); cno = rs.randint(0, 512, 200)
    ic_r, wc_r = sg.coords(bno, cno)
    ic_o, wc_o = oracle.sg_coords(o, sg.transform(), bno, cno)
    assert np.array_equal(ic_r, ic_o)
    np.testing.assert_allclose(wc_r, wc_o, rtol=1e-6, atol=1e-6)
    sg.close()


@pytest.mark.parametrize("scatter", [False, True])
def test_reorder_tiles_and_bht_reorder_match_reference(oracle, ref, scatter):
    """TileVector::reorderTiles + bht::reorder (the SparseGrid renumbering utilities) vs the reference containers"""
    rs = np.random.RandomState(12)
    nb, nch = 37, 3
    keys = _keys(rs, nb, span=5)
    perm = rs.permutation(nb).astype(np.int32)
    grid = rs.uniform(-1, 1, (nb, nch, 512)).astype(np.float32)
    sg = Ref.SparseGrid(ref, nb, nch)
    sg.table.insert(keys)
    sg.load_grid(grid)
    o = oracle.bht_new(nb)
    oracle.bht_insert(o, keys)
    o["active_keys"] = o["active_keys"][:nb].copy()
    # tiles
    assert np.array_equal(sg.reorder_tiles(perm, scatter), oracle.tilevector_reorder_tiles(grid, perm, scatter))
    # table
    sg.table.reorder(perm, scatter)
    a = sg.table.arrays()
    o2 = oracle.bht_reorder(o, perm, scatter)
    assert np.array_equal(a["active_keys"], o2["active_keys"])
    occ = a["keys16"][:, 0] != 0x3F3F3F3F
    assert np.array_equal(a["indices"][occ], o2["indices"][occ])
    # after renumbering, key k of the new order resolves to its new index
    assert np.array_equal(oracle.bht_query(o2, o2["active_keys"]), np.arange(nb))
    sg.close()

"""Host-side layout logic without a GPU: the iterator-port address formula (py_interop/GenericIterator.hpp:84-98)
against the TileVector AoSoA layout (container/TileVector.hpp:108,768-769), slab sharding, block-key packing."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def port_address(p, k):
    i = p.idx + k
    return (((i >> p.numTileBits) * p.numChns) << p.numTileBits) | (i & p.tileMask)


@pytest.mark.parametrize("L,nch,n", [(32, 25, 1000), (32, 3, 70001), (64, 7, 640)])
def test_port_formula_matches_tilevector_layout(L, nch, n):
    from zpc_b200 import api
    tv = api.TileVector(n, nch, L, device="cpu")
    for chn in (0, nch // 2, nch - 1):
        vals = torch.arange(n, dtype=torch.float32) + 1000 * chn
        tv.set_channel(chn, vals)
    flat = tv.data.numpy()
    for chn in (0, nch // 2, nch - 1):
        p = tv.port(chn)
        base_off = (p.base - tv.data.data_ptr()) // 4
        assert base_off == chn * L
        for k in (0, 1, L - 1, L, L + 1, n // 2, n - 1):
            assert flat[base_off + port_address(p, k)] == k + 1000 * chn
            # reference layout: element (chn, i) at (i // L * nch + chn) * L + i % L
            assert base_off + port_address(p, k) == (k // L * nch + chn) * L + k % L
        assert torch.equal(tv.channel(chn)[:, 0], torch.arange(n, dtype=torch.float32) + 1000 * chn)
    # a contiguous pointer is the degenerate port {ptr, 0, 0, 0, 1}
    q = api.zpc_port(0, 5, 0, 0, 1)
    assert [port_address(q, k) for k in range(4)] == [5, 6, 7, 8]
    # AoS vec3: element k component d at 3k + d
    v = api.zpc_port(0, 0, 0, 0, 3)
    assert [port_address(v, k) for k in range(3)] == [0, 3, 6] and v.tileMask + 1 == 1


def test_slab_shards_tile_the_cloud():
    from zpc_b200 import synth
    s = 10
    for world in (2, 3, 4, 8):
        ranges = [synth.slab_cell_range(s, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == s ** 3
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
    full = synth.elastic_cube(6, 32)
    parts = [synth.elastic_cube_slab(6, 32, r, 3) for r in range(3)]
    for k in ("x", "v", "m", "C", "F"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k])


def test_block_key_packing_is_order_preserving():
    from zpc_b200.dist_solver import pack_keys
    rs = np.random.RandomState(0)
    keys = rs.randint(-500, 500, (2000, 3)).astype(np.int32)
    keys = np.unique(keys, axis=0)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    codes = pack_keys(torch.from_numpy(keys[order])).numpy()
    assert (np.diff(codes) > 0).all()


def test_table_sizing_matches_reference_rule():
    from zpc_b200 import api
    for n, ts in ((1, 16), (2, 32), (3, 64), (1000, 16384), (1024, 16384), (1025, 32768)):
        assert api.next_2pow(n) * 16 == ts      # HashTable.hpp:87-90: next_2pow(entries) * reserve_ratio_v(16)


def test_bht_host_helpers_match_oracle(oracle):
    """zpcb200_bht_params / zpcb200_bht_table_size (host functions of the C ABI) against the pinned restatement"""
    import ctypes as C
    from zpc_b200 import api
    hf = (C.c_uint32 * 6)()
    api.lib().zpcb200_bht_params(hf)
    assert [int(v) for v in hf] == [int(v) for v in oracle.bht_params()]
    for n in (0, 1, 7, 100, 128, 129, 4096, 100000, 1 << 20):
        assert int(api.lib().zpcb200_bht_table_size(n)) == oracle.bht_table_size(n), n


def test_sparsegrid_host_transform_matches_reference(ref):
    """SparseGrid.scale / translate compose the index-to-world matrix like the reference's Transform"""
    from oracle.pyoracle import Ref
    from zpc_b200 import api
    sg = api.SparseGrid(4, 16, device="cpu")
    r = Ref.SparseGrid(ref, 16, 4)
    for obj in (sg, r):
        obj.scale(0.125)
        obj.translate([0.5, -1.0, 2.0])
    assert np.allclose(np.array(sg.transform, np.float32), r.transform(), rtol=0, atol=0)
    r.close()


def test_shard_by_blocks_cuts_equal_particle_counts_without_splitting_blocks():
    """sorted block keys, prefix sum of particles per block, contiguous ranges of equal particle counts"""
    import numpy as np
    from zpc_b200 import synth
    from zpc_b200.dist_solver import shard_by_blocks
    P = synth.elastic_cube(20, 64, shuffle_seed=3)
    x, dx, n = P["x"], P["dx"], P["x"].shape[0]
    blk = (np.floor(x / np.float32(dx) + np.float32(0.5)).astype(np.int64) - 2) >> 2
    for order in ("xmajor", "morton"):
        for world in (1, 2, 3, 8):
            owner, cuts, keys = shard_by_blocks(x, dx, world, order)
            assert owner.shape == (n,) and owner.min() == 0 and owner.max() == world - 1
            assert cuts[0] == 0 and cuts[-1] == len(keys) and (np.diff(cuts) > 0).all()
            # no block is split, and the sorted block list is cut into contiguous ranges
            code = (blk[:, 0] * 1000 + blk[:, 1]) * 1000 + blk[:, 2]
            for c in np.unique(code)[::7]:
                assert np.unique(owner[code == c]).size == 1
            kcode = (keys[:, 0].astype(np.int64) * 1000 + keys[:, 1]) * 1000 + keys[:, 2]
            assert np.unique(kcode).size == len(keys) == np.unique(code).size
            for r in range(world):
                mine = np.unique(code[owner == r])
                assert set(mine.tolist()) == set(kcode[cuts[r]:cuts[r + 1]].tolist())
            # balance: within one block's worth of particles (<= 512 at 8 ppc) of n / world
            per = np.bincount(owner, minlength=world)
            assert np.abs(per - n / world).max() <= 512, (order, world, per)
            if order == "xmajor" and world > 1:                   # x-major cuts are slabs: ranks are ordered along x
                xm = [x[owner == r, 0].mean() for r in range(world)]
                assert (np.diff(xm) > 0).all()
    owner, cuts, keys = shard_by_blocks(np.zeros((0, 3), np.float32), dx, 4)
    assert owner.shape == (0,) and len(keys) == 0

"""CUDA primitives (through the C ABI) against the oracle: bit-exact for integers and sort indices."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 31, 32, 33, 1000, 4095, 4096, 4097, 65536 + 17, 1000003]
NPDT = {"i32": np.int32, "u32": np.uint32, "i64": np.int64, "f32": np.float32, "u64": np.uint64}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def tdt(kind):
    return {"i32": torch.int32, "u32": torch.uint32, "i64": torch.int64, "f32": torch.float32, "u64": torch.uint64, "f64": torch.float64}[kind]


@pytest.fixture(scope="module")
def pol():
    from zpc_b200 import api
    assert torch.cuda.is_available()
    return api.cuda_exec()


def rand(kind, n, seed, lo=-1000, hi=1000):
    rs = np.random.RandomState(seed)
    if kind == "f32":
        return rs.uniform(-1, 1, n).astype(np.float32)
    if kind in ("u32", "u64"):
        return rs.randint(0, hi, size=n).astype(NPDT[kind])
    return rs.randint(lo, hi, size=n).astype(NPDT[kind])


@pytest.mark.parametrize("kind", ["i32", "u32", "i64"])
@pytest.mark.parametrize("n", SIZES)
def test_reduce_int_exact(pol, oracle, kind, n):
    a = rand(kind, n, n + 1)
    d = dev(a)
    out = torch.zeros(1, dtype=tdt(kind), device="cuda")
    for op in ("sum", "prod", "min", "max"):   # integer products wrap: associative, so the parallel fold is exact too
        pol.reduce(d, out, op)
        assert out.cpu().numpy()[0] == oracle.reduce(op, kind, a), (op, n)


def test_reduce_prod_float_and_c_layer(pol, oracle):
    """reduce_prod__b200_<T>_1 (py_interop/cuda/ExecutionPolicy.cpp:48-54: identity 1, zs::multiplies): floats near 1 so that the
    product stays in range; the parallel fold re-associates, so both folds are compared with the float64 product"""
    import ctypes as C
    from zpc_b200 import api
    L = api.lib()
    p = C.c_void_p(L.policy__b200())
    L.policy_set__b200(p, 0, None, 1)
    try:
        for kind, npdt, tol in (("f32", np.float32, 3e-4), ("f64", np.float64, 1e-12)):
            a = (1.0 + np.random.RandomState(9).uniform(-1e-3, 1e-3, 100003)).astype(npdt)
            d = dev(a)
            out = torch.zeros(1, dtype=tdt(kind), device="cuda")
            pol.reduce(d, out, "prod")
            # Both folds are compared with the float64 product.  A tree of products of numbers that straddle 1.0 is NOT the more
            # accurate one: results just above 1 round on a grid twice as coarse as results just below, which leaves a bias of about
            # -1e-9 per fp32 multiplication (measured; numpy's pairwise order gives -1.1e-4 on this input, the device's order -5.8e-5),
            # while the sequential accumulator leaves the neighbourhood of 1.0 after a few hundred factors (1.9e-6).
            exact = float(np.prod(a.astype(np.float64)))
            want = oracle.reduce("prod", kind, a)
            assert abs(out.item() / exact - 1.0) <= tol, (kind, out.item(), exact)
            assert abs(float(want) / exact - 1.0) <= tol, (kind, want, exact)
            out2 = torch.zeros_like(out)
            getattr(L, "reduce_prod__b200_%s_1" % ("float" if kind == "f32" else "double"))(p, api.port(d), api.port(d, a.size), api.port(out2))
            assert out2.item() == out.item()
        # odd integers: the product never collapses to 0 mod 2^32, every factor matters, and wrap-around arithmetic is associative
        odd = (np.random.RandomState(4).randint(-50000, 50000, 250001) * 2 + 1).astype(np.int32)
        do, oi = dev(odd), torch.zeros(1, dtype=torch.int32, device="cuda")
        pol.reduce(do, oi, "prod")
        assert oi.item() == int(oracle.reduce("prod", "i32", odd)) != 0
        e = torch.zeros(1, dtype=torch.int32, device="cuda")
        L.reduce_prod__b200_int_1(p, api.port(e), api.port(e), api.port(e))      # empty range -> the identity
        assert e.item() == 1
    finally:
        L.del_policy__b200(p)


@pytest.mark.parametrize("n", [1, 2, 7, 16, 128, 1024, 2000000])
def test_reference_reduce_test_on_gpu(pol, n):
    """test/cuda/main.cu + test/utils/parallel_primitives.hpp:9-32 replayed: getmax/getmin/plus<int>."""
    a = np.random.RandomState(n).randint(-2 ** 20, 2 ** 20, size=n).astype(np.int32)
    d = dev(a)
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    pol.reduce(d, out, "max"); assert out.item() == a.max()
    pol.reduce(d, out, "min"); assert out.item() == a.min()
    pol.reduce(d, out, "sum"); assert out.item() == int(a.sum(dtype=np.int32))


@pytest.mark.parametrize("n", [1, 1000, 4097, 1000003])
def test_reduce_f32(pol, oracle, n):
    a = rand("f32", n, n)
    d = dev(a)
    out = torch.zeros(1, dtype=torch.float32, device="cuda")
    pol.reduce(d, out, "max"); assert out.item() == a.max()
    pol.reduce(d, out, "min"); assert out.item() == a.min()
    pol.reduce(d, out, "sum")
    exact = float(a.sum(dtype=np.float64))
    assert abs(out.item() - exact) <= 1e-6 * float(np.abs(a).sum(dtype=np.float64))   # rel. 1e-6 of the L1 mass
    r1 = out.item(); pol.reduce(d, out, "sum"); assert out.item() == r1   # run-to-run deterministic


@pytest.mark.parametrize("kind", ["i32", "u32", "i64"])
@pytest.mark.parametrize("n", SIZES)
def test_scans_int_exact(pol, oracle, kind, n):
    a = rand(kind, n, n + 3, 0, 100)
    d = dev(a)
    out = torch.empty_like(d)
    pol.exclusive_scan(d, out)
    assert np.array_equal(out.cpu().numpy(), oracle.scan("exclusive", kind, a))
    pol.inclusive_scan(d, out)
    assert np.array_equal(out.cpu().numpy(), oracle.scan("inclusive", kind, a))


def test_scan_in_place_and_unaligned(pol, oracle):
    a = rand("i32", 100001, 7, 0, 9)
    big = dev(np.concatenate([[0], a]).astype(np.int32))
    view = big[1:]                      # 4-byte aligned only
    out = torch.empty(100002, dtype=torch.int32, device="cuda")[1:]
    pol.exclusive_scan(view, out)
    assert np.array_equal(out.cpu().numpy(), oracle.scan("exclusive", "i32", a))
    pol.inclusive_scan(view, view)      # in place
    assert np.array_equal(view.cpu().numpy(), oracle.scan("inclusive", "i32", a))


def test_scan_f32_tolerance(pol):
    a = rand("f32", 300000, 1)
    d = dev(a); out = torch.empty_like(d)
    pol.inclusive_scan(d, out)
    exact = np.cumsum(a.astype(np.float64))
    assert np.abs(out.cpu().numpy() - exact).max() <= 1e-5 * np.abs(a).sum()


def test_aosoa_ports(pol, oracle):
    """TileVector<int,32> channel 'b' as in the reference test (TileVector{{"a",1},{"b",1}}), reduce + scan + sort."""
    from zpc_b200 import api
    n = 70001
    tv = api.TileVector(n, 3, 32, dtype=torch.int32)
    a = rand("i32", n, 11)
    tv.set_channel(1, dev(a))
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    for op in ("sum", "min", "max"):
        pol.reduce((tv, 1), out, op)
        assert out.item() == oracle.reduce(op, "i32", a)
    pol.exclusive_scan((tv, 1), (tv, 2))
    assert np.array_equal(tv.channel(2)[:, 0].cpu().numpy(), oracle.scan("exclusive", "i32", a))
    ko = torch.empty(n, dtype=torch.int32, device="cuda")
    pol.radix_sort((tv, 1), ko)
    assert np.array_equal(ko.cpu().numpy(), np.sort(a, kind="stable"))
    assert np.array_equal(tv.channel(0)[:, 0].cpu().numpy(), np.zeros(n, np.int32))   # untouched channel


@pytest.mark.parametrize("kind", ["u32", "i32", "u64"])
@pytest.mark.parametrize("n", SIZES)
def test_radix_sort_pair_index_exact(pol, oracle, kind, n):
    rs = np.random.RandomState(n + 17)
    if kind == "u64":
        k = (rs.randint(0, 2 ** 32, size=n, dtype=np.uint64) << np.uint64(32)) | rs.randint(0, 2 ** 32, size=n, dtype=np.uint64)
    else:
        k = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
        if kind == "i32":
            k = k.view(np.int32)
    v = np.arange(n, dtype=np.int32)
    ko = torch.empty(n, dtype=tdt(kind), device="cuda"); vo = torch.empty(n, dtype=torch.int32, device="cuda")
    pol.radix_sort_pair(dev(k), dev(v), ko, vo, kind=kind)
    eo, ev = oracle.radix_sort_pair(kind, k, v)
    assert np.array_equal(ko.cpu().numpy(), eo) and np.array_equal(vo.cpu().numpy(), ev)
    ko2 = torch.empty_like(ko)
    pol.radix_sort(dev(k), ko2, kind=kind)
    assert np.array_equal(ko2.cpu().numpy(), eo)


@pytest.mark.parametrize("sbit,ebit", [(0, 8), (6, 24), (3, 17), (0, 12), (20, 32), (5, 5)])
def test_radix_sort_bit_windows_and_duplicates(pol, oracle, sbit, ebit):
    n = 200003
    k = np.random.RandomState(sbit * 40 + ebit).randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    v = np.arange(n, dtype=np.int32)
    ko = torch.empty(n, dtype=torch.uint32, device="cuda"); vo = torch.empty(n, dtype=torch.int32, device="cuda")
    pol.radix_sort_pair(dev(k), dev(v), ko, vo, sbit=sbit, ebit=ebit, kind="u32")
    eo, ev = oracle.radix_sort_pair("u32", k, v, sbit, ebit)
    assert np.array_equal(ko.cpu().numpy(), eo) and np.array_equal(vo.cpu().numpy(), ev)


def test_radix_sort_skewed_keys(pol, oracle):
    """block-key-like input: few distinct, long runs (the MPM binning distribution) and all-equal keys."""
    n = 500000
    k = np.repeat(np.arange(n // 500, dtype=np.uint32)[::-1], 500)[:n].copy()
    v = np.arange(n, dtype=np.int32)
    for keys in (k, np.full(n, 77, np.uint32)):
        ko = torch.empty(n, dtype=torch.uint32, device="cuda"); vo = torch.empty(n, dtype=torch.int32, device="cuda")
        pol.radix_sort_pair(dev(keys), dev(v), ko, vo, kind="u32")
        eo, ev = oracle.radix_sort_pair("u32", keys, v)
        assert np.array_equal(ko.cpu().numpy(), eo) and np.array_equal(vo.cpu().numpy(), ev)


def test_large_sort_scan_properties(pol):
    """2^26 elements: size-independent properties (sortedness, permutation, checksum; scan end == reduce)."""
    n = 1 << 26
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    k = torch.randint(0, 2 ** 31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
    v = torch.arange(n, device="cuda", dtype=torch.int32)
    ko = torch.empty_like(k); vo = torch.empty_like(v)
    pol.radix_sort_pair(k, v, ko, vo, kind="i32")
    assert bool((ko[1:] >= ko[:-1]).all())
    assert bool((k[vo.long()] == ko).all())                       # values follow their keys
    eq = ko[1:] == ko[:-1]
    assert bool((vo[1:][eq] > vo[:-1][eq]).all())                 # stability
    assert int(vo.long().sum().item()) == n * (n - 1) // 2        # permutation checksum
    ones = torch.ones(n, dtype=torch.int32, device="cuda")
    out = torch.empty_like(ones)
    pol.inclusive_scan(ones, out)
    assert bool((out == torch.arange(1, n + 1, device="cuda", dtype=torch.int32)).all())
    r = torch.zeros(1, dtype=torch.int32, device="cuda")
    pol.reduce(ones, r, "sum")
    assert r.item() == n


def test_policy_object_abi():
    """reference-style object ABI: policy__b200 / <op>__b200_int_1 (py_interop/cuda/ExecutionPolicy.cpp)."""
    import ctypes as C
    from zpc_b200 import api
    L = api.lib()
    p = C.c_void_p(L.policy__b200())
    L.policy_set__b200(p, 0, None, 1)
    a = np.random.RandomState(1).randint(-50, 50, 9999).astype(np.int32)
    d = dev(a); out = torch.zeros(1, dtype=torch.int32, device="cuda"); sc = torch.empty_like(d)
    first, last = api.port(d, 0), api.port(d, a.size)
    L.reduce_sum__b200_int_1(p, first, last, api.port(out))
    assert out.item() == int(a.sum())
    L.exclusive_scan_sum__b200_int_1(p, first, last, api.port(sc))
    assert np.array_equal(sc.cpu().numpy(), np.concatenate([[0], np.cumsum(a)[:-1]]).astype(np.int32))
    vo = torch.empty_like(d)
    L.radix_sort_pair__b200_int_1(p, first, api.port(torch.arange(a.size, dtype=torch.int32, device="cuda")), api.port(sc),
                                  api.port(vo), C.c_size_t(a.size))
    assert np.array_equal(sc.cpu().numpy(), np.sort(a, kind="stable"))
    assert np.array_equal(vo.cpu().numpy(), np.argsort(a, kind="stable").astype(np.int32))
    assert L.policy_last_error__b200(p) == 0
    L.del_policy__b200(p)


# ---- merge_sort(_pair): stable, in place, {i32, f32, f64} keys (ExecutionPolicy.cuh:686-760) ---------------------------
MS_DT = {"i32": (np.int32, torch.int32), "f32": (np.float32, torch.float32), "f64": (np.float64, torch.float64),
         "u32": (np.uint32, torch.uint32), "i64": (np.int64, torch.int64), "u64": (np.uint64, torch.uint64)}


@pytest.mark.parametrize("kind", list(MS_DT))
@pytest.mark.parametrize("n", [0, 1, 2, 33, 4097, 300007])
def test_merge_sort_pair_matches_oracle(pol, oracle, kind, n):
    npdt, _ = MS_DT[kind]
    rs = np.random.RandomState(n + 3)
    k = rs.randint(-200, 200, size=n).astype(npdt)          # many duplicates: the value order proves stability
    if kind in ("u32", "i64", "u64"):                       # integers: the high bits take part (negatives wrap to the top of unsigned)
        k = (k.astype(np.int64) * (1 << (20 if kind == "u32" else 40))).astype(npdt)
    elif kind != "i32":
        k = (k * npdt(0.37)).astype(npdt)
        if n > 40:
            k[5] = -0.0; k[9] = 0.0; k[17] = -0.0; k[30] = np.inf; k[31] = -np.inf   # +-0 equal under <
    v = np.arange(n, dtype=np.int32)
    dk, dv = dev(k), dev(v)
    pol.merge_sort_pair(dk, dv)
    ko, vo = oracle.merge_sort_pair(kind, k, v)
    assert np.array_equal(dv.cpu().numpy(), vo)
    assert np.array_equal(dk.cpu().numpy().view(np.uint8), ko.view(np.uint8))
    dk2 = dev(k)
    pol.merge_sort(dk2)
    assert np.array_equal(dk2.cpu().numpy().view(np.uint8), ko.view(np.uint8))


def test_merge_sort_pair_on_tilevector_channels(pol, oracle):
    """keys and values living in channels of an AoSoA TileVector (iterator ports), sorted in place"""
    from zpc_b200 import api
    n = 5003
    rs = np.random.RandomState(1)
    k = rs.uniform(-10, 10, n).astype(np.float32).round(1)
    tv = api.TileVector(n, 5, 32)
    tv.set_channel(2, dev(k))
    vals = dev(np.arange(n, dtype=np.int32))
    pol.merge_sort_pair((tv, 2), vals)
    ko, vo = oracle.merge_sort_pair("f32", k, np.arange(n, dtype=np.int32))
    assert np.array_equal(tv.channel(2)[:, 0].cpu().numpy(), ko) and np.array_equal(vals.cpu().numpy(), vo)


@pytest.mark.parametrize("n", [0, 1, 33, 4097, 1000003])
def test_f64_scan_and_reduce(pol, oracle, n):
    """integer-valued doubles: every partial sum is exact, so any association gives the oracle's left fold bit for bit"""
    a = np.random.RandomState(n).randint(-1000, 1000, size=n).astype(np.float64)
    d = dev(a)
    out = torch.empty_like(d)
    pol.exclusive_scan(d, out)
    assert np.array_equal(out.cpu().numpy(), oracle.scan("exclusive", "f64", a))
    pol.inclusive_scan(d, out)
    assert np.array_equal(out.cpu().numpy(), oracle.scan("inclusive", "f64", a))
    r = torch.zeros(1, dtype=torch.float64, device="cuda")
    for op in ("sum", "min", "max"):
        pol.reduce(d, r, op)
        assert r.cpu().numpy()[0] == oracle.reduce(op, "f64", a), (op, n)


def _chunks(n, step=1 << 27):
    for a in range(0, n, step):
        yield a, min(a + step, n)


def test_one_billion_keys_sort_scan_reduce_properties(pol):
    """C5 upper end (BASELINE configs[4] "1M-1B keys"): n = 2^30 — the API limit, where an all-one-digit pass drives a look-back
    prefix to exactly 2^30 (csrc/prims.cu: RS_VAL_MASK) — pairs sorted with every key sharing its top byte, then sortedness, stability,
    keys-follow-values and the permutation checksum, verified in slices; exclusive scan of ones and reduce at the same size."""
    n = 1 << 30
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    k = torch.randint(0, 1 << 24, (n,), device="cuda", dtype=torch.int32, generator=g)
    k |= 0x35000000                                                # one digit value in the fourth pass: its prefix reaches n
    v = torch.arange(n, device="cuda", dtype=torch.int32)
    ko = torch.empty_like(k); vo = torch.empty_like(v)
    pol.radix_sort_pair(k, v, ko, vo, kind="i32")
    total = 0
    for a, b in _chunks(n):
        e = min(b + 1, n)
        kk, vv = ko[a:e], vo[a:e]
        assert bool((kk[1:] >= kk[:-1]).all())
        eq = kk[1:] == kk[:-1]
        assert bool((vv[1:][eq] > vv[:-1][eq]).all())              # stability
        assert bool((k[vo[a:b].long()] == ko[a:b]).all())           # values follow their keys
        total += int(vo[a:b].long().sum().item())
    assert total == n * (n - 1) // 2                               # a permutation
    del ko, vo, v
    ones = k; ones.fill_(1)
    out = torch.empty_like(ones)
    pol.exclusive_scan(ones, out)
    for a, b in _chunks(n):
        assert bool((out[a:b] == torch.arange(a, b, device="cuda", dtype=torch.int32)).all())
    r = torch.zeros(1, dtype=torch.int32, device="cuda")
    pol.reduce(ones, r, "sum")
    assert r.item() == n


@pytest.mark.parametrize("kind,ebit", [("u32", 12), ("u32", 20), ("u32", 24), ("u64", 64), ("u64", 40)])
def test_block_key_style_sorts_at_2_26(pol, kind, ebit):
    """12- / 20- / 24-bit restricted u32 keys through `ebit` (what block-key and binning sorts use) and u64 keys, 2^26
    pairs: result equals torch's stable sort of the masked keys, index for index"""
    n = 1 << 26
    g = torch.Generator(device="cuda"); g.manual_seed(ebit)
    if kind == "u32":
        k = torch.randint(0, 2 ** 31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
        masked = (k & ((1 << ebit) - 1)).to(torch.int64)
        keys = k.view(torch.uint32)
    else:
        k = torch.randint(0, 2 ** 62, (n,), device="cuda", dtype=torch.int64, generator=g)
        masked = k if ebit == 64 else (k & ((1 << ebit) - 1))
        keys = k.view(torch.uint64)
    v = torch.arange(n, device="cuda", dtype=torch.int32)
    ko = torch.empty_like(keys); vo = torch.empty_like(v)
    pol.radix_sort_pair(keys, v, ko, vo, kind=kind, sbit=0, ebit=ebit)
    want_k, want_i = torch.sort(masked, stable=True)
    assert bool((vo.long() == want_i).all())                       # index-exact (stable)
    got = ko.view(torch.int32 if kind == "u32" else torch.int64)
    assert bool((got == k[want_i]).all())                          # whole keys travel, only the window orders them

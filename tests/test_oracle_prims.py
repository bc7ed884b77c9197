"""Serial-primitive oracle against independent numpy statements, including the reference's own test
(test/utils/parallel_primitives.hpp:9-32: integer reduce with getmax/getmin/plus on sizes
1, 2, 7, 16, 128, 1024, 2 000 000 — exact equality)."""
import numpy as np
import pytest

REF_TEST_SIZES = [1, 2, 7, 16, 128, 1024, 2000000]


@pytest.mark.parametrize("n", REF_TEST_SIZES)
def test_reference_reduce_cases(oracle, n):
    rs = np.random.RandomState(n)
    a = rs.randint(-2 ** 20, 2 ** 20, size=n).astype(np.int32)   # reference uses rand(); any ints do
    assert oracle.reduce("max", "i32", a) == a.max()
    assert oracle.reduce("min", "i32", a) == a.min()
    assert oracle.reduce("sum", "i32", a) == np.int32(a.sum(dtype=np.int64) & 0xFFFFFFFF if False else a.sum(dtype=np.int32))


@pytest.mark.parametrize("kind,dt", [("u32", np.uint32), ("i32", np.int32), ("u64", np.uint64)])
@pytest.mark.parametrize("n", [0, 1, 33, 5000])
def test_radix_sort_pair_is_stable_sort(oracle, kind, dt, n):
    rs = np.random.RandomState(n + 1)
    k = rs.randint(0, 50, size=n).astype(dt) if n else np.zeros(0, dt)   # many duplicates: stability matters
    if kind == "i32":
        k = (k.astype(np.int64) - 25).astype(np.int32)
    v = np.arange(n, dtype=np.int32)
    ko, vo = oracle.radix_sort_pair(kind, k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(ko, k[order]) and np.array_equal(vo, v[order])
    assert np.array_equal(oracle.radix_sort(kind, k), k[order])


def test_radix_sort_bit_window(oracle):
    rs = np.random.RandomState(5)
    k = rs.randint(0, 2 ** 32, size=3000, dtype=np.uint64).astype(np.uint32)
    v = np.arange(3000, dtype=np.int32)
    ko, vo = oracle.radix_sort_pair("u32", k, v, 6, 24)
    sub = (k >> np.uint32(6)) & np.uint32((1 << 18) - 1)
    order = np.argsort(sub, kind="stable")
    assert np.array_equal(ko, k[order]) and np.array_equal(vo, v[order])


@pytest.mark.parametrize("kind,dt", [("i32", np.int32), ("u32", np.uint32), ("i64", np.int64)])
def test_scans(oracle, kind, dt):
    a = np.random.RandomState(2).randint(0, 100, size=10001).astype(dt)
    inc = np.cumsum(a, dtype=dt)
    assert np.array_equal(oracle.scan("inclusive", kind, a), inc)
    assert np.array_equal(oracle.scan("exclusive", kind, a), np.concatenate([[0], inc[:-1]]).astype(dt))
    assert oracle.scan("exclusive", kind, a[:0]).size == 0


def test_float_reduce_is_left_fold(oracle):
    a = np.random.RandomState(4).uniform(-1, 1, 1000).astype(np.float32)
    acc = np.float32(0)
    for x in a:
        acc = np.float32(acc + x)
    assert oracle.reduce("sum", "f32", a) == acc
    assert oracle.reduce("max", "f32", a) == a.max() and oracle.reduce("min", "f32", a) == a.min()


@pytest.mark.parametrize("kind,dt", [("i32", np.int32), ("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("n", [0, 1, 2, 33, 4097])
def test_merge_sort_pair_is_the_stable_sort(oracle, kind, dt, n):
    rs = np.random.RandomState(n + 7)
    k = (rs.randint(-20, 20, size=n)).astype(dt)
    if kind != "i32" and n > 4:
        k[1] = -0.0; k[3] = 0.0                     # equal under operator<: input order must survive
    v = np.arange(n, dtype=np.int32)
    ko, vo = oracle.merge_sort_pair(kind, k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(vo, v[order]) and np.array_equal(ko.view(np.uint8), k[order].view(np.uint8))

"""The C-ABI library loads and exports every symbol include/zpcb200.h declares (no compute calls: this
runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zpcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    # expand the declaration macros the header uses
    for S in ("i32", "u32", "i64", "f32"):
        for op in ("reduce_sum", "reduce_min", "reduce_max", "exclusive_scan_sum", "inclusive_scan_sum"):
            names.add("zpcb200_%s_%s" % (op, S))
    for S in ("u32", "i32", "u64"):
        names.add("zpcb200_radix_sort_pair_" + S)
        names.add("zpcb200_radix_sort_" + S)
    for T in ("int", "float"):
        for op in ("reduce_sum", "reduce_min", "reduce_max", "exclusive_scan_sum", "inclusive_scan_sum"):
            names.add("%s__b200_%s_1" % (op, T))
    body = "\n".join(l for l in src.splitlines() if not l.rstrip().endswith("\\") and not l.startswith("#"))
    for m in re.finditer(r"\b([a-zA-Z_][a-zA-Z0-9_]*)\s*\(", body):
        n = m.group(1)
        if n.startswith("zpcb200_") or n.endswith("__b200") or "__b200_" in n:
            names.add(n)
    names = {n for n in names if "##" not in n and not n.startswith("ZPCB200_")}
    return sorted(names)


def test_library_exports_header():
    from zpc_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.zpcb200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.zpcb200_version()


def test_struct_layouts_match_header():
    from zpc_b200 import api
    assert ctypes.sizeof(api.zpc_port) == 24
    assert ctypes.sizeof(api.zpc_particles_view) == 72
    assert ctypes.sizeof(api.zpc_hashtable_view) == 48
    assert ctypes.sizeof(api.zpc_grids_view) == 24
    assert ctypes.sizeof(api.zpc_tilevector_view) == 24
    assert ctypes.sizeof(api.zpc_bins_view) == 80
    assert ctypes.sizeof(api.zpc_fixed_corotated) == 20


def test_size_queries_without_gpu():
    """temp == NULL size queries are pure host arithmetic."""
    from zpc_b200 import api
    L = api.lib()
    nb = ctypes.c_size_t(0)
    none = api.zpc_port(None, 0, 0, 0, 1)
    assert L.zpcb200_radix_sort_pair_u32(None, ctypes.byref(nb), none, none, none, none, ctypes.c_size_t(1 << 20), 0, 32, None) == 0
    assert nb.value >= 2 * 4 * (1 << 20)
    assert L.zpcb200_exclusive_scan_sum_i32(None, ctypes.byref(nb), none, none, ctypes.c_size_t(1 << 20), None) == 0
    assert 0 < nb.value < (1 << 20)
    assert L.zpcb200_radix_sort_u32(None, ctypes.byref(nb), none, none, ctypes.c_size_t(1 << 31), 0, 32, None) == -3


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "libzpcoracle" not in text and "libzpcref" not in text, f

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
    for S in ("i32", "u32", "i64", "f32", "f64"):
        for op in ("reduce_sum", "reduce_min", "reduce_max", "exclusive_scan_sum", "inclusive_scan_sum"):
            names.add("zpcb200_%s_%s" % (op, S))
    for S in ("u32", "i32", "u64"):
        names.add("zpcb200_radix_sort_pair_" + S)
        names.add("zpcb200_radix_sort_" + S)
    for S in ("i32", "f32", "f64"):
        names.add("zpcb200_merge_sort_pair_" + S)
        names.add("zpcb200_merge_sort_" + S)
    for T in ("int", "float", "double"):
        for op in ("reduce_sum", "reduce_min", "reduce_max", "exclusive_scan_sum", "inclusive_scan_sum", "merge_sort", "merge_sort_pair"):
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
    assert ctypes.sizeof(api.zpc_bins_view) == 88
    assert ctypes.sizeof(api.zpc_fixed_corotated) == 20
    assert ctypes.sizeof(api.zpc_vonmises_fixed_corotated) == 24
    assert ctypes.sizeof(api.zpc_equation_of_state) == 24
    assert ctypes.sizeof(api.zpc_drucker_prager) == 44
    assert ctypes.sizeof(api.zpc_nacc) == 40
    assert ctypes.sizeof(api.zpc_collider) == 32 + 4 * (3 + 3 + 9 + 3 + 2)
    assert ctypes.sizeof(api.zpc_bht_view) == 4 * 8 + 8 + 2 * 8 + 24
    assert ctypes.sizeof(api.zpc_sparsegrid_view) == ctypes.sizeof(api.zpc_bht_view) + 8 + 8 + 4 + 64 + 4


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


def test_new_entries_host_side_behaviour():
    """size queries and argument checks of the entries added for the SparseGrid variant, merge sort and the kernel tuning
    hook — none of them launches anything"""
    from zpc_b200 import api
    L = api.lib()
    nb = ctypes.c_size_t(0)
    none = api.zpc_port(None, 0, 0, 0, 1)
    for S, ksz in (("i32", 4), ("f32", 4), ("f64", 8)):
        assert getattr(L, "zpcb200_merge_sort_pair_" + S)(None, ctypes.byref(nb), none, none, ctypes.c_size_t(1 << 20), None) == 0
        assert nb.value >= (2 * ksz + 2 * 4 + ksz + 4) * (1 << 20)          # image, index (x2), gathered keys and values
        assert getattr(L, "zpcb200_merge_sort_" + S)(None, ctypes.byref(nb), none, ctypes.c_size_t(0), None) == 0
    assert L.zpcb200_reduce_sum_f64(None, ctypes.byref(nb), none, none, ctypes.c_size_t(1 << 20), None) == 0
    # kernel tuning hook: validated values only, -1 keeps
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    assert L.zpcb200_get_tuning(ctypes.byref(a), ctypes.byref(b)) == 0 and (a.value, b.value) == (4, 1)
    assert L.zpcb200_set_tuning(7, -1) == -1 and L.zpcb200_set_tuning(2, -1) == -1 and L.zpcb200_set_tuning(-1, 5) == -1
    assert L.zpcb200_set_tuning(3, 0) == 0
    L.zpcb200_get_tuning(ctypes.byref(a), ctypes.byref(b))
    assert (a.value, b.value) == (3, 0)
    assert L.zpcb200_set_tuning(-1, 128) == 0 and L.zpcb200_set_tuning(4, -1) == 0
    L.zpcb200_get_tuning(ctypes.byref(a), ctypes.byref(b))
    assert (a.value, b.value) == (4, 128)
    assert L.zpcb200_set_tuning(6, -1) == 0 and L.zpcb200_set_tuning(8, -1) == -1   # 6 = the plane sweep
    L.zpcb200_get_tuning(ctypes.byref(a), ctypes.byref(b))
    assert (a.value, b.value) == (6, 128)
    assert L.zpcb200_set_tuning(4, 1) == 0
    # a SparseGrid with a rotated / translated transform is refused by the MPM functors before anything is launched
    sg = api.SparseGrid(7, 64, device="cpu")
    sg.scale(0.1)
    sg.translate([0.5, 0.0, 0.0])
    e = (ctypes.c_float * 3)(0.0, -9.8, 0.0)
    assert L.zpcb200_sg_grid_update(sg.view(), ctypes.c_float(1e-4), e, 2, None, None) == -1      # bad mode / NULL maxVel
    pv = api.zpc_particles_view(None, None, None, None, None, None, None, None, 0)
    assert L.zpcb200_sg_g2p_apic(pv, sg.view(), ctypes.c_float(1e-4), None) == -3                   # ZPCB200_E_UNSUPPORTED
    nbytes = ctypes.c_size_t(0)
    assert L.zpcb200_sg_partition_build(None, ctypes.byref(nbytes), none, ctypes.c_size_t(0), sg.view(), 0, 2, None, None) == -3
    sg2 = api.SparseGrid(7, 64, device="cpu")
    sg2.scale(0.1)
    assert L.zpcb200_sg_partition_build(None, ctypes.byref(nbytes), none, ctypes.c_size_t(0), sg2.view(), 0, 2, None, None) == 0
    assert nbytes.value > 0
    # static collider helper
    L.zpcb200_collider_static.restype = api.zpc_collider
    o = (ctypes.c_float * 3)(0.0, 0.3, 0.0)
    nrm = (ctypes.c_float * 3)(0.0, 1.0, 0.0)
    c = L.zpcb200_collider_static(0, 2, o, nrm)
    assert (c.geometry, c.type, c.s, c.dsdt) == (0, 2, 1.0, 0.0) and list(c.R) == [1, 0, 0, 0, 1, 0, 0, 0, 1] and list(c.b) == [0, 0, 0]


def test_plastic_model_entries_check_their_arguments():
    """zpcb200_p2g_apic_drucker_prager / _nacc refuse a particle view without logJp and a non-3D NACC model before
    anything is launched; an empty range is a no-op"""
    from zpc_b200 import api
    L = api.lib()
    dp, nacc = api.model_drucker_prager(1e-6), api.model_nacc(1e-6)
    assert abs(dp.yieldSurface - 0.816496580927726 * 2 * 0.5 / 2.5) < 1e-7 and dp.volumeCorrection == 1 and nacc.hardeningOn == 1
    tv = api.zpc_hashtable_view(1, 1, 1, 1, 16, 1)          # non-null dummies: the checks below fail before any use
    gv = api.zpc_grids_view(1, 1, 7, 0.1)
    no_logjp = api.zpc_particles_view(1, 1, 1, None, None, 1, 1, None, 10)
    empty = api.zpc_particles_view(None, None, None, None, None, None, None, None, 0)
    for fn, model in ((L.zpcb200_p2g_apic_drucker_prager, dp), (L.zpcb200_p2g_apic_nacc, nacc)):
        assert fn(no_logjp, tv, gv, ctypes.c_float(1e-4), model, None) == -1
        assert fn(empty, tv, gv, ctypes.c_float(1e-4), model, None) == 0
        assert fn(empty, tv, api.zpc_grids_view(1, 1, 4, 0.1), ctypes.c_float(1e-4), model, None) == -1
    bad = api.model_nacc(1e-6)
    bad.dim = 2
    assert L.zpcb200_p2g_apic_nacc(empty, tv, gv, ctypes.c_float(1e-4), bad, None) == -1


def test_lbvh_entries_host_side_behaviour():
    from zpc_b200 import api
    L = api.lib()
    assert ctypes.sizeof(api.zpc_lbvh_view) == 40
    nb = ctypes.c_size_t(0)
    v = api.zpc_lbvh_view(None, None, None, None, None)
    assert L.zpcb200_lbvh_build(None, ctypes.byref(nb), None, ctypes.c_size_t(1 << 20), v, 1, None) == 0
    assert nb.value > 13 * 4 * (1 << 20)               # nine index arrays, codes and ids twice, flags, plus sort scratch
    assert L.zpcb200_lbvh_build(None, ctypes.byref(nb), None, ctypes.c_size_t(2), v, 1, None) == 0 and nb.value == 256
    assert L.zpcb200_lbvh_build(None, ctypes.byref(nb), None, ctypes.c_size_t((1 << 30) + 1), v, 1, None) == -3
    assert L.zpcb200_lbvh_refit(None, ctypes.byref(nb), None, ctypes.c_size_t(1000), v, None) == 0 and nb.value >= 8000
    assert L.zpcb200_lbvh_build(None, None, None, ctypes.c_size_t(10), v, 1, None) == -1
    small = ctypes.c_size_t(16)
    buf = ctypes.create_string_buffer(16)
    assert L.zpcb200_lbvh_build(buf, ctypes.byref(small), None, ctypes.c_size_t(1000), v, 1, None) == -2


def test_index_buckets_entry_host_side_behaviour():
    from zpc_b200 import api
    L = api.lib()
    nb = ctypes.c_size_t(0)
    none = api.zpc_port(None, 0, 0, 0, 3)
    tv = api.zpc_hashtable_view(None, None, None, None, 16 * 1024, None)
    args = (none, ctypes.c_size_t(1000), ctypes.c_float(0.1), ctypes.c_float(0.5), tv, None, None, None, None, None)
    assert L.zpcb200_index_buckets_build(None, ctypes.byref(nb), *args) == 0 and nb.value > 3 * 4 * 1000
    buf = ctypes.create_string_buffer(nb.value)
    assert L.zpcb200_index_buckets_build(buf, ctypes.byref(nb), *args) == -1        # null table arrays are refused before any launch
    assert L.zpcb200_index_buckets_build(None, None, *args) == -1
    bad = api.zpc_hashtable_view(None, None, None, None, 0, None)
    assert L.zpcb200_index_buckets_build(None, ctypes.byref(nb), none, ctypes.c_size_t(10), ctypes.c_float(0.1), ctypes.c_float(0.5), bad,
                                         None, None, None, None, None) == -1


def test_g2p2g_entry_checks_its_arguments():
    from zpc_b200 import api
    L = api.lib()
    tv = api.zpc_hashtable_view(1, 1, 1, 1, 16, 1)
    m = api.model_fcr(1e-6)
    empty = api.zpc_particles_view(None, None, None, None, None, None, None, None, 0)
    some = api.zpc_particles_view(1, 1, 1, None, None, 1, 1, None, 10)
    f = ctypes.c_float
    assert L.zpcb200_g2p2g_apic(empty, tv, f(0.1), f(1e-4), 0, ctypes.byref(m), None, None, None) == 0          # empty range: no-op
    assert L.zpcb200_g2p2g_apic(some, tv, f(0.1), f(1e-4), 0, ctypes.byref(m), None, None, None) == -1         # null DOF vectors
    assert L.zpcb200_g2p2g_apic(some, tv, f(0.1), f(1e-4), 7, ctypes.byref(m), None, None, None) == -1         # unknown model kind
    assert L.zpcb200_g2p2g_apic(some, tv, f(0.1), f(1e-4), 0, None, None, None, None) == -1
    nacc = api.model_nacc(1e-6)
    assert L.zpcb200_g2p2g_apic(some, tv, f(0.1), f(1e-4), 3, ctypes.byref(nacc), ctypes.c_void_p(8), ctypes.c_void_p(8), None) == -1   # no logJp


def test_grid_momentum_entries_check_their_arguments():
    from zpc_b200 import api
    L = api.lib()
    g = api.zpc_grids_view(1, 100, 7, 0.1)
    tv = api.zpc_hashtable_view(1, 1, 1, 1, 16, 1)
    one = ctypes.c_void_p(8)
    for m_chn, mv_chn in ((-1, 1), (0, 5), (7, 1), (2, 1), (0, -1)):     # outside the 7 channels, or mass inside the momentum range
        assert L.zpcb200_grid_momentum_to_velocity(g, one, m_chn, mv_chn, one, None) == -1, (m_chn, mv_chn)
    assert L.zpcb200_grid_momentum_to_velocity(g, None, 0, 1, one, None) == -1
    assert L.zpcb200_grid_momentum_to_velocity(g, one, 0, 1, None, None) == -1
    assert L.zpcb200_grid_angular_momentum(g, tv, 0, 1, None, None) == -1
    assert L.zpcb200_grid_angular_momentum(g, tv, 0, 5, one, None) == -1
    assert L.zpcb200_grid_angular_momentum(api.zpc_grids_view(None, 100, 7, 0.1), tv, 0, 1, one, None) == -1


def test_overlay_binds_the_reference_headers_to_the_library():
    """oracle/_ref/libzpcref_cuda.so = the unmodified reference (CUDA backend) + include/zpcb200/zs_overlay.cuh.  It cannot be loaded
    without a driver, but its dynamic symbol table shows what the overlay resolved to: generic zs::radix_sort_pair / exclusive_scan /
    reduce calls with b200_exec() and the five functor launches must reference zpcb200_* entries that libzpcb200.so exports."""
    import subprocess
    so = os.path.join(ROOT, "oracle", "_ref", "libzpcref_cuda.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libzpcref_cuda.so not built (make -C oracle refcuda where /root/reference is mounted)")
    from zpc_b200 import build
    lib = ctypes.CDLL(build.build())
    out = subprocess.check_output(["nm", "-D", so], text=True)
    undefined = {l.split()[-1] for l in out.splitlines() if " U zpcb200_" in l}
    want = {"zpcb200_partition_build", "zpcb200_clean_grid", "zpcb200_p2g_apic_fcr", "zpcb200_grid_update", "zpcb200_g2p_apic",
            "zpcb200_radix_sort_pair_u32", "zpcb200_exclusive_scan_sum_i32", "zpcb200_reduce_sum_i32", "zpcb200_reduce_max_i32"}
    assert want <= undefined, want - undefined
    assert all(hasattr(lib, n) for n in undefined)
    defined = {l.split()[-1] for l in out.splitlines() if " T zpcrefcuda_" in l}
    assert {"zpcrefcuda_mpm_p2g", "zpcrefcuda_overlay_p2g", "zpcrefcuda_overlay_prims"} <= defined


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "libzpcoracle" not in text and "libzpcref" not in text, f

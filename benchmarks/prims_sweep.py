"""C5: radix_sort_pair / exclusive_scan / reduce sweep on one GPU, GB/s against the algorithmic byte counts of
SURVEY §8(d), next to torch's CUB-backed ops (torch.sort / cumsum / sum — the library path the reference's
CudaExecutionPolicy wraps, without zpc's extra copy kernels).  python benchmarks/prims_sweep.py [--max-log2 30]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zpc_b200 import api  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_prims(lg, peak, with_torch=False, extras=False, iters=10):
    """one row of the sweep: n = 2^lg i32 keys (mt19937-free device generator, seed 12345), i32 values = iota"""
    n = 1 << lg
    pol = api.cuda_exec().sync(False)
    g = torch.Generator(device="cuda"); g.manual_seed(12345)
    keys = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
    vals = torch.arange(n, device="cuda", dtype=torch.int32)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    r = torch.zeros(1, device="cuda", dtype=torch.int32)
    it = max(3, iters if lg <= 28 else iters // 3)
    t_sort = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32"), it)
    t_sort24 = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32", sbit=0, ebit=24), it)
    t_scan = timeit(lambda: pol.exclusive_scan(vals, vo), it)
    t_red = timeit(lambda: pol.reduce(vals, r, "sum"), it)
    row = dict(log2n=lg, n=n, sort_pair_ms=t_sort, sort_pair_gbps=68 * n / t_sort / 1e6, sort_pair_gkeys=n / t_sort / 1e6,
               sort_pair_24bit_ms=t_sort24, sort_pair_24bit_gbps=52 * n / t_sort24 / 1e6,
               scan_ms=t_scan, scan_gbps=8 * n / t_scan / 1e6, reduce_ms=t_red, reduce_gbps=4 * n / t_red / 1e6,
               bytes_per_key=dict(sort_pair=68, sort_pair_24bit=52, scan=8, reduce=4))
    if extras:
        row["sort_pair_12bit_ms"] = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32", sbit=0, ebit=12), it)
        row["sort_pair_20bit_ms"] = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32", sbit=0, ebit=20), it)
        row["sort_keys_ms"] = timeit(lambda: pol.radix_sort(keys, ko, kind="i32"), it)
        row["sort_keys_gbps"] = 36 * n / row["sort_keys_ms"] / 1e6
        if lg <= 29:
            k64 = torch.randint(0, 2 ** 62, (n,), device="cuda", dtype=torch.int64, generator=g).view(torch.uint64)
            k64o = torch.empty_like(k64)
            row["sort_pair_u64_ms"] = timeit(lambda: pol.radix_sort_pair(k64, vals, k64o, vo, kind="u64"), it)
            row["sort_pair_u64_gbps"] = (8 + 8 * 2 * 12) * n / row["sort_pair_u64_ms"] / 1e6
            del k64, k64o
    if with_torch:
        row["torch_sort_ms"] = timeit(lambda: torch.sort(keys, stable=True), it)
        row["torch_cumsum_ms"] = timeit(lambda: torch.cumsum(vals, 0, dtype=torch.int32), it)
        row["torch_sum_ms"] = timeit(lambda: torch.sum(vals, dtype=torch.int32), it)
    row["sort_frac"] = row["sort_pair_gbps"] / peak
    row["scan_frac"] = row["scan_gbps"] / peak
    row["reduce_frac"] = row["reduce_gbps"] / peak
    del keys, vals, ko, vo
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--min-log2", type=int, default=20)
    args = ap.parse_args()
    peak = 6536.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        row = time_prims(lg, peak, with_torch=True, extras=True)
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()

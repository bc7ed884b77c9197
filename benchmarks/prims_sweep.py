"""C5: radix_sort_pair / exclusive_scan / reduce sweep on one GPU, GB/s against the algorithmic byte counts of
SURVEY §8(d), next to torch's CUB-backed ops (torch.sort / cumsum / sum — the library path the reference's
CudaExecutionPolicy wraps, without zpc's extra copy kernels).  python benchmarks/prims_sweep.py [--max-log2 28]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=28)
    ap.add_argument("--min-log2", type=int, default=20)
    args = ap.parse_args()
    peak = 6536.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    pol = api.cuda_exec().sync(False)
    rows = []
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg
        g = torch.Generator(device="cuda"); g.manual_seed(12345)
        keys = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
        vals = torch.arange(n, device="cuda", dtype=torch.int32)
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        ones = torch.ones(n, device="cuda", dtype=torch.int32)
        out = torch.empty_like(ones)
        r = torch.zeros(1, device="cuda", dtype=torch.int32)
        t_sort = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32"))
        t_sort24 = timeit(lambda: pol.radix_sort_pair(keys, vals, ko, vo, kind="i32", sbit=0, ebit=24))
        t_scan = timeit(lambda: pol.exclusive_scan(ones, out))
        t_red = timeit(lambda: pol.reduce(ones, r, "sum"))
        t_tsort = timeit(lambda: torch.sort(keys, stable=True))
        t_tscan = timeit(lambda: torch.cumsum(ones, 0, dtype=torch.int32))
        t_tsum = timeit(lambda: torch.sum(ones, dtype=torch.int32))
        row = dict(log2n=lg, sort_pair_ms=t_sort, sort_pair_gbps=68 * n / t_sort / 1e6, sort_pair_gkeys=n / t_sort / 1e6,
                   sort_pair_24bit_ms=t_sort24, sort_pair_24bit_gbps=52 * n / t_sort24 / 1e6,
                   torch_sort_ms=t_tsort, scan_ms=t_scan, scan_gbps=8 * n / t_scan / 1e6, torch_cumsum_ms=t_tscan,
                   reduce_ms=t_red, reduce_gbps=4 * n / t_red / 1e6, torch_sum_ms=t_tsum)
        row["sort_frac"] = row["sort_pair_gbps"] / peak
        row["scan_frac"] = row["scan_gbps"] / peak
        row["reduce_frac"] = row["reduce_gbps"] / peak
        rows.append(row)
        print(json.dumps(row))
    return rows


if __name__ == "__main__":
    main()

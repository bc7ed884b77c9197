#!/usr/bin/env python
"""HBM bandwidth of plain streaming torch ops at different read:write mixes (context for the roofline fractions of
write-heavy kernels such as G2P: 48 B read + 96 B written per particle).  Measurement aid only."""
import json

import torch


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    N = 1 << 30
    a32 = torch.ones(N, dtype=torch.float32, device="cuda")
    b32 = torch.empty(N, dtype=torch.float32, device="cuda")
    a16 = torch.ones(N, dtype=torch.float16, device="cuda")
    out = {}
    t = timeit(lambda: b32.fill_(1.0)); out["write_only"] = 4 * N / t / 1e6
    t = timeit(lambda: a32.sum()); out["read_only"] = 4 * N / t / 1e6
    t = timeit(lambda: b32.copy_(a32)); out["copy_1r_1w"] = 8 * N / t / 1e6
    t = timeit(lambda: b32.copy_(a16)); out["r1_w2"] = 6 * N / t / 1e6
    t = timeit(lambda: a16.copy_(a32)); out["r2_w1"] = 6 * N / t / 1e6
    print(json.dumps({k: round(v, 1) for k, v in out.items()}) + "  (GB/s)")


if __name__ == "__main__":
    main()

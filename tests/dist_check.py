"""torchrun --nproc-per-node N tests/dist_check.py : N-GPU sharded substeps against the single-GPU solver on
the same cloud (rank 0 holds both).  Exit code 0 = parity within the multi-step tolerance.  The check itself is
zpc_b200/selfcheck.py (bench.py --gpus N runs the same before it times anything).
ZPC_MIGRATE=1: after half of the substeps every particle is handed to the rank that owns its current home block
(DistMpmSolver.migrate; ownership = BlockOwnership over shard_by_blocks of the initial cloud) — results must not change.
ZPC_E2E=1: the substeps go through DistMpmSolver.substep_host (host buffers per rank, AoS kernels) instead.
ZPC_GRAPH=1: a re-bin cycle captured as one CUDA graph and replayed (DistMpmSolver.capture_cycle)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zpc_b200.selfcheck import multi_gpu_parity  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    r = multi_gpu_parity(migrate=os.environ.get("ZPC_MIGRATE") == "1", e2e=os.environ.get("ZPC_E2E") == "1",
                         transport=os.environ.get("ZPC_HALO", "auto"), graph=os.environ.get("ZPC_GRAPH") == "1")
    if rank == 0:
        print(("dist_check ok: " if r["ok"] else "dist_check FAILED: ") + json.dumps(r))
    dist.destroy_process_group()
    sys.exit(0 if r["ok"] else 1)


if __name__ == "__main__":
    main()

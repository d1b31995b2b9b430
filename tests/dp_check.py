"""Data-parallel equivalence on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 -m tests.dp_check

W ranks on B/W rows each == one rank on all B rows (jck_generation_b200/train/dp_selfcheck.py, which bench.py also runs at
every N > 1 and reports as `dp_check`).  Tolerances and why they are what they are: see that module's docstring."""
import sys

import torch

from jck_generation_b200 import parallel
from jck_generation_b200.train import dp_selfcheck


def main():
    comm = parallel.init_from_env()
    res = dp_selfcheck.run(comm, per_rank=16)
    if comm.rank == 0:
        print("dp_check", res, flush=True)
        assert res["ok"], res
        print("dp_check OK", flush=True)
    comm.barrier()
    if comm.world_size > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

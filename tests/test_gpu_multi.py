"""Multi-GPU checks as pytest cases: each launches its torchrun script (tests/dp_check.py, tests/comm_check.py,
tests/fid_dp_check.py) on two GPUs of the box and looks for the script's own OK line.  Skipped on a box with fewer than two
GPUs (the round-end GPU test box has one); the scripts are also run by hand under `gpurun --gpus N` (DESIGN.md sections 6, 9)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(module, nproc=2, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "-m", module]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("module,marker", [("tests.comm_check", "comm_check OK"), ("tests.dp_check", "dp_check OK"),
                                           ("tests.fid_dp_check", "fid_dp_check OK")])
def test_two_gpu_script(module, marker):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    import __graft_entry__ as entry
    entry.build()
    rc, out = _torchrun(module)
    assert marker in out, out[-3000:]

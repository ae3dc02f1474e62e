"""Multi-GPU parity on a box with >= 2 GPUs: the slab-decomposed CG / BiCGSTAB / Jacobi / Euler
(P = 2) against the single-GPU run of the same global problem (tools/dist_check.py under torchrun).
Skipped on single-GPU boxes; the host-side slab logic is covered on CPU by test_parallel_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slab_solvers_match_single_gpu():
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = "\n".join((out.stdout + out.stderr).splitlines()[-60:])
    assert out.returncode == 0 and "DIST_CHECK PASS" in out.stdout, tail

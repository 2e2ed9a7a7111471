"""Multi-GPU parity under the driver's `-m gpu` tier: self-launches torchrun on 2 GPUs when the box has them
(tests/dist_worker.py), skips on a single-GPU box.  Covers what a single process cannot: the NVLink peer-memory
exchange (csrc/exchange.cu) against NCCL and against the oracle, every `method`, arg-reductions with ties across ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_parity(nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ)
    env.pop("FLOXB_PEER_XCHG", None)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29530 + nproc), os.path.join(ROOT, "tests", "dist_worker.py")],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    tail = (p.stdout + p.stderr)[-4000:]
    assert p.returncode == 0 and "dist_worker OK" in p.stdout, tail

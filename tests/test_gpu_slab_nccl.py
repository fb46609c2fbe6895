"""Multi-process slab decomposition over NCCL (needs >= 2 GPUs; skipped on a 1-GPU box,
where test_gpu_slab.py covers the same device code through the in-process ring)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("nproc", [2, 4])
def test_slab_nccl_ring_matches_single_engine(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc),
           os.path.join(HERE, "slab_nccl_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SLAB_NCCL_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]

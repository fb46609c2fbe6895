"""Multi-process slab decomposition, one rank per GPU (needs >= 2 GPUs; skipped on a 1-GPU box,
where test_gpu_slab.py covers the same device code through the in-process ring), with both
transports of jax_sph_b200/slab.py: "direct" (pack kernels store into the ring neighbours'
symmetric-memory buffers over NVLink, flags instead of collectives) and "nccl"."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("transport", ["direct", "nccl"])
@pytest.mark.parametrize("nproc", [2, 4])
def test_slab_nccl_ring_matches_single_engine(nproc, transport):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc + (10 if transport == "nccl" else 0)),
           os.path.join(HERE, "slab_nccl_worker.py")]
    env = dict(os.environ, SPHB200_SLAB_TRANSPORT=transport)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and f"SLAB_NCCL_OK transport={transport}" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]

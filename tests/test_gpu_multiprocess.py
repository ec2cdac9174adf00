"""The fused all-reduce ACROSS PROCESSES on real peers (CUDA IPC mailboxes over NVLink): one process per GPU via torchrun, checked against
NCCL on the local sums (1e-12), against the un-sharded context, and through a whole Gauss-Newton align (tests/mp_fused_check.py).
Needs >= 2 GPUs in the box (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multiprocess.py -m gpu`); skipped on a single-GPU box, where
tests/test_gpu_fused_allreduce.py runs the same protocol between contexts of one process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_allreduce_across_processes(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "mp_fused_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    if r.returncode != 0:  # torchrun's own summary buries the worker's traceback: keep everything, show the lines that name the error
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"mp_fused_check_world{world}.log"), "w") as f:
            f.write(r.stdout + "\n==== stderr ====\n" + r.stderr)
    culprit = [ln for ln in r.stderr.splitlines() if ("Error" in ln or "assert" in ln) and "ChildFailedError" not in ln][:20]
    assert r.returncode == 0, "\n".join(culprit) + "\n" + r.stderr[-1500:]
    assert "mp_fused_check ok" in r.stdout

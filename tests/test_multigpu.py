"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): torchrun with one rank per GPU runs
tests/_mgpu_worker.py, which checks the decomposed paths against the single-GPU results."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposed_paths_match_single_gpu(cuda_device, world):
    """Grid 256^3 (hand-written transform path; the worker repeats the spectrum at 128^3 = cuFFT path)."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    env = dict(os.environ, FAVA_MGPU_N="256")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world), str(ROOT / "tests" / "_mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0 and "MGPU_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]

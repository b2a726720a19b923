"""The head-sharded path inside `pytest -m gpu`: tests/gpu_dist_parity.py is launched under torchrun with two ranks -- one
per GPU over NCCL when the box has two GPUs, else both on cuda:0 over gloo -- and must print DIST_PARITY PASS
(sharded train_phase1 / train_phase2 == single-GPU: check logs, GC on every rank, weights <= 1e-4)."""
import os
import subprocess
import sys

import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_training_equals_single_gpu(world):
    env = dict(os.environ)
    if torch.cuda.device_count() < world:
        env["DIST_ONE_GPU"] = "1"
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gpu_dist_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert f"DIST_PARITY PASS world={world}" in r.stdout, tail

"""Multi-GPU path on real devices: needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise.
The CPU-side logic of the same path is covered with gloo in tests/test_host_logic.py."""
import os
import subprocess
import sys

import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_ensemble_is_bit_identical_and_pooled_stats_allreduce():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "scripts",
                                                                               "dist_invariance.py")]
    # windows wide enough that every vector of the script — also the multi-chunk ones and the shared-covariance path's
    # moment vector — goes through the one-shot NVLink all-reduce (by default only vectors up to 1,024 doubles do)
    env = dict(os.environ, ME_PEER_WINDOW_DOUBLES="32768")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "OK" in out.stdout

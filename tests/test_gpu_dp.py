"""Multi-GPU data parallelism on real devices (skipped with fewer than 2 GPUs): G ranks reproduce the 1-rank trainer on
the global batch within 1e-5 and the replicas stay bit-identical -- with the gradient exchange inside the update kernel
over peer memory (the default) and with the NCCL all-reduce (VV_DP_MODE=nccl).  scripts/dp_check.py does the work."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
@pytest.mark.parametrize("prec", ["tf32x3", "f16x3", "bf16"])
def test_two_rank_matches_single_rank(prec, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dp_check.py"), prec]
    env = dict(os.environ, VV_DP_MODE=mode, VV_DP_TIMEOUT_MS="5000")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert "DP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert ("mode=" + mode) in r.stdout, r.stdout[-2000:]

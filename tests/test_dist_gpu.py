"""K5 on real GPUs: needs >= 2 B200s on the box (skipped on a 1-GPU lease)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused", ["1", "0"])
def test_sharded_search_matches_oracle(tss, fused):
    """fused=1: the scan's last CTA exchanges over peer memory; fused=0: ncclAllGather + merge
    kernel for every search (TSS_FUSED_XCHG=0), pending searches included."""
    ngpu = tss.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, TSS_FUSED_XCHG=fused)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "DIST_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]

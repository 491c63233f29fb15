"""The multi-GPU build on real GPUs (needs >= 2 devices; skipped on a one-GPU box): torchrun launches
tests/multigpu_check.py, which compares the index built by 2 (and 4) ranks with the oracle's."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,lo_bits", [(2, None), (2, "14"), (4, None)])
def test_index_built_by_several_gpus_matches_the_oracle(world, lo_bits):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    if lo_bits:
        env["DSMFM_POS_LO_BITS"] = lo_bits
    else:
        env["DSMFM_CHECK_LARGE"] = "1"  # 400k reads per rank: the N-GPU file equals the 1-GPU build of the same documents
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multigpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_CHECK_OK" in r.stdout

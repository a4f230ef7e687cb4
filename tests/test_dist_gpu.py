"""Multi-GPU parity of the CUDA code itself (not the numpy twin): tools/dist_check.py under torch.distributed.run, one rank per
GPU over NCCL, and tools/mg_check.py — the single-process multi-GPU path behind the unchanged C API — both against the
compiled reference.  Skipped on boxes with fewer than 2 GPUs; logs of 2- and 8-GPU runs are committed under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_check_row_partitioned(world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "dist_check.py")],
                         capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert out.returncode == 0 and "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 8])
def test_single_process_multi_gpu_behind_c_api(world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, RSVD_B200_DEVICES="0-%d" % (world - 1))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mg_check.py")], capture_output=True, text=True, timeout=1200, cwd=ROOT, env=env)
    assert out.returncode == 0 and "MG_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

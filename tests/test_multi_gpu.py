"""Multi-GPU parity (skips below 2 GPUs): the sharded prefilter and the whole sharded pipeline under
torchrun with 2 ranks, against the single-GPU result (tools/check_sharded.py,
tools/check_pipeline_sharded.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("tool", ["check_sharded.py", "check_pipeline_sharded.py"])
def test_two_ranks_match_single_gpu(tool):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", tool)],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "ALL OK" in out.stdout, out.stdout[-3000:]


def test_in_process_multi_device_entry_matches_single_gpu():
    """galah_b200_cluster_packed_multi (C ABI, one host thread per device inside the library, peer copies
    and peer-mapped K3 tables) on 2 GPUs: clusters identical to galah_b200_cluster_packed on one."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_multi_device.py"), "--genomes", "1536",
                          "--len", "500000", "--devices", "2"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "ALL OK" in out.stdout, out.stdout[-3000:]

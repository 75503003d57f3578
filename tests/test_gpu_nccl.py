"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU under
torch.distributed.run, the library's NCCL transport and the torch.distributed CALLBACK transport
against the oracle with the same x-slab decomposition."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode", ["nccl", "callback", "mailboxes", "mailboxes-particles"])
def test_two_gpu_parity(mode, cylgpu_lib):
    """nccl: ncclSend / ncclRecv only (CYLGPU_P2P=0); mailboxes: the NCCL transport with its peer-memory data path
    (the default, csrc/transport.cu); mailboxes-particles: the mailboxes for the counted particle messages only"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env["CYLGPU_P2P"] = {"mailboxes": "1", "mailboxes-particles": "particles"}.get(mode, "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "nccl_parity_worker.py"),
           "nccl" if mode.startswith("mailboxes") else mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "NCCL_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("mode", ["mailboxes", "nccl"])
def test_four_gpu_parity(mode, cylgpu_lib):
    """interior ranks: two different neighbours per rank (two mailbox links, two NCCL peers), the closing exchange of
    the communication-avoiding field phases in both directions at once, the window deck over four slabs"""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    env = dict(os.environ)
    env["CYLGPU_P2P"] = "1" if mode == "mailboxes" else "0"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(HERE, "nccl_parity_worker.py"), "nccl"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert "NCCL_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

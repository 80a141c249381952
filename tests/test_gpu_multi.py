"""GPU suite, N > 1: sharded subsets and the row-split all-reduce on 2 (or more) real GPUs against the single-GPU
result and the oracle. Skipped on a box with one GPU; tools/multi_gpu_check.py holds the assertions (it also runs
with one rank: the exchange then loops back through the rank's own mailbox)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from correlation_b200 import engine
    return engine.load_library().dic_device_count()


def _run(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_one_rank_loopback():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_and_rowsplit_on_n_gpus(n):
    if _device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    out = _run(n)
    assert out.returncode == 0 and "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]

"""GPU: the data-parallel gradient all-reduce captured inside the backward CUDA graph (NCCL, world size 1 in a subprocess)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_overlapped_grad_sync_inside_the_graph():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_overlap_worker.py")
    r = subprocess.run([sys.executable, worker], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OVERLAP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

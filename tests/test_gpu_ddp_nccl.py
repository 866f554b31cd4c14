"""GPU x 2 (NCCL): data-parallel gradient parity of the real model: the mean of the two ranks'
gradients, each on 40 rows of the config-2 batch, equals the single-process gradient on the concatenated 80 rows -- checked
against autograd through the ORACLE on all 80 rows.  Skipped on a one-GPU box (run with `gpurun --gpus 2`; the log of that run
is profiles/r2_ddp_nccl_2gpu.log)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from oracle import rfnet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_two_rank_mean_gradient_equals_single_process_gradient(tmp_path):
    from tests.test_gpu_training import _config2_batch, _oracle_xe_grads
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "ddp_grads.pt")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ddp_worker.py")
    procs = []
    for r in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, worker, out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    cfg, fc, att, labels, masks, top = _config2_batch()
    sd = O.make_state_dict(cfg, seed=1234)
    want_loss, want = _oracle_xe_grads(cfg, sd, fc, att, labels, masks, top)   # overlaps the workers
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs) and all("DDP_OK" in l for l in logs), "\n".join(l[-3000:] for l in logs)
    got = torch.load(out)
    # the margin terms are batch means too, so the mean of the two 40-row losses is the 80-row loss
    assert abs(got["loss_mean"] - want_loss) <= 1e-4 * max(1.0, abs(want_loss))
    worst = 0.0
    for k, w in want.items():
        g = got["grads"][k]
        scale = float(w.abs().max()) + 1e-6
        err = float((g - w).abs().max())
        worst = max(worst, err / scale)
        assert err / scale <= 5e-4 or err <= 1e-6, f"{k}: rel err {err / scale:.3g}"
    print(f"2-rank NCCL mean gradient vs oracle 80-row gradient: worst relative error {worst:.3g} over {len(want)} tensors")

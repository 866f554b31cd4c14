"""CPU, world_size 2, gloo: the N>1 host logic -- contiguous image sharding, the ragged caption gather
and the gradient averaging -- produces exactly what one process produces on the whole job.  The decode
itself is the oracle here (no GPU in this suite); the CUDA path uses the same dist.py functions."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rfnet_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recurrent_fusion_network_b200 import dist as D
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, n_images, seed=21)

    def decode(fcs, atts):
        seq, slp, *_ = O.sample_beam(sd, cfg, fcs, atts, beam_size=3)
        return seq, slp

    seq, slp = D.sharded_decode(decode, fc, att)
    # gradient averaging: mean over ranks, clamp after the average
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.tensor([4.0, -4.0, 0.5, 1.0, 3.0]) * (rank + 1)
    D.average_gradients([p], grad_clip=1.0, bucket_bytes=8)
    # divide=False: the SUM is left for FusedAdam(grad_scale=1/world), which divides and clamps in the optimizer pass
    q = torch.nn.Parameter(torch.zeros(3))
    q.grad = torch.tensor([1.0, -2.0, 0.25]) * (rank + 1)
    D.average_gradients([q], divide=False)
    if rank == 0:
        out_q.put((seq, slp, p.grad.clone(), q.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 8, 1])
def test_sharded_decode_equals_single_process(n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    seq, slp, grad, grad_sum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, n_images, seed=21)
    want_seq, want_slp, *_ = O.sample_beam(sd, cfg, fc, att, beam_size=3)
    assert torch.equal(seq, want_seq)
    assert torch.equal(slp, want_slp)
    # mean of (g, 2g) = 1.5 g, clamped to +-1 afterwards
    assert torch.allclose(grad, (torch.tensor([4.0, -4.0, 0.5, 1.0, 3.0]) * 1.5).clamp(-1, 1))
    assert torch.allclose(grad_sum, torch.tensor([1.0, -2.0, 0.25]) * 3)


def test_shard_range_covers_everything():
    from recurrent_fusion_network_b200.dist import shard_range
    for n in (0, 1, 5, 8, 5000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))

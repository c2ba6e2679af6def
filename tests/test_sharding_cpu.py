"""N>1 host logic on CPU: 2 processes over gloo (127.0.0.1): weight broadcast, scene sharding, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevgen_b200.sharding import broadcast_module_weights, max_over_ranks, scene_shard


def test_scene_shard_partition():
    for n, w in [(128, 8), (16, 3), (5, 8), (1, 2), (0, 4)]:
        seen = []
        for r in range(w):
            seen += list(scene_shard(n, r, w))
        assert seen == list(range(n))
        sizes = [len(scene_shard(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_view_generation.modules.losses.vqperceptual import DummyLoss
    from multi_view_generation.modules.stage1.vqgan import VQModel
    from oracle import synth
    dd = synth.vqgan_ddconfig(ch=32)
    torch.manual_seed(100 + rank)                        # different random init on every rank
    m = VQModel(dd, DummyLoss(), 64, 256, (256, 256), (16, 16), 256)
    if rank == 0:
        m.load_state_dict(synth.vqgan_state_dict(dd, seed=1, n_embed=64))
    sent = broadcast_module_weights(m, src=0, bucket_bytes=1 << 20)
    ref = synth.vqgan_state_dict(dd, seed=1, n_embed=64)
    ok = all(torch.equal(m.state_dict()[k], v) for k, v in ref.items())
    t = max_over_ranks(float(rank + 1))
    q.put((rank, ok, sent, t, list(scene_shard(7, rank, world))))
    dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_timing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "weights differ after broadcast"
    assert res[0][2] == res[1][2] > 0
    assert res[0][3] == res[1][3] == 2.0                 # max over ranks
    assert res[0][4] + res[1][4] == list(range(7))

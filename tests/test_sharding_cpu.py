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


def _layout_worker(rank, world, port, q):
    """at density < 1 every rank draws its block layouts from its own RNG state; the start-up broadcast must ship the int64
    `master_layout` buffers too (the reference's dist.broadcast(master_layout), sparse_self_attention.py:50-52)."""
    import zlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_view_generation.modules.transformer.mingpt_sparse import GPT, GPTConfig
    from tests.cases import GPT_SMALL
    torch.manual_seed(100 + rank)
    gpt = GPT(GPTConfig(**{**GPT_SMALL, "density": 0.3}))
    lay = lambda: torch.stack([b.attention.sparse_self_attention.master_layout for b in gpt.blocks])
    before = zlib.crc32(lay().numpy().tobytes())
    gpt._engine, gpt._engine_key = "stale", "stale"               # a packed engine built before the broadcast must be dropped
    broadcast_module_weights(gpt, src=0)
    after = zlib.crc32(lay().numpy().tobytes())
    w = zlib.crc32(gpt.head.weight.detach().numpy().tobytes())
    q.put((rank, before, after, w, gpt._engine is None, float(lay().float().mean())))
    dist.destroy_process_group()


def test_two_rank_gloo_broadcast_ships_integer_layout_buffers_and_drops_engines():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_layout_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] != res[1][1], "the two ranks should have drawn different layouts (different RNG state)"
    assert res[0][2] == res[1][2] == res[0][1], "layouts differ after the broadcast"
    assert res[0][3] == res[1][3], "weights differ after the broadcast"
    assert res[0][4] and res[1][4], "pre-packed engines must be invalidated by the broadcast"
    assert 0.2 < res[0][5] < 0.45

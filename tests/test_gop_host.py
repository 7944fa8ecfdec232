"""Host-side logic of the GOP pipeline: the dyadic schedule, padding rule, static GOP sharding and the statistics
gather (world_size 2 on gloo) -- no GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from learned_pmctf_b200 import gop as G
from learned_pmctf_b200 import parallel as par
from oracle import oracle as orc


def reference_schedule(gop_size):
    """Transcription of the loop variables of test_pMCTF_flex.py:91-146 (independent of gop.dyadic_schedule)."""
    stages = 1
    while 2 ** stages < gop_size:
        stages += 1
    assert 2 ** stages == gop_size
    out, num_frames = [], gop_size
    for stage_idx in range(stages):
        num_frames = num_frames // 2
        pairs = []
        for group_idx in range(num_frames):
            group_step = 2 ** stage_idx
            frame_idx_gop = group_idx * 2 * group_step
            pairs.append((frame_idx_gop, frame_idx_gop + group_step))
        out.append(pairs)
    return out


@pytest.mark.parametrize("g", [2, 4, 8, 16, 32])
def test_dyadic_schedule(g):
    s = G.dyadic_schedule(g)
    assert s == reference_schedule(g)
    assert sum(len(p) for p in s) == g - 1          # 8+4+2+1 = 15 pairs for GOP-16
    cur = sorted(c for p in s for _, c in p)
    assert cur == list(range(1, g))                  # every frame but 0 becomes an H frame exactly once
    # the strided-view batching of GopCodec.analysis picks exactly these frames
    level = list(range(g))
    for pairs in s:
        assert list(zip(level[0::2], level[1::2])) == pairs
        level = level[0::2]
    assert level == [0]


def test_bad_gop_size():
    with pytest.raises(ValueError):
        G.dyadic_schedule(12)


def test_padding_rule():
    assert G.get_padding_size(1080, 1920, 128) == (0, 0, 0, 72)   # 1080p -> 1152 x 1920
    assert G.get_padding_size(256, 448, 128) == (0, 64, 0, 0)
    assert G.get_padding_size(128, 128, 128) == (0, 0, 0, 0)


def test_frame_of_plane_order():
    class M:  # GopCodec only stores the model
        pass
    c = G.GopCodec(M(), 8)
    assert c._frame_of_plane() == [1, 3, 5, 7, 2, 6, 4, 0]


def test_sharding_covers_every_item_once():
    items = par.work_items([0, 4, 8, 12, 16, 20], 7, 6)   # C4: 6 q points x 7 sequences x 6 GOPs
    assert len(items) == 252 and items[0] == (0, 0, 0) and items[7] == (0, 1, 1)
    for world in (1, 2, 4, 8):
        shards = [par.shard(items, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == sorted(items)
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
        assert max(map(len, shards)) == par.max_items_per_rank(len(items), world)
    with pytest.raises(ValueError):
        par.shard(items, 2, 2)


def test_oracle_host_pieces():
    g = np.random.default_rng(3)
    u = g.integers(0, 256, (3, 10, 12), dtype=np.uint8)
    p = orc.unpack_u8(u, 16, 16)
    assert p.shape == (3, 16, 16) and np.array_equal(p[:, :10, :12], u) and p[:, 10:].sum() == 0 and p[:, :, 12:].sum() == 0
    rec = p + g.normal(0, 3, p.shape).astype(np.float32)
    sse = orc.frame_sse(rec, u)
    want = [int(((np.rint(np.clip(rec[i, :10, :12], 0, 255)).astype(np.int64) - u[i]) ** 2).sum()) for i in range(3)]
    assert sse.tolist() == want
    assert abs(G.psnr_from_sse(10 * 12 * 4.0, 120) - 10 * np.log10(255 ** 2 / 4.0)) < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = par.init_from_env("gloo")
    items = list(range(n_items))
    mine = par.shard(items, r, w)
    # each item's "statistics" encode its global index, so the gathered order can be checked
    local = torch.stack([torch.full((4, G.N_STATS), float(i), dtype=torch.float64) + torch.arange(G.N_STATS) for i in mine]) \
        if mine else torch.zeros((0, 4, G.N_STATS), dtype=torch.float64)
    out = par.gather_stats(local, n_items, r, w)
    t = par.max_over_ranks(10.0 + r, "cpu")
    q.put((r, out.numpy(), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [6, 7, 1])
def test_gather_stats_world2_gloo(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, out, t in res:
        assert out.shape == (n_items, 4, G.N_STATS)
        for i in range(n_items):
            assert np.array_equal(out[i], np.full((4, G.N_STATS), float(i)) + np.arange(G.N_STATS))
        assert t == 11.0   # max over ranks

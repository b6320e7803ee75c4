"""CPU, world_size 2 over gloo: the scene-sharding host logic used by bench.py / multi-GPU sampling."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffuscene_b200.parallel import gather_scenes, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                off, cnt = shard_range(total, r, world)
                seen += list(range(off, off + cnt))
            assert seen == list(range(total))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 7
    off, cnt = shard_range(total, rank, world)
    local = torch.arange(off, off + cnt, dtype=torch.float32).reshape(cnt, 1, 1).expand(cnt, 2, 3).contiguous()
    full = gather_scenes(local, total)
    ms = max_over_ranks(10.0 + rank, torch.device("cpu"))
    q.put((rank, full[:, 0, 0].tolist(), ms))
    dist.destroy_process_group()


def test_two_rank_gather_and_timing_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29571
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scenes, ms in res:
        assert scenes == [float(i) for i in range(7)]       # global scene order, every scene exactly once
        assert ms == 11.0                                    # max over ranks

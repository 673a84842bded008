"""World-size-2 tests of the scene-sharding host logic on CPU (gloo); the GPU path uses the same helpers
with NCCL (bench.py --gpus N under torchrun)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scenes, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from pixie_b200 import dist_utils as D
    D.init("gloo")
    mine = D.shard_scenes(n_scenes, rank, world)
    # per-rank "timing": rank 1 is slower; the reported time must be the max
    t = D.max_over_ranks(10.0 + 5.0 * rank)
    total = D.sum_over_ranks(float(len(mine)))
    D.barrier()
    recs = D.gather_records([{"scene": i, "rank": rank} for i in mine])
    q.put((rank, mine, t, total, recs))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_scenes", [7, 8, 1])
def test_scene_sharding_matches_distributed_sampler_world2(n_scenes):
    from torch.utils.data import DistributedSampler
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        expect = list(DistributedSampler(range(n_scenes), num_replicas=world, rank=rank, shuffle=False))
        assert res[rank][1] == expect
        assert res[rank][2] == 15.0                       # max over ranks
        assert res[rank][3] == float(sum(len(res[r][1]) for r in range(world)))
    recs = res[0][4]
    assert res[1][4] is None
    assert [r["scene"] for r in recs] == res[0][1] + res[1][1]          # rank order, as gather_object
    covered = {r["scene"] for r in recs}
    assert covered == set(range(n_scenes))                               # every scene processed (some twice when padded)


def test_single_process_helpers():
    sys.path.insert(0, ROOT)
    from pixie_b200 import dist_utils as D
    assert D.shard_scenes(5, 0, 1) == [0, 1, 2, 3, 4]
    assert D.shard_scenes(0, 0, 4) == []
    assert D.shard_scenes(3, 3, 4) == [0]                                # wrap-around padding
    assert D.max_over_ranks(3.5) == 3.5 and D.gather_records([1, 2]) == [1, 2]

"""N > 1 host logic on CPU: env-id sharding and the episode-stat reduction, world_size 2 over gloo, with the CPU oracle
standing in for the per-rank device shard (the reduction code path is the one bench.py uses with nccl)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range(qlb):
    sh = importlib.import_module("q-learning_b200.sharding")
    for world in (1, 2, 3, 8):
        for total in (1, 7, 8, 4096, 65536 * 8 + 5):
            spans = [sh.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(2, 2, 10)


def _worker(rank, world, port, total_envs, steps, seed, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    sh = importlib.import_module("q-learning_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sh.shard_range(rank, world, total_envs)
    env = O.VecEnv(hi - lo, seed=seed, env_id_base=lo)
    acts = O.synthetic_actions(seed, lo, hi - lo, 0, steps)
    for t in range(steps):
        env.step(acts[t])
    s = env.stats()
    vec = torch.tensor([s["sum_return"], s["episodes"], s["steps"],
                        -s["min_return"] if s["episodes"] else -1e300, s["max_return"] if s["episodes"] else -1e300], dtype=torch.float64)
    red = sh.reduce_episode_stats(vec, dist)
    st = env.state()
    out_q.put((rank, lo, hi, red, st["score"].tolist(), st["bricks"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(O):
    import torch.multiprocessing as mp
    total, steps, seed, world = 12, 400, 41, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    # single process over all envs
    env = O.VecEnv(total, seed=seed)
    acts = O.synthetic_actions(seed, 0, total, 0, steps)
    for t in range(steps):
        env.step(acts[t])
    s = env.stats()
    score = np.concatenate([np.array(r[4]) for r in results])
    bricks = np.concatenate([np.array(r[5], dtype=np.uint64) for r in results])
    assert np.array_equal(score, env.state()["score"]) and np.array_equal(bricks, env.state()["bricks"])
    for r in results:
        red = r[3]
        assert red["sum_return"] == s["sum_return"] and red["episodes"] == s["episodes"] and red["steps"] == s["steps"]
        assert red["min_return"] == s["min_return"] and red["max_return"] == s["max_return"]

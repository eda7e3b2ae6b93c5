"""Multi-GPU host logic: envs and replay shard trivially (no data-path collective); the only exchange is the
episode-statistics reduction, off the step path. Works with any torch.distributed backend
(nccl on the GPU box, gloo in the CPU tests)."""


def shard_range(rank, world, total_envs):
    """Global env ids [lo, hi) owned by `rank`: contiguous blocks, sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


STATS_FIELDS = ("sum_return", "episodes", "steps", "neg_min_return", "max_return")


def reduce_episode_stats(vec, dist=None, group=None):
    """vec: 5-element float64 tensor in qlc_stats_export layout {sum_return, episodes, steps, -min_return, max_return}
    (on the backend's device). Sums the first three and max-reduces the last two over all ranks; returns a dict."""
    import torch
    if dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        a = vec[:3].clone()
        b = vec[3:].clone()
        dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
        vec = torch.cat([a, b])
    v = vec.detach().cpu().tolist()
    episodes = v[1]
    return {"sum_return": v[0], "episodes": episodes, "steps": v[2],
            "min_return": (-v[3]) if episodes > 0 else None, "max_return": v[4] if episodes > 0 else None,
            "mean_return": (v[0] / episodes) if episodes > 0 else None}

"""Multi-GPU plumbing: envs are independent, so N GPUs = N shards of the global env-id range with no
collective on the step path.  The only exchange is one end-of-run reduction of episode
statistics (payoff sums per seat, episodes, env-steps)."""
import os

import torch


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def shard_env_id_base(envs_per_rank, rank):
    """Global id of env 0 of this rank's shard: ids [rank*E, (rank+1)*E).  Philox streams are keyed by the
    global id, so results do not depend on how many GPUs the id range is split over."""
    return int(rank) * int(envs_per_rank)


def episode_stats(payoffs, done, env_steps):
    """[P] payoff sums over finished episodes | episodes | env-steps, float64 on the trajectory's device."""
    d = done.bool()
    tot = (payoffs * d.unsqueeze(-1)).sum(tuple(range(payoffs.dim() - 1))).double()
    return torch.cat([tot, d.sum().double().reshape(1),
                      torch.tensor([float(env_steps)], dtype=torch.float64, device=payoffs.device)])


def reduce_stats(stats, group=None):
    """SUM over ranks (NCCL over NVLink on GPUs, gloo in the CPU tests); no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def reduce_max(value, device, group=None):
    """MAX over ranks of a scalar (timings are reported as the slowest rank's)."""
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

"""Multi-GPU plumbing of the LOB step: environments are independent, so the batch is sharded by environment across
ranks (one process per GPU) and the step path has NO collective.  The only exchange is the reduction of episode
statistics once per rollout (the reference's data-parallel trainer does its one all-reduce on gradients,
jaxrl/MARL/ippo_rnn_JAXMARL_pmap.py:566-567; env state is reshaped to (N_DEVICES, NUM_ENVS/N, ...) at :300-334).

Backend: NCCL over NVLink on GPUs, gloo on CPU (tests)."""
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    start: int    # first global environment of this rank
    count: int    # environments owned by this rank


def shard_range(num_envs: int, rank: int, world_size: int) -> Shard:
    """Contiguous blocks, remainder spread over the first ranks (block layout == reshape_pytree_leading_dim when
    num_envs % world_size == 0)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(num_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return Shard(rank, world_size, start, count)


def init_from_env(backend=None):
    """Join the process group torchrun described (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)   # every launch of this rank goes to its own GPU (one process per GPU)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


STAT_KEYS = ("count", "sum", "sumsq", "min", "max")


def local_episode_stats(x: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
    """[5] float64 = count, sum, sum of squares, min, max of ``x`` (optionally where ``mask``) on x's device."""
    x = x.reshape(-1).to(torch.float64)
    if mask is not None:
        x = x[mask.reshape(-1).bool()]
    if x.numel() == 0:
        return torch.tensor([0.0, 0.0, 0.0, float("inf"), float("-inf")], dtype=torch.float64, device=x.device)
    return torch.stack([torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device), x.sum(), (x * x).sum(),
                        x.min(), x.max()])


def reduce_episode_stats(stats: torch.Tensor) -> dict:
    """Combine per-rank [K,5] (or [5]) statistics over the ranks: SUM for count / sum / sumsq, MIN / MAX for the extremes.
    ONE tiny collective (an all-gather of K*5 doubles per rank, combined locally; latency-bound, off the step path).
    Returns mean / std / min / max / count tensors."""
    s = stats.reshape(-1, 5).clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.empty_like(s) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, s.contiguous())
        g = torch.stack(parts)                       # [world, K, 5]
        s = torch.cat([g[:, :, 0:3].sum(dim=0), g[:, :, 3].min(dim=0).values[:, None], g[:, :, 4].max(dim=0).values[:, None]],
                      dim=1)
    n = s[:, 0].clamp(min=1.0)
    mean = s[:, 1] / n
    var = (s[:, 2] / n - mean * mean).clamp(min=0.0)
    return {"count": s[:, 0], "mean": mean, "std": var.sqrt(), "min": s[:, 3], "max": s[:, 4]}

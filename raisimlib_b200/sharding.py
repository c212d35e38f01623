"""Environment sharding across GPUs: contiguous blocks, one process per GPU, and the
single collective of the path -- the all-gather of the observation rows.  Backend-agnostic
(nccl on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(num_envs_total, world_size, rank):
    """rank r owns [r*N/G, (r+1)*N/G); N must divide evenly (weak scaling keeps N/G fixed)."""
    if num_envs_total % world_size != 0:
        raise ValueError(f"{num_envs_total} environments do not split evenly over {world_size} ranks")
    per = num_envs_total // world_size
    return rank * per, (rank + 1) * per


def shard_seed(config_index, rank):
    """seed = config index x 1000 + rank"""
    return config_index * 1000 + rank


def allgather_observations(obs_local, out=None):
    """obs_local [N/G, obDim] -> [N, obDim] in rank order on every rank (no-op without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if out is not None:
            out.copy_(obs_local)
            return out
        return obs_local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * obs_local.shape[0], obs_local.shape[1]), dtype=obs_local.dtype, device=obs_local.device)
    dist.all_gather_into_tensor(out, obs_local.contiguous())
    return out

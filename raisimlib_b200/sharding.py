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


class SharedHostRows:
    """A [world * rows_per_rank, width] float32 array in POSIX shared memory that every rank maps and page-locks.

    Host-side consumers (a trainer process on rank 0 reading observation rows) do not need the NVLink all-gather:
    each rank's step kernel writes its own block of rows IN PLACE over its own PCIe link (the C-ABI recognises the
    page-locked buffer and binds it zero-copy), all links in parallel; one barrier later every row is visible to
    every process (`publish_and_wait`, shared-memory flags).  `local` is this rank's block (a numpy view to hand to
    Batch.control_step), `all` the whole array.
    """

    def __init__(self, tag, world, rank, rows_per_rank, width, register=True, barrier=None):
        import os
        import numpy as np
        self.path = f"/dev/shm/rsb_{tag}"
        self.rank, self.registered = rank, False
        shape = (world * rows_per_rank, width)
        row_bytes = shape[0] * shape[1] * 4
        flag_off = (row_bytes + 4095) // 4096 * 4096          # one cache line of flags per rank, behind the rows
        nbytes = flag_off + world * 64
        self.world = world
        barrier = barrier or (dist.barrier if (dist.is_available() and dist.is_initialized()) else (lambda: None))
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        barrier()
        self.all = np.memmap(self.path, dtype=np.float32, mode="r+", shape=shape)
        self._flags = np.memmap(self.path, dtype=np.int64, mode="r+", offset=flag_off, shape=(world, 8))
        self.local = self.all[rank * rows_per_rank:(rank + 1) * rows_per_rank]
        self._ptr, self._nbytes = self.all.ctypes.data, row_bytes
        if register:
            rc = torch.cuda.cudart().cudaHostRegister(self._ptr, row_bytes, 1 | 2)   # portable | mapped
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed with {rc}")
            self.registered = True
        barrier()

    def publish_and_wait(self, step_id, timeout_s=60.0):
        """Host-side barrier for the rows of control step `step_id` (> 0, increasing): call after this rank's step call
        returned (its rows are then in host memory), returns when every rank has published the same step.  Plain
        shared-memory flags (x86 store order: rows before flag) -- no GPU work, no NCCL on the host-bound path."""
        import time
        self._flags[self.rank, 0] = step_id
        f = self._flags[:, 0]
        spins, t0 = 0, None
        while int(f.min()) < step_id:
            spins += 1
            if spins % 4096 == 0:                      # a dead peer must not hang the job
                t0 = t0 or time.monotonic()
                if time.monotonic() - t0 > timeout_s:
                    raise RuntimeError(f"SharedHostRows: rank(s) {[int(r) for r in (f < step_id).nonzero()[0]]} did not publish step {step_id} within {timeout_s} s")

    def close(self, barrier=None):
        import os
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self._ptr)
            self.registered = False
        barrier = barrier or (dist.barrier if (dist.is_available() and dist.is_initialized()) else (lambda: None))
        barrier()
        self.local = None
        self.all = None
        self._flags = None
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass

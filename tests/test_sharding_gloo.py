"""N > 1 host logic on CPU: world_size-2 gloo process group, contiguous env shards, observation
all-gather in rank order (the path's only collective)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _obs_rows(gc, gv):
    """numpy restatement of the observation row (same layout as rsb_observe_kernel)."""
    from helpers import quat_to_rot
    out = np.zeros((gc.shape[0], 34), np.float32)
    for e in range(gc.shape[0]):
        R = quat_to_rot(gc[e, 3:7])
        out[e] = np.r_[gc[e, 2], R[2], gc[e, 7:], R.T @ gv[e, 0:3], R.T @ gv[e, 3:6], gv[e, 6:]]
    return out


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from raisimlib_b200.sharding import shard_range, allgather_observations
    from helpers import random_state
    from oracle.urdf_tables import load_tables
    t = load_tables(os.path.join(ROOT, "raisimlib_b200", "rsc", "anymal_c_like.urdf"))
    gc, gv = random_state(t, np.random.default_rng(5), total)          # the same global state on every rank
    lo, hi = shard_range(total, world, rank)
    local = torch.from_numpy(_obs_rows(gc[lo:hi], gv[lo:hi]))
    full = allgather_observations(local)
    # the gather object bench.py drives: "nccl" mode = the collective after the step (gloo here), two alternating row buffers
    from raisimlib_b200.sharding import ObservationGather
    og = ObservationGather(None, world, rank, hi - lo, 34, mode="nccl")
    r1 = og.gather(local).clone(); r2 = og.gather(2 * local)
    assert torch.equal(r1, full) and torch.equal(r2, 2 * full) and r1.data_ptr() != r2.data_ptr()
    assert og.report()["mode"] == "nccl"
    # host-side consumers: every rank writes its own block of a shared, mapped array; one barrier makes all rows visible
    from raisimlib_b200.sharding import SharedHostRows
    sh = SharedHostRows(f"test_{port}", world, rank, hi - lo, 34, register=False)
    sh.local[:] = local.numpy()
    sh.publish_and_wait(1)                       # shared-memory flag barrier (no collective)
    shared_copy = np.array(sh.all)
    sh.close()
    q.put((rank, lo, hi, full.numpy().copy(), shared_copy))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    from raisimlib_b200.sharding import shard_range, shard_seed
    assert [shard_range(32768, 8, r) for r in (0, 7)] == [(0, 4096), (28672, 32768)]
    assert shard_range(4096, 1, 0) == (0, 4096)
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
    assert shard_seed(3, 5) == 3005


def test_allgather_two_ranks_gloo():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import random_state
    from oracle.urdf_tables import load_tables
    total, world, port = 16, 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = load_tables(os.path.join(ROOT, "raisimlib_b200", "rsc", "anymal_c_like.urdf"))
    gc, gv = random_state(t, np.random.default_rng(5), total)
    ref = _obs_rows(gc, gv)
    for rank, lo, hi, full, shared_copy in got:
        assert (lo, hi) == (rank * 8, rank * 8 + 8)
        assert full.shape == (total, 34)
        assert np.array_equal(full, ref)          # every rank holds all rows, in rank order
        assert np.array_equal(shared_copy, ref)   # the shared host array holds the same rows on every rank


def test_allgather_single_process_is_identity():
    from raisimlib_b200.sharding import allgather_observations
    x = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    assert allgather_observations(x) is x

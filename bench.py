#!/usr/bin/env python3
"""bench.py -- env-steps/s of the batched World::integrate() hot path (BASELINE.json metric).

Workload (BASELINE.json configs[2], weak-scaled per GPU -> configs[4] at 8 GPUs):
  4096 ANYmal-C-like environments per GPU on a 513x513 random rough height field (51.2 m square,
  3-octave value noise, +-0.10 m), PD stance control with per-control-step target jitter.
A bench "step" is one RaisimGym control step: ONE fused launch of 4 sub-steps of World::integrate()
for every environment (= 4 x 4096 env-steps per GPU), followed by the observation kernel and, at
N > 1, the NCCL all-gather of the observation rows.

  value   env-steps/s with state, targets and terrain resident in HBM (targets read in place)
  e2e     same metric through the C-ABI with HOST buffers: pinned H2D of the PD targets and D2H of
          the (all-gathered) observation rows inside the timed region
  --impl reference   the CPU path (oracle port of World::integrate(), OpenMP over envs, all host
          cores) on the same workload -- /root/reference holds no buildable source
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
SUBSTEPS = 4
GC0 = np.array([0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8], dtype=np.float64)
KP, KD = 300.0, 8.0
HM = dict(xs=513, ys=513, size=51.2, amp=0.10)
RING = 8            # distinct PD-target sets cycled through (synthetic "policy output")
SETTLE = 40         # untimed control steps run while BUILDING the workload: robots are dropped from 0.75 m and must stand
                    # on the terrain before warm-up starts, whatever --warmup the caller passes
L2_FLUSH_BYTES = 256 << 20


def value_noise(rng, n, size_cells):
    """bilinear value noise on an n x n grid with lattice spacing size_cells"""
    m = (n - 1) // size_cells + 2
    lat = rng.uniform(-1, 1, (m, m))
    x = np.arange(n) / size_cells
    i = x.astype(int); f = x - i
    a = lat[np.ix_(i, i)]; b = lat[np.ix_(i, i + 1)]; c = lat[np.ix_(i + 1, i)]; d = lat[np.ix_(i + 1, i + 1)]
    fy, fx = f[:, None], f[None, :]
    return a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx


def make_workload(rank, n_envs):
    rng = np.random.default_rng(3000 + rank)
    n = HM["xs"]
    H = value_noise(rng, n, 32) + 0.5 * value_noise(rng, n, 8) + 0.25 * value_noise(rng, n, 2)
    H = (HM["amp"] * H / np.abs(H).max()).astype(np.float32)
    gc = np.tile(GC0, (n_envs, 1))
    gc[:, 0:2] = rng.uniform(-0.4 * HM["size"], 0.4 * HM["size"], (n_envs, 2))      # inner 80 % of the map
    gc[:, 2] = 0.75
    gc[:, 7:] += rng.uniform(-0.2, 0.2, (n_envs, 12))
    yaw = rng.uniform(-np.pi, np.pi, n_envs)
    gc[:, 3] = np.cos(yaw / 2); gc[:, 6] = np.sin(yaw / 2)
    gv = np.zeros((n_envs, 18))
    targets = np.tile(GC0, (RING, n_envs, 1))
    targets[:, :, 7:] += rng.uniform(-0.15, 0.15, (RING, n_envs, 12))
    kp = np.r_[np.zeros(6), KP * np.ones(12)]; kd = np.r_[np.zeros(6), KD * np.ones(12)]
    return H, gc, gv, targets, kp, kd


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in a
    background thread every 5 ms (nvidia-smi -lms needs > 100 ms to produce its first row, longer
    than a short timed region)."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self.stop_flag, self.t = index, [], set(), None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid_order = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(uuid_order.split(",")[index]) if uuid_order and uuid_order.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self.stop_flag:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.BAD.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def usable_cores():
    """host threads this process may really use: CPU affinity, capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); p_ = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = max(1, min(n, q // p_))
        except Exception:
            pass
    return n


def algorithmic_bytes_per_env_step(nq, nv, nj, kbar):
    # reads gc, gv, tau_ff, pTarget, vTarget; writes gc+, gv+, tau_applied, count, 12 words per contact
    return 4.0 * (2 * nq + 3 * nv + 2 * nj + 1 + 12.0 * kbar)


def cpu_baseline(sample_envs, max_steps, budget_s=12.0):
    """the oracle (a port: no reference binary exists) on all host cores, bounded sample of the workload"""
    from oracle.oracle import Oracle
    from oracle.urdf_tables import load_tables
    from raisimlib_b200 import RSC_DIR
    H, gc, gv, targets, kp, kd = make_workload(0, sample_envs)
    o = Oracle(load_tables(os.path.join(RSC_DIR, "anymal_c_like.urdf")), params=dict(threshold=1e-6, slip_bisect=1))   # CPU-tuned slip search
    o.set_heightmap(HM["xs"], HM["ys"], HM["size"], HM["size"], 0.0, 0.0, H.astype(np.float64))
    cores = usable_cores()
    vt = np.zeros((sample_envs, 18))
    for k in range(SETTLE + 3):     # workload construction + warm-up: robots land on the terrain and settle, like the GPU arm
        o.step(gc, gv, n_steps=SUBSTEPS, ptarget=targets[k % RING], vtarget=vt, kp=kp, kd=kd, nthreads=cores)
    t0 = time.perf_counter(); done = 0; k = 0
    while done < max_steps and time.perf_counter() - t0 < budget_s:
        o.step(gc, gv, n_steps=SUBSTEPS, ptarget=targets[k % RING], vtarget=vt, kp=kp, kd=kd, nthreads=cores)
        done += SUBSTEPS; k += 1
    dt = time.perf_counter() - t0
    return {"value": sample_envs * done / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{sample_envs} envs x {done} sub-steps of the same workload (float64 oracle, bisection slip search, OpenMP over envs, after the same settling phase as the GPU arm)"}


def workload_string(n):
    return (f"{n} ANYmal-C-like envs per GPU on 513x513 rough height field (+-0.10 m), PD stance kp={KP} kd={KD}, dt=0.0025, "
            f"{SUBSTEPS} sub-steps fused per step")


def run_reference(args):
    """--impl reference: the CPU path on this box's host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import Oracle
    from oracle.urdf_tables import load_tables
    from raisimlib_b200 import RSC_DIR
    n = ENVS_PER_GPU
    H, gc, gv, targets, kp, kd = make_workload(0, n)
    o = Oracle(load_tables(os.path.join(RSC_DIR, "anymal_c_like.urdf")), params=dict(threshold=1e-6, slip_bisect=1))   # CPU-tuned slip search
    o.set_heightmap(HM["xs"], HM["ys"], HM["size"], HM["size"], 0.0, 0.0, H.astype(np.float64))
    vt = np.zeros((n, 18))
    cores = usable_cores()          # explicit: torchrun exports OMP_NUM_THREADS=1
    for k in range(SETTLE + max(args.warmup, 3)):
        o.step(gc, gv, n_steps=SUBSTEPS, ptarget=targets[k % RING], vtarget=vt, kp=kp, kd=kd, nthreads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        o.step(gc, gv, n_steps=SUBSTEPS, ptarget=targets[k % RING], vtarget=vt, kp=kp, kd=kd, nthreads=cores)
    dt = time.perf_counter() - t0
    val = n * SUBSTEPS * args.steps / dt
    line = {"impl": "reference", "metric": "env-steps/s", "value": val, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(n), "envs_per_gpu": n, "substeps_per_step": SUBSTEPS,
                       "arm": f"CPU oracle port (a restatement, not RaiSim's binary), float64, OpenMP over envs, {cores} threads; each step = the full per-GPU workload"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n} envs x {SUBSTEPS * args.steps} sub-steps (the full per-GPU workload), float64, OpenMP over envs"},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from raisimlib_b200 import capi, RSC_DIR
    from raisimlib_b200.sharding import allgather_observations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    n = args.envs
    H, gc, gv, targets, kp, kd = make_workload(rank, n)
    model = capi.Model(os.path.join(RSC_DIR, "anymal_c_like.urdf"))
    bt = capi.Batch(model, n, device=local)
    bt.set_params(threshold=1e-6)
    bt.set_heightmap(HM["xs"], HM["ys"], HM["size"], HM["size"], 0.0, 0.0, H)
    bt.set_pd_gains(kp, kd)
    stream = torch.cuda.current_stream()
    bt.set_stream(stream.cuda_stream)
    bt.set_state(gc.astype(np.float32), gv.astype(np.float32))
    od = bt.ob_dim()
    tg_dev = torch.tensor(targets, dtype=torch.float32, device="cuda")                  # [RING, n, nq] resident in HBM
    tg_pin = torch.tensor(targets, dtype=torch.float32).pin_memory()                    # host copies for the e2e arm
    vt_dev = torch.zeros((n, 18), dtype=torch.float32, device="cuda")
    bt.set_pd_target(tg_dev[0], vt_dev)
    obs = torch.empty((n, od), dtype=torch.float32, device="cuda")
    obs_all = torch.empty((world * n, od), dtype=torch.float32, device="cuda") if world > 1 else obs
    obs_host = torch.empty((n, od), dtype=torch.float32).pin_memory() if world == 1 else None
    shared = None
    if world > 1:
        # e2e arm at N > 1: the host-side consumer reads one [world*n, obDim] array in shared, page-locked memory; every
        # rank's step kernel writes its own rows in place over its own PCIe link (no NVLink hop, no D2H funnel on rank 0)
        from raisimlib_b200.sharding import SharedHostRows
        shared = SharedHostRows(f"bench_{os.environ.get('MASTER_PORT', '0')}", world, rank, n, od)
        e2e_step = [0]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")

    def control_step(k, host_io):
        assert not host_io
        bt.bind_pd_target(tg_dev[k % RING])                  # resident targets read in place (zero-copy)
        bt.control_step(None, SUBSTEPS, obs)                 # ONE fused launch: 4 x World::integrate() + observation rows
        if world > 1:
            allgather_observations(obs, obs_all)             # the only collective of the path

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host_io, steps, step0):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for k in range(steps):
            flush.fill_(float(k))                            # evict L2 between timed iterations (256 MiB > 126 MB L2)
            ev[k][0].record(stream)
            if host_io and world == 1:
                # the call a user makes: targets (pinned host) in, 4 fused sub-steps, observation rows (pinned host) out
                kev[k][0].record(stream)
                bt.control_step(tg_pin[(step0 + k) % RING], SUBSTEPS, obs_host)
                kev[k][1].record(stream)
                ev[k][1].record(stream)
                continue
            if host_io:
                # N > 1: same call per rank (it returns when this rank's rows are in host memory); shared-memory flags are
                # the barrier after which the trainer's rank may read every row
                kev[k][0].record(stream)
                bt.control_step(tg_pin[(step0 + k) % RING], SUBSTEPS, shared.local)
                kev[k][1].record(stream)
                e2e_step[0] += 1
                shared.publish_and_wait(e2e_step[0])         # host-side barrier: every rank's rows are in the shared array
                ev[k][1].record(stream)
                continue
            bt.bind_pd_target(tg_dev[(step0 + k) % RING])    # resident targets read in place (zero-copy)
            kev[k][0].record(stream)
            bt.control_step(None, SUBSTEPS, obs)             # ONE launch: 4 fused sub-steps + observation rows
            kev[k][1].record(stream)
            if world > 1:
                allgather_observations(obs, obs_all)
            ev[k][1].record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        kms = sum(a.elapsed_time(b) for a, b in kev)
        t = torch.tensor([ms, kms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)         # max over ranks
        return float(t[0]), float(t[1])

    for k in range(SETTLE):          # workload construction (untimed): land and settle
        control_step(k, False)
    for k in range(W):
        control_step(k, False)
    barrier()
    l0 = bt.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, ms_kernel = timed(False, args.steps, W)
    launches = bt.launch_count() - l0
    ms_e2e, _ = timed(True, args.steps, W + args.steps)
    clocks = sampler.stop() if rank == 0 else None
    _, cnt = bt.contacts()
    it = bt.solver_iterations()
    g, _v = bt.get_state()
    stats = torch.tensor([float(cnt.mean()), float(it.mean()), float(it.max()), float((g[:, 2] > 0.25).mean())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        stats /= world
    if rank == 0:
        total_env_steps = world * n * SUBSTEPS * args.steps
        value = total_env_steps / (ms_dev * 1e-3)
        e2e = total_env_steps / (ms_e2e * 1e-3)
        kbar = float(stats[0])
        B = algorithmic_bytes_per_env_step(19, 18, 12, kbar)
        launch_ms = ms_kernel / args.steps
        achieved = B * n * SUBSTEPS / (launch_ms * 1e-3) / 1e9
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(n),
                       "envs_per_gpu": n, "substeps_per_step": SUBSTEPS, "l2": "flushed between timed iterations (256 MiB write)",
                       "mean_contacts_per_env": kbar, "mean_solver_iters": float(stats[1]), "max_solver_iters": float(stats[2]),
                       "standing_fraction": float(stats[3]), "parallelism": f"env-shard x{world}" + (", NCCL obs all-gather" if world > 1 else "")},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": int(world * n * 19 * 4), "d2h_bytes_per_step": int(world * n * od * 4),
                    "path": "pinned host targets and observation rows read / written in place by the step kernel (zero-copy over PCIe)"
                            + ("; rows of all ranks land in one shared page-locked array, shared-memory flags as barrier" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "rsb_step_kernel (fused FK+CRBA+RNEA+narrow-phase+contact solver+integrate)",
                         "algorithmic_bytes_per_env_step": B, "launch_ms": launch_ms, "peak_source": peak_src,
                         "note": "path is FP32-latency/issue bound, not HBM bound (SURVEY.md 7 hard part 2); see DESIGN.md roofline"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(n, 400)
        print(json.dumps(line), flush=True)
    if shared is not None:
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Timing of VectorizedEnvironment::step() on the device (rsb_batch_gym_step), device-resident buffers, CUDA events (exploratory,
run under gpurun).  usage: gym_probe.py [path/to/librsb.so] -- an explicit library path lets two builds be compared in one call."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raisimlib_b200 import capi, RSC_DIR
if len(sys.argv) > 1:
    capi.LIB_PATH = os.path.abspath(sys.argv[1])

n, substeps = 4096, 4
gc0 = np.array([0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8])
m = capi.Model(os.path.join(RSC_DIR, "anymal_c_like.urdf"))
bt = capi.Batch(m, n)
stream = torch.cuda.current_stream()
bt.set_stream(stream.cuda_stream)
bt.set_ground(0.0)
bt.set_params(threshold=1e-6)
bt.set_pd_gains(np.r_[np.zeros(6), 100.0 * np.ones(12)], np.r_[np.zeros(6), 2.0 * np.ones(12)])
feet = [m.body_index(f"{leg}_FOOT") for leg in ("LF", "RF", "LH", "RH")]
bt.gym_configure(gc0, np.zeros(18), gc0[7:], 0.6 * np.ones(12), feet)
bt.gym_reset()
g = torch.Generator(device="cuda"); g.manual_seed(7)
acts = [0.15 * torch.randn((n, 12), device="cuda", generator=g) for _ in range(8)]
obs = torch.empty((n, 34), device="cuda"); rew = torch.empty(n, device="cuda"); done = torch.empty(n, dtype=torch.uint8, device="cuda")
for k in range(30):
    bt.gym_step(acts[k % 8], substeps, obs, rew, done)
torch.cuda.synchronize()
l0 = bt.launch_count()
steps = 60
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
ndone = 0
for k in range(steps):
    ev[k][0].record(stream); bt.gym_step(acts[k % 8], substeps, obs, rew, done); ev[k][1].record(stream)
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in ev])
print(f"{capi.LIB_PATH}: gym step median {np.median(ms):.4f} ms (min {ms.min():.4f}) = {n * substeps / np.median(ms) * 1e3:.3e} env-steps/s; "
      f"launches per step {(bt.launch_count() - l0) / steps:.2f}; reward mean {float(rew.mean()):.4f}, done {int(done.sum())}, obs checksum {float(obs.double().sum()):.6f}")

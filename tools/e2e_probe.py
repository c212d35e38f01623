#!/usr/bin/env python3
"""Where does the e2e arm (pinned host buffers read / written in place by the step kernel) lose time against the device-resident
arm?  Times one control step with CUDA events for the four combinations of {targets, observation rows} x {HBM, pinned host}."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    wl = bench.Workload("c3", 0, bench.ENVS_PER_GPU)
    with torch.cuda.stream(stream):
        sim = bench.GpuSim(wl, 0, stream)
        for k in range(wl.settle + 20):
            sim.step_resident(k)
        torch.cuda.synchronize()
        flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
        k0 = wl.settle + 20
        g, v = sim.bt.get_state()
        for tin in ("hbm", "host"):
            for tout in ("hbm", "host"):
                sim.bt.set_state(g, v)
                ts = []
                for k in range(60):
                    flush.fill_(float(k))
                    tgt = (sim.ring_dev if tin == "hbm" else sim.ring_pin)[(k0 + k) % bench.RING]
                    out = sim.obs[k & 1] if tout == "hbm" else sim.obs_host
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    sim.bt.control_step(tgt, bench.SUBSTEPS, out)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                print("targets %-4s  observations %-4s : median %.4f ms  (min %.4f)" % (tin, tout, np.median(ts[10:]), min(ts[10:])))
        # the same host buffers through the copy engines instead of in-place PCIe access by the kernel: H2D of the targets before the
        # launch and / or D2H of the observation rows after it, on the launching stream, inside the event window
        tgt_dev = torch.empty_like(sim.ring_dev[0])
        for tin in ("zero-copy", "dma"):
            for tout in ("zero-copy", "dma"):
                sim.bt.set_state(g, v)
                ts = []
                for k in range(60):
                    flush.fill_(float(k))
                    src = sim.ring_pin[(k0 + k) % bench.RING]
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    if tin == "dma":
                        tgt_dev.copy_(src, non_blocking=True)
                    sim.bt.control_step(tgt_dev if tin == "dma" else src, bench.SUBSTEPS, sim.obs[k & 1] if tout == "dma" else sim.obs_host)
                    if tout == "dma":
                        sim.obs_host.copy_(sim.obs[k & 1], non_blocking=True)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                print("host targets by %-9s  host observations by %-9s : median %.4f ms  (min %.4f)" % (tin, tout, np.median(ts[10:]), min(ts[10:])))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Exploratory GPU-vs-oracle error report (run under gpurun); informs the tolerances in tests/."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from raisimlib_b200 import capi, RSC_DIR
from oracle.oracle import Oracle
from oracle.urdf_tables import load_tables
from helpers import ANYMAL_GC0, random_state


def report(name, a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    print(f"  {name:12s} max_abs={err.max():.3e} mean_abs={err.mean():.3e} ref_scale={np.abs(b).max():.3e}", flush=True)


def stage_errors(urdf, n=256, seed=0, base_z=0.45, terrain="ground"):
    print(f"== {urdf} n={n} terrain={terrain}", flush=True)
    path = os.path.join(RSC_DIR, urdf)
    t = load_tables(path)
    m = capi.Model(path)
    bt = capi.Batch(m, n)
    rng = np.random.default_rng(seed)
    gc, gv = random_state(t, rng, n, vel_scale=0.5, base_z=base_z, pos_scale=2.0)
    tau = rng.uniform(-20, 20, (n, t["nv"])); tau[:, :6] = 0
    prm = dict(threshold=1e-6)
    o64, o32 = Oracle(t, params=prm), Oracle(t, precision="f32", params=prm)
    if terrain == "ground":
        for o in (o64, o32): o.set_ground(0.0)
        bt.set_ground(0.0)
    else:
        xs = ys = 65
        H = 0.1 * rng.uniform(-1, 1, (ys, xs))
        for o in (o64, o32): o.set_heightmap(xs, ys, 12.8, 12.8, 0.0, 0.0, H)
        bt.set_heightmap(xs, ys, 12.8, 12.8, 0.0, 0.0, H)
    gc32, gv32 = gc.astype(np.float32), gv.astype(np.float32)
    gc, gv = gc32.astype(np.float64), gv32.astype(np.float64)
    bt.set_control_mode(capi.FORCE_AND_TORQUE)
    bt.set_state(gc32, gv32)
    bt.set_generalized_force(tau.astype(np.float32))
    tau = tau.astype(np.float32).astype(np.float64)
    bt.integrate1()
    M, h = bt.mass_matrix(), bt.nonlinearities()
    R, p = bt.body_poses()
    ct, cnt = bt.contacts(); pts = bt.contact_points()
    a64, v64 = gc.copy(), gv.copy()
    d64 = o64.step(a64, v64, tau_ff=tau, debug=True)
    a32, v32 = gc.copy(), gv.copy()
    d32 = o32.step(a32, v32, tau_ff=tau, debug=True)
    report("M", M, d64["M"]); report("h", h, d64["h"]); report("R", R, d64["R"]); report("p", p, d64["p"])
    print("  oracle f32 vs f64:"); report("M32", d32["M"], d64["M"]); report("h32", d32["h"], d64["h"])
    print("  ncontacts gpu/f32/f64:", cnt.sum(), d32["ncontacts"].sum(), d64["ncontacts"].sum(), " mismatching envs vs f32:",
          int((pts != d32["c_pt"]).any(1).sum()), "vs f64:", int((pts != d64["c_pt"]).any(1).sum()), flush=True)
    bt.integrate2()
    g1, v1 = bt.get_state()
    it = bt.solver_iterations()
    print("  iters gpu mean/max", it.mean(), it.max(), " oracle64", d64["iters"].mean(), d64["iters"].max(), " oracle32", d32["iters"].mean(), d32["iters"].max())
    same = (pts == d64["c_pt"]).all(1)
    report("gc+ (all)", g1, a64); report("gv+ (all)", v1, v64)
    report("gc+ (same ct)", g1[same], a64[same]); report("gv+ (same ct)", v1[same], v64[same])
    print("  oracle f32 vs f64 one step:"); report("gv32", v32[same], v64[same])
    worst = np.abs(v1 - v64).max(1)
    wi = int(np.argmax(np.where(same, worst, 0)))
    print("  worst env", wi, "K", cnt[wi], "iters", it[wi], d64["iters"][wi], "err", worst[wi])
    # trajectory
    for nsteps in (10, 50):
        bt.set_state(gc32, gv32)
        bt.integrate(nsteps)
        gN, vN = bt.get_state()
        aN, bN = gc.copy(), gv.copy()
        o64.step(aN, bN, n_steps=nsteps, tau_ff=tau)
        cN, dN = gc.copy(), gv.copy()
        o32.step(cN, dN, n_steps=nsteps, tau_ff=tau)
        e = np.abs(gN - aN).max(1); e32 = np.abs(cN - aN).max(1)
        print(f"  {nsteps} steps: gc err gpu-vs-f64 median={np.median(e):.2e} p90={np.quantile(e, 0.9):.2e} max={e.max():.2e} | f32-vs-f64 median={np.median(e32):.2e} p90={np.quantile(e32, 0.9):.2e} max={e32.max():.2e}", flush=True)


def timing(urdf, n=4096, terrain="ground", substeps=4, reps=50, z0=None, kp_val=300.0, kd_val=8.0):
    import torch
    path = os.path.join(RSC_DIR, urdf)
    t = load_tables(path)
    m = capi.Model(path)
    bt = capi.Batch(m, n)
    rng = np.random.default_rng(1)
    if t["nq"] == 19:
        gc = np.tile(ANYMAL_GC0, (n, 1)); gc[:, 7:] += rng.uniform(-0.2, 0.2, (n, 12)); gc[:, 2] = rng.uniform(0.55, 0.65, n)
        gc[:, :2] = rng.uniform(-5, 5, (n, 2))
        if z0 is not None: gc[:, 2] = z0
    else:
        gc = np.zeros((n, t["nq"])); gc[:, 2] = 0.95; gc[:, 3] = 1.0
    gv = 0.1 * rng.standard_normal((n, t["nv"]))
    if terrain == "ground":
        bt.set_ground(0.0)
    elif terrain == "hm":
        xs = ys = 129
        bt.set_heightmap(xs, ys, 12.8, 12.8, 0.0, 0.0, 0.05 * rng.uniform(-1, 1, (ys, xs)))
    kp = np.r_[np.zeros(6), kp_val * np.ones(t["nv"] - 6)]; kd = np.r_[np.zeros(6), kd_val * np.ones(t["nv"] - 6)]
    bt.set_pd_gains(kp, kd)
    bt.set_state(gc.astype(np.float32), gv.astype(np.float32))
    bt.set_pd_target(gc.astype(np.float32), np.zeros((n, t["nv"]), np.float32))
    s = torch.cuda.Stream()
    bt.set_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        for _ in range(40): bt.integrate(substeps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(reps): bt.integrate(substeps)
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _, cnt = bt.contacts(); it = bt.solver_iterations()
    g, v = bt.get_state()
    print(f"TIMING {urdf} n={n} {terrain} substeps={substeps}: {ms:.3f} ms/launch -> {n * substeps / ms * 1e3:.3e} env-steps/s ; meanK={cnt.mean():.2f} iters mean={it.mean():.1f} max={it.max()} finite={np.isfinite(g).all()}", flush=True)
    bt.set_stream(0)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "err"):
        stage_errors("anymal_c_like.urdf")
        stage_errors("anymal_c_like.urdf", terrain="hm", seed=3)
        stage_errors("atlas_like.urdf", base_z=0.9)
    if which == "prof":     # one scenario for ncu: -k regex:rsb_step -s 42 -c 1
        timing("anymal_c_like.urdf", reps=3)
    if which in ("all", "time"):
        timing("anymal_c_like.urdf", terrain="none", z0=1000.0, reps=20)      # airborne: no contacts at all
        timing("anymal_c_like.urdf", terrain="none", z0=1000.0, substeps=1, reps=20)
        timing("anymal_c_like.urdf")                                            # standing on flat ground
        timing("anymal_c_like.urdf", substeps=1)
        timing("anymal_c_like.urdf", terrain="hm")
        timing("anymal_c_like.urdf", kp_val=50.0, kd_val=0.5)                   # weak PD: robots kneel (knee + foot contacts)
        timing("atlas_like.urdf", terrain="none", z0=1000.0, reps=10)
        timing("atlas_like.urdf")

#!/usr/bin/env python3
"""small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel path once"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from raisimlib_b200 import capi, RSC_DIR
from helpers import ANYMAL_GC0, PENDULUM_URDF, random_state

for urdf, n, z in (("anymal_c_like.urdf", 40, 0.4), ("atlas_like.urdf", 20, 0.6), (PENDULUM_URDF, 10, 0.0)):
    path = os.path.join(RSC_DIR, urdf) if urdf.endswith(".urdf") else urdf
    m = capi.Model(path)
    t = m.tables()          # same keys as the Python restatement's tables; no oracle import outside tests/
    b = capi.Batch(m, n)
    rng = np.random.default_rng(0)
    gc, gv = random_state(t, rng, n, vel_scale=0.5, base_z=z)
    H = (0.05 * rng.uniform(-1, 1, (33, 33))).astype(np.float32)
    b.set_heightmap(33, 33, 8.0, 8.0, 0.0, 0.0, H) if t["floating"] else b.set_ground(-0.55)
    b.set_state(gc.astype(np.float32), gv.astype(np.float32))
    kp = np.r_[np.zeros(6 if t["floating"] else 0), 50.0 * np.ones(t["nv"] - (6 if t["floating"] else 0))]
    b.set_pd_gains(kp, 0.1 * kp)
    b.set_pd_target(gc.astype(np.float32), np.zeros((n, t["nv"]), np.float32))
    b.integrate1(); b.integrate2(); b.integrate(3)
    b.mass_matrix(); b.contacts(); b.observe()
    g, v = b.get_state()
    print(urdf[:20], "finite", np.isfinite(g).all())
print("SANITIZE_PROBE_DONE")

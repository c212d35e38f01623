#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the float64 CPU oracle (the reference itself cannot run: its
snapshot holds no source -- SURVEY.md 8c).  The fixtures pin the ORACLE across refactors (CPU test)
and give the GPU tests a reference that does not need the oracle at run time.
Re-run only when a convention in DESIGN.md section 2 changes on purpose."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle
from oracle.urdf_tables import load_tables
from helpers import random_state

OUT = os.path.join(ROOT, "tests", "golden")
PRM = dict(threshold=1e-6, stall_window=0, accel_m=0)     # the published method as is: plain sweeps, maxIter semantics


def make(name, urdf, n, seed, terrain, base_z, tau_scale, steps):
    t = load_tables(os.path.join(ROOT, "raisimlib_b200", "rsc", urdf))
    rng = np.random.default_rng(seed)
    gc, gv = random_state(t, rng, n, vel_scale=0.5, base_z=base_z, pos_scale=2.0, joint_scale=0.4)
    tau = rng.uniform(-tau_scale, tau_scale, (n, t["nv"])); tau[:, :6] = 0
    gc, gv, tau = (x.astype(np.float32).astype(np.float64) for x in (gc, gv, tau))
    o = Oracle(t, params=PRM)
    H = None
    if terrain == "hm":
        H = (0.1 * rng.uniform(-1, 1, (33, 41))).astype(np.float32)
        o.set_heightmap(41, 33, 8.0, 6.4, 0.2, -0.1, H.astype(np.float64))
    else:
        o.set_ground(0.0)
    a, b = gc.copy(), gv.copy()
    d = o.step(a, b, tau_ff=tau, debug=True)
    out = dict(gc0=gc, gv0=gv, tau=tau, gc1=a.copy(), gv1=b.copy(), M=d["M"], h=d["h"], ncontacts=d["ncontacts"], c_pt=d["c_pt"],
               c_body=d["c_body"], c_pair=d["c_pair"], c_depth=d["c_depth"], c_lambda=d["c_lambda"], iters=d["iters"])
    o.step(a, b, n_steps=steps - 1, tau_ff=tau)
    out.update(gcN=a, gvN=b, steps=np.int32(steps))
    if H is not None:
        out["H"] = H
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "contacts", int(d["ncontacts"].sum()), "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make("anymal_ground", "anymal_c_like.urdf", 48, 201, "ground", 0.45, 20.0, 10)
    make("anymal_heightmap", "anymal_c_like.urdf", 48, 202, "hm", 0.45, 20.0, 10)
    make("atlas_ground", "atlas_like.urdf", 24, 203, "ground", 0.9, 2.0, 5)

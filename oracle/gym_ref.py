"""ORACLE (test infrastructure): numpy restatement of the RaisimGym ANYmal task wrapped around the CPU oracle.

Restates, from SURVEY.md 3.1 [RECALL] (raisimGymTorch VectorizedEnvironment.hpp::perAgentStep and
envs/rsg_anymal/Environment.hpp; not in the reference snapshot):
    pTarget.tail(nJoints) = action * actionStd + actionMean ; setPdTarget(pTarget, 0)
    world.integrate() x (control_dt / simulation_dt)
    reward = torqueCoeff * |getGeneralizedForce()|^2 + forwardVelCoeff * min(4, bodyLinearVel.x)
    done   = any contact whose local body is not a foot  -> reward += terminalReward ; reset()
    observe() -> [z, R^T e_z, joint q, R^T v, R^T w, joint rates]
"""
import numpy as np


def quat_to_rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def observation(gc, gv):
    out = np.zeros((gc.shape[0], gc.shape[1] + gv.shape[1] - 3))
    for e in range(gc.shape[0]):
        R = quat_to_rot(gc[e, 3:7])
        out[e] = np.r_[gc[e, 2], R[2], gc[e, 7:], R.T @ gv[e, 0:3], R.T @ gv[e, 3:6], gv[e, 6:]]
    return out


class GymRef:
    def __init__(self, oracle, gc_init, gv_init, action_mean, action_std, foot_bodies, kp, kd,
                 torque_coeff=-4e-5, forward_vel_coeff=0.3, terminal_reward=-10.0):
        self.o, self.gc_init, self.gv_init = oracle, np.asarray(gc_init, float), np.asarray(gv_init, float)
        self.mean, self.std, self.feet = np.asarray(action_mean, float), np.asarray(action_std, float), set(int(b) for b in foot_bodies)
        self.kp, self.kd = kp, kd
        self.tc, self.fc, self.term = torque_coeff, forward_vel_coeff, terminal_reward

    def reset(self, n):
        self.gc = np.tile(self.gc_init, (n, 1)); self.gv = np.tile(self.gv_init, (n, 1))
        self.pt = self.gc.copy()
        self.o.reset_warm_start(n)

    def step(self, action, substeps):
        n = self.gc.shape[0]
        self.pt[:, 7:] = action * self.std + self.mean
        d = self.o.step(self.gc, self.gv, n_steps=substeps, ptarget=self.pt, vtarget=np.zeros_like(self.gv), kp=self.kp, kd=self.kd, debug=True)
        reward, done = np.zeros(n), np.zeros(n, bool)
        for e in range(n):
            R = quat_to_rot(self.gc[e, 3:7])
            vb = R.T @ self.gv[e, 0:3]
            reward[e] = self.tc * float(d["tau_applied"][e] @ d["tau_applied"][e]) + self.fc * min(4.0, vb[0])
            K = d["ncontacts"][e]
            done[e] = any(int(b) not in self.feet for b in d["c_body"][e, :K])
            if done[e]:
                reward[e] += self.term
                self.gc[e] = self.gc_init; self.gv[e] = self.gv_init; self.pt[e] = self.gc_init
        return observation(self.gc, self.gv), reward, done, d

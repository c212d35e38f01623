"""Shared test helpers: independent numpy restatements used to pin the oracle."""
import numpy as np

ANYMAL_GC0 = np.array([0, 0, 0.60, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8], dtype=np.float64)

PENDULUM_URDF = """<robot name="pend">
  <link name="world"/>
  <link name="l1"><inertial><origin xyz="0 0 -0.5"/><mass value="2.0"/><inertia ixx="0.1" ixy="0" ixz="0" iyy="0.1" iyz="0" izz="0.01"/></inertial></link>
  <joint name="j1" type="revolute"><parent link="world"/><child link="l1"/><origin xyz="0 0 1.0"/><axis xyz="0 1 0"/></joint>
  <link name="l2"><inertial><origin xyz="0 0 -0.3"/><mass value="1.0"/><inertia ixx="0.03" ixy="0" ixz="0" iyy="0.03" iyz="0" izz="0.005"/></inertial></link>
  <joint name="j2" type="revolute"><parent link="l1"/><child link="l2"/><origin xyz="0 0.1 -1.0"/><axis xyz="0 1 0"/></joint>
  <link name="l3"><inertial><origin xyz="0.1 0 0"/><mass value="0.5"/><inertia ixx="0.01" ixy="0.001" ixz="0" iyy="0.02" iyz="0" izz="0.02"/></inertial></link>
  <joint name="j3" type="prismatic"><parent link="l2"/><child link="l3"/><origin xyz="0 0 -0.6" rpy="0.3 0.2 0.1"/><axis xyz="1 0 0"/></joint>
</robot>"""

SPHERE_URDF = """<robot name="ball">
  <link name="ball"><inertial><origin xyz="0 0 0"/><mass value="2.0"/><inertia ixx="0.008" ixy="0" ixz="0" iyy="0.008" iyz="0" izz="0.008"/></inertial>
  <collision><origin xyz="0 0 0"/><geometry><sphere radius="0.1"/></geometry></collision></link>
</robot>"""

BOX_URDF = """<robot name="box">
  <link name="box"><inertial><origin xyz="0 0 0"/><mass value="3.0"/><inertia ixx="0.02" ixy="0" ixz="0" iyy="0.03" iyz="0" izz="0.04"/></inertial>
  <collision><origin xyz="0 0 0"/><geometry><box size="0.4 0.3 0.2"/></geometry></collision></link>
</robot>"""


def quat_to_rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axis_angle(a, q):
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def fk_numpy(t, gc):
    """Independent FK: returns R[nb,3,3], p[nb,3], world joint axes a[nb,3]."""
    nb = t["nb"]
    R, p, a = np.zeros((nb, 3, 3)), np.zeros((nb, 3)), np.zeros((nb, 3))
    for i in range(nb):
        pr = t["parent"][i]
        if pr < 0:
            if t["floating"]:
                R[i], p[i] = quat_to_rot(gc[3:7]), gc[0:3]
            else:
                R[i], p[i] = t["jrot"][i].reshape(3, 3), t["jpos"][i]
            continue
        Rj = R[pr] @ t["jrot"][i].reshape(3, 3)
        a[i] = Rj @ t["axis"][i]
        p[i] = p[pr] + R[pr] @ t["jpos"][i]
        q = gc[t["qidx"][i]]
        if t["jtype"][i] == 1:
            R[i] = Rj @ axis_angle(t["axis"][i], q)
        else:
            R[i] = Rj
            p[i] = p[i] + q * a[i]
    return R, p, a


def body_jacobians(t, gc):
    """Per body: Jv (COM linear velocity) and Jw (angular velocity), each [3,nv]; the textbook
    M = sum_b m Jv^T Jv + Jw^T (R I R^T) Jw identity, independent of the oracle's CRBA."""
    R, p, a = fk_numpy(t, gc)
    nb, nv = t["nb"], t["nv"]
    out = []
    for b in range(nb):
        c = p[b] + R[b] @ t["com"][b]
        Jv, Jw = np.zeros((3, nv)), np.zeros((3, nv))
        if t["floating"]:
            Jv[:, 0:3] = np.eye(3)
            Jv[:, 3:6] = -skew(c - p[0])
            Jw[:, 3:6] = np.eye(3)
        j = b
        while t["parent"][j] >= 0:
            vi = t["vidx"][j]
            if t["jtype"][j] == 1:
                Jv[:, vi] = np.cross(a[j], c - p[j])
                Jw[:, vi] = a[j]
            else:
                Jv[:, vi] = a[j]
            j = t["parent"][j]
        out.append((Jv, Jw, R[b], c))
    return out


def mass_matrix_numpy(t, gc):
    M = np.zeros((t["nv"], t["nv"]))
    for b, (Jv, Jw, Rb, c) in enumerate(body_jacobians(t, gc)):
        I6 = t["inertia"][b]
        I = np.array([[I6[0], I6[1], I6[2]], [I6[1], I6[3], I6[4]], [I6[2], I6[4], I6[5]]])
        M += t["mass"][b] * Jv.T @ Jv + Jw.T @ (Rb @ I @ Rb.T) @ Jw
    return M


def potential_energy(t, gc, g=9.81):
    R, p, _ = fk_numpy(t, gc)
    return sum(t["mass"][b] * g * (p[b] + R[b] @ t["com"][b])[2] for b in range(t["nb"]))


def integrate_gc(t, gc, dv, eps):
    """gc (+) eps*dv with the oracle's conventions (world-frame base angular velocity)."""
    out = gc.copy()
    if t["floating"]:
        out[0:3] += eps * dv[0:3]
        w = dv[3:6] * eps
        ang = np.linalg.norm(w)
        if ang > 0:
            dq = np.r_[np.cos(ang / 2), np.sin(ang / 2) * w / ang]
        else:
            dq = np.array([1.0, 0, 0, 0])
        q = gc[3:7]
        out[3:7] = np.array([dq[0] * q[0] - dq[1:] @ q[1:], *(dq[0] * q[1:] + q[0] * dq[1:] + np.cross(dq[1:], q[1:]))])
    for i in range(1, t["nb"]):
        out[t["qidx"][i]] += eps * dv[t["vidx"][i]]
    return out


def random_state(t, rng, n, pos_scale=0.0, vel_scale=1.0, joint_scale=0.5, base_z=0.6):
    gc = np.zeros((n, t["nq"])); gv = vel_scale * rng.standard_normal((n, t["nv"]))
    if t["floating"]:
        gc[:, 0:2] = pos_scale * rng.uniform(-1, 1, (n, 2))
        gc[:, 2] = base_z
        q = rng.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        gc[:, 3:7] = q
        gc[:, 7:] = joint_scale * rng.uniform(-1, 1, (n, t["nq"] - 7))
    else:
        gc[:] = joint_scale * rng.uniform(-1, 1, (n, t["nq"]))
    return gc, gv

# a link set as real robot descriptions have it: visuals, meshes, cylinders, gazebo / transmission tags
REALISTIC_URDF = """<?xml version="1.0" ?>
<!-- generated by xacro -->
<robot name="mini" xmlns:xacro="http://www.ros.org/wiki/xacro">
  <material name="grey"><color rgba="0.5 0.5 0.5 1"/></material>
  <link name="trunk">
    <visual><origin xyz="0 0 0" rpy="0 0 0"/><geometry><mesh filename="package://x/meshes/trunk.dae" scale="1 1 1"/></geometry><material name="grey"/></visual>
    <collision><geometry><box size="0.4 0.2 0.1"/></geometry></collision>
    <collision><geometry><mesh filename="package://x/meshes/trunk_col.stl"/></geometry></collision>
    <inertial><origin xyz="0.01 0 0"/><mass value="5"/><inertia ixx="0.02" ixy="0" ixz="0" iyy="0.06" iyz="0" izz="0.07"/></inertial>
  </link>
  <link name="imu_link"/>
  <joint name="imu_joint" type="fixed"><parent link="trunk"/><child link="imu_link"/><origin rpy="0 0 0" xyz="0 0 0.05"/></joint>
  <link name="leg">
    <visual><geometry><cylinder length="0.3" radius="0.02"/></geometry></visual>
    <collision><origin xyz="0 0 -0.15" rpy="0 0 0"/><geometry><cylinder length="0.3" radius="0.025"/></geometry></collision>
    <inertial><origin xyz="0 0 -0.15"/><mass value="0.8"/><inertia ixx="0.006" ixy="0" ixz="0" iyy="0.006" iyz="0" izz="0.0003"/></inertial>
  </link>
  <joint name="hip" type="continuous"><parent link="trunk"/><child link="leg"/><origin xyz="0.15 0.1 0"/><axis xyz="0 1 0"/>
    <dynamics damping="0.0" friction="0.0"/><limit effort="30" velocity="20"/></joint>
  <transmission name="t"><type>transmission_interface/SimpleTransmission</type><joint name="hip"><hardwareInterface>hardware_interface/EffortJointInterface</hardwareInterface></joint></transmission>
  <gazebo reference="leg"><mu1>0.8</mu1><self_collide>1</self_collide></gazebo>
</robot>"""

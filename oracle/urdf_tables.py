"""ORACLE (test infrastructure, not product): URDF -> model tables, restated in Python/numpy.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference behaviour restated ([RECALL]; the reference snapshot holds
no source for it -- `/root/reference/.SUBMODULES.json:2` records "bytes": 0):
`raisim::World::addArticulatedSystem(urdf)` parses the URDF, merges links connected by
fixed joints into their parent body, orders the movable bodies depth-first, and produces
per-body constants {parent, joint type, joint axis, joint origin in parent, mass, COM,
inertia} plus collision bodies {shape, size, pose in body}.

This file is an independent restatement (xml.etree) of what the product's C++ loader
(raisimlib_b200/csrc/model.cpp) does, so that the two can be compared table by table.

Conventions fixed here (parity unpinned -- no reference artefact can confirm them):
  * body 0 is the URDF root link; a root link named "world" means a fixed base
  * children are visited in the order their joints appear in the file (DFS pre-order)
  * body frame == child-link frame of the body's movable joint
  * inertia stored as (xx, xy, xz, yy, yz, zz) about the body COM, in body axes
  * collision shapes expand to candidate points (sphere -> 1, capsule -> 2 end spheres,
    box -> 8 zero-radius corners, corner k has sign bits x=k&1, y=k&2, z=k&4;
    cylinder -> 4 zero-radius rim points per end cap, point k at angle 90deg*(k&3), cap z sign k&4;
    mesh collision bodies are skipped)
"""
import xml.etree.ElementTree as ET
import numpy as np

JT_FIXED, JT_REVOLUTE, JT_PRISMATIC, JT_FLOATING = 0, 1, 2, 3
CT_SPHERE, CT_BOX, CT_CAPSULE, CT_CYLINDER = 0, 1, 2, 3


def rpy_to_rot(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    # URDF fixed-axis roll-pitch-yaw: R = Rz(y) Ry(p) Rx(r)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _vec(s, n=3, default=None):
    if s is None:
        return np.array(default if default is not None else [0.0] * n, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    assert v.shape == (n,)
    return v


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.zeros(3), np.eye(3)
    return _vec(o.get("xyz")), rpy_to_rot(_vec(o.get("rpy")))


def _sym(I6):
    xx, xy, xz, yy, yz, zz = I6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])


def _skew2(d):
    # -[d]x[d]x  = (d.d) 1 - d d^T
    return np.dot(d, d) * np.eye(3) - np.outer(d, d)


class _Body:
    def __init__(self, name):
        self.name = name
        self.mass = 0.0
        self.com = np.zeros(3)
        self.I = np.zeros((3, 3))
        self.colls = []      # (type, size[3], pos[3], rot[3,3], link name)
        self.frames = []     # (name, pos, rot)

    def add_inertia(self, m, c, I):
        if m <= 0.0:
            return
        mt = self.mass + m
        cn = (self.mass * self.com + m * c) / mt
        self.I = self.I + self.mass * _skew2(self.com - cn) + I + m * _skew2(c - cn)
        self.mass, self.com = mt, cn


def load_tables(path_or_xml):
    if path_or_xml.lstrip().startswith("<"):
        root = ET.fromstring(path_or_xml)
    else:
        root = ET.parse(path_or_xml).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_links = {j.find("child").get("link") for j in joints}
    roots = [n for n in links if n not in child_links]
    assert len(roots) == 1, "URDF must have exactly one root link"
    root_name = roots[0]
    children = {}
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)

    bodies, parent, jtype, jpos, jrot, axis, jname, jlimit = [], [], [], [], [], [], [], []
    jeffort = []

    def absorb(body, link_name, T_pos, T_rot):
        """merge link `link_name` (frame at T in body coords) into body, then recurse."""
        l = links[link_name]
        body.frames.append((link_name, T_pos.copy(), T_rot.copy()))
        ine = l.find("inertial")
        if ine is not None:
            cpos, crot = _origin(ine)
            m = float(ine.find("mass").get("value"))
            it = ine.find("inertia")
            I6 = [float(it.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")]
            Rl = T_rot @ crot
            body.add_inertia(m, T_pos + T_rot @ cpos, Rl @ _sym(I6) @ Rl.T)
        for c in l.findall("collision"):
            cpos, crot = _origin(c)
            g = c.find("geometry")
            pos, rot = T_pos + T_rot @ cpos, T_rot @ crot
            if g.find("sphere") is not None:
                body.colls.append((CT_SPHERE, [float(g.find("sphere").get("radius")), 0, 0], pos, rot, link_name))
            elif g.find("box") is not None:
                sz = _vec(g.find("box").get("size"))
                body.colls.append((CT_BOX, list(0.5 * sz), pos, rot, link_name))
            elif g.find("capsule") is not None:
                e = g.find("capsule")
                body.colls.append((CT_CAPSULE, [float(e.get("radius")), 0.5 * float(e.get("length")), 0], pos, rot, link_name))
            elif g.find("cylinder") is not None:
                e = g.find("cylinder")
                body.colls.append((CT_CYLINDER, [float(e.get("radius")), 0.5 * float(e.get("length")), 0], pos, rot, link_name))
            elif g.find("mesh") is not None:
                continue                                 # mesh collision bodies are outside this path
            else:
                raise ValueError("unsupported collision geometry in link " + link_name)
        my_index = bodies.index(body)
        for j in children.get(link_name, []):
            opos, orot = _origin(j)
            t = j.get("type")
            child = j.find("child").get("link")
            if t == "fixed":
                absorb(body, child, T_pos + T_rot @ opos, T_rot @ orot)
            elif t in ("revolute", "continuous", "prismatic"):
                nb = _Body(child)
                bodies.append(nb)
                parent.append(my_index)
                jtype.append(JT_PRISMATIC if t == "prismatic" else JT_REVOLUTE)
                jpos.append(T_pos + T_rot @ opos)
                jrot.append(T_rot @ orot)
                a = _vec(j.find("axis").get("xyz")) if j.find("axis") is not None else np.array([1.0, 0, 0])
                axis.append(a / np.linalg.norm(a))
                jname.append(j.get("name"))
                lim = j.find("limit")
                jlimit.append([float(lim.get("lower", "-1e30")), float(lim.get("upper", "1e30"))]
                              if (lim is not None and t != "continuous") else [-1e30, 1e30])
                eff = float(lim.get("effort", "0")) if lim is not None else 0.0
                jeffort.append(eff if eff > 0 else 1e30)          # <limit effort>: 0 or absent = unlimited
                absorb(nb, child, np.zeros(3), np.eye(3))
            else:
                raise ValueError("unsupported joint type " + t)

    floating = root_name != "world"
    b0 = _Body(root_name)
    bodies.append(b0)
    parent.append(-1)
    jtype.append(JT_FLOATING if floating else JT_FIXED)
    jpos.append(np.zeros(3)); jrot.append(np.eye(3)); axis.append(np.array([0.0, 0, 1])); jname.append("root")
    jlimit.append([-1e30, 1e30]); jeffort.append(1e30)
    absorb(b0, root_name, np.zeros(3), np.eye(3))

    nb = len(bodies)
    qidx, vidx, depth = np.zeros(nb, np.int32), np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    nq = 7 if floating else 0
    nv = 6 if floating else 0
    for i in range(1, nb):
        qidx[i], vidx[i] = nq, nv
        nq += 1; nv += 1
        depth[i] = depth[parent[i]] + 1
    I6 = np.array([[b.I[0, 0], b.I[0, 1], b.I[0, 2], b.I[1, 1], b.I[1, 2], b.I[2, 2]] for b in bodies])
    t = dict(
        nb=nb, nq=nq, nv=nv, floating=int(floating),
        parent=np.array(parent, np.int32), jtype=np.array(jtype, np.int32), qidx=qidx, vidx=vidx, depth=depth,
        jpos=np.array(jpos), jrot=np.array([r.reshape(9) for r in jrot]), axis=np.array(axis),
        mass=np.array([b.mass for b in bodies]), com=np.array([b.com for b in bodies]), inertia=I6,
        jlimit=np.array(jlimit), jeffort=np.array(jeffort, np.float64),
        body_names=[b.name for b in bodies], joint_names=jname,
    )
    # collision bodies and their candidate points
    cbody, ctype, csize, cpos, crot, cname = [], [], [], [], [], []
    pt_body, pt_pos, pt_rad, pt_coll, pt_feat = [], [], [], [], []
    for bi, b in enumerate(bodies):
        for (ty, size, pos, rot, lname) in b.colls:
            ci = len(cbody)
            cbody.append(bi); ctype.append(ty); csize.append(size); cpos.append(pos); crot.append(rot.reshape(9)); cname.append(lname)
            if ty == CT_SPHERE:
                pts = [(pos, size[0])]
            elif ty == CT_CAPSULE:
                pts = [(pos + rot @ np.array([0, 0, -size[1]]), size[0]), (pos + rot @ np.array([0, 0, size[1]]), size[0])]
            elif ty == CT_CYLINDER:
                cx, sy = [1.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, -1.0]
                pts = [(pos + rot @ np.array([size[0] * cx[k & 3], size[0] * sy[k & 3], size[1] if k & 4 else -size[1]]), 0.0) for k in range(8)]
            else:
                pts = []
                for k in range(8):
                    sgn = np.array([1.0 if k & 1 else -1.0, 1.0 if k & 2 else -1.0, 1.0 if k & 4 else -1.0])
                    pts.append((pos + rot @ (sgn * np.array(size)), 0.0))
            for f, (pp, rr) in enumerate(pts):
                pt_body.append(bi); pt_pos.append(pp); pt_rad.append(rr); pt_coll.append(ci); pt_feat.append(f)
    pt_type = [0] * len(pt_body); pt_pos2 = [p.copy() for p in pt_pos]
    # shape features after the points, in collision-body order: one segment (type 1) per capsule / cylinder axis, one box-face
    # feature (type 2) per box -- what can touch a height map where no end sphere / corner does
    for ci in range(len(cbody)):
        ty, size, pos, rot = ctype[ci], csize[ci], cpos[ci], crot[ci].reshape(3, 3)
        if ty == CT_SPHERE:
            continue
        pt_body.append(cbody[ci]); pt_coll.append(ci); pt_feat.append(0)
        if ty == CT_BOX:
            pt_type.append(2); pt_pos.append(np.array(pos, float)); pt_pos2.append(np.array(pos, float)); pt_rad.append(0.0)
        else:
            ax = rot[:, 2] * size[1]
            pt_type.append(1); pt_pos.append(np.array(pos, float) - ax); pt_pos2.append(np.array(pos, float) + ax); pt_rad.append(size[0])
            if ty == CT_CYLINDER:      # two rim features (type 3): the lowest point of each cap's rim circle
                for cap, sg in enumerate((-1.0, 1.0)):
                    pt_body.append(cbody[ci]); pt_coll.append(ci); pt_feat.append(cap); pt_type.append(3)
                    pt_pos.append(np.array(pos, float) + sg * ax); pt_pos2.append(np.array(pos, float) - sg * ax); pt_rad.append(size[0])
    t.update(
        ncoll=len(cbody), cbody=np.array(cbody, np.int32), ctype=np.array(ctype, np.int32),
        csize=np.array(csize, np.float64).reshape(-1, 3), cpos=np.array(cpos, np.float64).reshape(-1, 3),
        crot=np.array(crot, np.float64).reshape(-1, 9), coll_names=cname,
        npts=len(pt_body), pt_body=np.array(pt_body, np.int32), pt_pos=np.array(pt_pos, np.float64).reshape(-1, 3),
        pt_rad=np.array(pt_rad, np.float64), pt_coll=np.array(pt_coll, np.int32), pt_feat=np.array(pt_feat, np.int32),
        pt_type=np.array(pt_type, np.int32), pt_pos2=np.array(pt_pos2, np.float64).reshape(-1, 3),
    )
    return t

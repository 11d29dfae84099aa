"""Host-side scene compiler: compiled assets (json / npz) -> the POD structs of the C ABI.

The URDF link table written by tools/compile_assets.py keeps pybullet's link order and frames.  Here the
fixed joints are folded into their moving parent ("reduced model"), which is what the CUDA kernels run on:
a UR5 becomes 6 bodies, an MG400 8.  Frames the reference reads through getLinkState()[0:2] (the
INERTIAL frame of tcp_link and of <sensor>_body_link, robots/arms/base_robot_arm.py:136-151,
sensors/tactile_sensor.py:150-187) are kept as rigid offsets on their owning body.
"""
import json
import os

import numpy as np

from . import _lib as L

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def quat_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_from_euler(rpy):
    """p.getQuaternionFromEuler"""
    hr, hp, hy = rpy[0] * 0.5, rpy[1] * 0.5, rpy[2] * 0.5
    cr, sr, cp, sp, cy, sy = np.cos(hr), np.sin(hr), np.cos(hp), np.sin(hp), np.cos(hy), np.sin(hy)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def load_model_json(arm, sensor, typ):
    path = os.path.join(ASSETS, "models", "%s_%s_%s.json" % (arm, typ, sensor))
    if not os.path.isfile(path):
        raise ValueError("no compiled model for arm_type=%r tactile_sensor_name=%r type=%r" % (arm, sensor, typ))
    with open(path) as f:
        return json.load(f)


def load_sensor_json(sensor, typ):
    with open(os.path.join(ASSETS, "sensors.json")) as f:
        sj = json.load(f)
    if sensor not in sj or typ not in sj[sensor]["types"]:
        raise ValueError("unknown tactile sensor %r / %r" % (sensor, typ))
    return sj[sensor]


def load_rest_pose(env, arm, sensor, typ, control_links):
    with open(os.path.join(ASSETS, "rest_poses.json")) as f:
        rp = json.load(f)[env][arm]
    rp = rp[sensor] if sensor in rp else rp
    if typ not in rp:
        # the reference's tables decide which pairings exist (e.g. surface_follow has MG400 rest poses for `forward` only):
        # its own constructor dies with the same KeyError (rest_poses_dict[arm][sensor][type])
        raise KeyError("the reference has no %s rest pose for %s + %s / %s" % (env, arm, sensor, typ))
    return np.asarray(rp[typ], dtype=np.float64)[control_links].copy()


def load_refimg(sensor, typ, S):
    path = os.path.join(ASSETS, "refimg", "%s_%s_%d.npz" % (sensor, typ, S))
    if not os.path.isfile(path):
        raise ValueError("no reference images for %s/%s at %dx%d" % (sensor, typ, S, S))
    d = np.load(path)
    return (np.ascontiguousarray(d["nodef_dep"], dtype=np.float32), np.ascontiguousarray(d["nodef_gray"], dtype=np.float32),
            np.ascontiguousarray(d["border_mask"], dtype=np.uint8))


def load_stimulus(name):
    return np.ascontiguousarray(np.load(os.path.join(ASSETS, "stimuli", name + ".npz"))["tris"], dtype=np.float64)


def merge_coplanar(tris, tol=1e-9):
    """Stimulus triangles [T,3,3] -> convex planar primitives [P,4,3] + vertex counts [P].

    Two triangles that share an edge, lie in one plane and whose union is a convex quadrilateral become one quad
    (the OBJ quads of the box stimuli, fan-triangulated by the loader, come back as quads).  The union has the same
    pixel coverage and the same plane, so the render is unchanged; the raster just sees half as many primitives.
    """
    tris = np.asarray(tris, dtype=np.float64)
    used = np.zeros(len(tris), dtype=bool)
    prims, nv = [], []

    def normal(t):
        n = np.cross(t[1] - t[0], t[2] - t[0])
        l = np.linalg.norm(n)
        return n / l if l > 0 else n

    for i in range(len(tris)):
        if used[i]:
            continue
        used[i] = True
        a, na = tris[i], normal(tris[i])
        quad = None
        for j in range(i + 1, len(tris)):
            if used[j]:
                continue
            b, nb = tris[j], normal(tris[j])
            scale = max(np.abs(a).max(), 1e-12)
            if np.linalg.norm(na - nb) > 1e-9 or abs(np.dot(na, b[0] - a[0])) > tol * max(scale, 1.0):
                continue
            # shared edge (as vertex pairs)
            shared = [(p, q) for p in range(3) for q in range(3) if np.allclose(a[p], b[q], atol=tol, rtol=0)]
            if len(shared) != 2:
                continue
            pa = [p for p, _ in shared]
            qb = [q for _, q in shared]
            oa = [p for p in range(3) if p not in pa][0]          # a's vertex not on the shared edge
            ob = [q for q in range(3) if q not in qb][0]
            # walk a's vertices in order, inserting b's free vertex between the two shared ones
            k = (oa + 1) % 3
            cand = np.array([a[oa], a[k], b[ob], a[(k + 1) % 3]])
            # convex and consistently oriented?
            ok = True
            for m in range(4):
                e1 = cand[(m + 1) % 4] - cand[m]
                e2 = cand[(m + 2) % 4] - cand[(m + 1) % 4]
                if np.dot(np.cross(e1, e2), na) <= 1e-14:
                    ok = False
            if ok:
                quad = cand
                used[j] = True
                break
        if quad is not None:
            prims.append(quad); nv.append(4)
        else:
            prims.append(np.vstack([a, a[2:3]])); nv.append(3)
    return np.ascontiguousarray(np.array(prims), dtype=np.float64), np.ascontiguousarray(np.array(nv), dtype=np.int32)


def reduce_model(mj, sensor, cam_pos, cam_rpy):
    """Fold fixed joints; returns (TgArm, control_links)."""
    links = mj["links"]
    n = len(links)
    names = [l["link_name"] for l in links]
    moving = [i for i, l in enumerate(links) if l["joint_type"] == 1]
    body_of_link = [-1] * n           # owning moving body (index into `moving`), -1 = static base
    T_body_link = [None] * n          # (R, t): link frame expressed in its owning body's frame (q = 0)
    arm = L.TgArm()
    arm.nb = len(moving)
    if arm.nb > L.TG_MAXB:
        raise ValueError("arm has %d dofs (max %d)" % (arm.nb, L.TG_MAXB))
    arm.topo = L.TG_TOPO_CHAIN6 if arm.nb == 6 else L.TG_TOPO_MG400
    expected_parent = {L.TG_TOPO_CHAIN6: [-1, 0, 1, 2, 3, 4], L.TG_TOPO_MG400: [-1, 0, 1, 2, 3, 0, 5, 6]}[arm.topo]
    for i, l in enumerate(links):
        p = l["parent"]
        Rj, tj = rpy_to_mat(l["joint_rpy"]), np.array(l["joint_xyz"], dtype=np.float64)
        if p < 0:
            Rp, tp, pb = np.eye(3), np.zeros(3), -1
        else:
            Rp, tp, pb = T_body_link[p][0], T_body_link[p][1], body_of_link[p]
        R, t = Rp @ Rj, tp + Rp @ tj      # child link frame in the parent's owning body frame
        if l["joint_type"] == 1:
            b = moving.index(i)
            if pb != expected_parent[b]:
                raise ValueError("unexpected arm topology at %s" % names[i])
            a = np.array(l["axis"], dtype=np.float64)
            a = a / np.linalg.norm(a)
            for c in range(3):
                arm.jpos[b][c] = t[c]
                arm.axis[b][c] = a[c]
            for c in range(9):
                arm.jrot[b][c] = R.reshape(-1)[c]
            body_of_link[i] = b
            T_body_link[i] = (np.eye(3), np.zeros(3))
        elif l["joint_type"] == 0:
            body_of_link[i] = pb
            T_body_link[i] = (R, t)
        else:
            raise ValueError("joint type of %s not supported" % names[i])
    # composite inertias + damping sub-links
    nsub = 0
    for b in range(arm.nb):
        parts = []
        arm.sub_start[b] = nsub
        for i, l in enumerate(links):
            if body_of_link[i] != b or l["mass"] <= 0:
                continue
            R, t = T_body_link[i]
            Rin = R @ rpy_to_mat(l["inertial_rpy"])
            c = t + R @ np.array(l["inertial_xyz"], dtype=np.float64)
            I = Rin @ np.diag(l["inertia_diag"]) @ Rin.T
            parts.append((l["mass"], c, I))
            if nsub >= L.TG_MAXSUB:
                raise ValueError("too many mass-carrying links")
            arm.sub_body[nsub] = b
            arm.sub_mass[nsub] = l["mass"]
            for k in range(3):
                arm.sub_com[nsub][k] = c[k]
                arm.sub_inertia[nsub][k] = l["inertia_diag"][k]
            for k in range(9):
                arm.sub_rot[nsub][k] = Rin.reshape(-1)[k]
            nsub += 1
        m = sum(p[0] for p in parts)
        if m <= 0:
            raise ValueError("body %d has no mass" % b)
        com = sum(p[0] * p[1] for p in parts) / m
        I = np.zeros((3, 3))
        for mi, ci, Ii in parts:
            d = ci - com
            I += Ii + mi * (d.dot(d) * np.eye(3) - np.outer(d, d))
        arm.mass[b] = m
        for k in range(3):
            arm.com[b][k] = com[k]
        for k, (r, c) in enumerate([(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]):
            arm.inertia[b][k] = I[r, c]
    arm.nsub = nsub
    for b in range(arm.nb, L.TG_MAXB + 1):
        arm.sub_start[b] = nsub

    def inertial_frame(link_name):
        i = names.index(link_name)
        R, t = T_body_link[i]
        l = links[i]
        return body_of_link[i], t + R @ np.array(l["inertial_xyz"], dtype=np.float64), R @ rpy_to_mat(l["inertial_rpy"])

    b, t, R = inertial_frame("tcp_link")
    arm.tcp_body = b
    for k in range(3):
        arm.tcp_pos[k] = t[k]
    for k in range(9):
        arm.tcp_rot[k] = R.reshape(-1)[k]
    b, t, R = inertial_frame(sensor + "_body_link")
    Rc = quat_to_mat(quat_from_euler(cam_rpy))   # multiplyTransforms(body, cam) (tactile_sensor.py:184-187)
    tc, Rc = t + R @ np.array(cam_pos, dtype=np.float64), R @ Rc
    arm.cam_body = b
    for k in range(3):
        arm.cam_pos[k] = tc[k]
    for k in range(9):
        arm.cam_rot[k] = Rc.reshape(-1)[k]
    A = _align_body_frames(arm)
    # kept for link_points_in_body(): URDF link frames in their (world-aligned) body frame
    arm._link_frames = {names[i]: (body_of_link[i], A[body_of_link[i]] @ T_body_link[i][0], A[body_of_link[i]] @ T_body_link[i][1])
                        for i in range(n) if body_of_link[i] >= 0}
    return arm, moving


def link_points_in_body(arm, link_name, pts):
    """Points given in a URDF link frame (e.g. the tip core hull, assets/models/*_meshes.npz) -> the frame of the reduced
    body that carries the link.  Returns (body index, points [K,3])."""
    b, R, t = arm._link_frames[link_name]
    return b, np.ascontiguousarray(np.asarray(pts, dtype=np.float64) @ R.T + t)


def load_tip_hull(arm_type, sensor, typ):
    d = np.load(os.path.join(ASSETS, "models", "%s_%s_%s_meshes.npz" % (arm_type, typ, sensor)))
    return np.ascontiguousarray(d["tip_core_hull"], dtype=np.float64)


def load_object_json(name):
    with open(os.path.join(ASSETS, "objects", name + ".json")) as f:
        return json.load(f)


def _align_body_frames(arm):
    """Re-express every body-fixed quantity in a frame that is world-aligned at q = 0.

    With A_i = A_parent . jrot_i, the rotated frame R_i' = R_i A_i^T obeys R_i' = R_parent' . Rot(A_i axis_i, q_i):
    the constant joint rotation disappears from the forward kinematics (one 3x3 product per body per substep
    saved); vectors become A_i v, tensors A_i I A_i^T, attached frames A_i R.  jrot becomes the identity."""
    nb = arm.nb
    parents = {0: [-1, 0, 1, 2, 3, 4], 1: [-1, 0, 1, 2, 3, 0, 5, 6]}[arm.topo]
    A = []
    for b in range(nb):
        Ap = np.eye(3) if parents[b] < 0 else A[parents[b]]
        A.append(Ap @ np.array(arm.jrot[b][:]).reshape(3, 3))
    def setv(dst, v):
        for k in range(len(v)):
            dst[k] = float(v[k])
    for b in range(nb):
        Ap = np.eye(3) if parents[b] < 0 else A[parents[b]]
        setv(arm.jpos[b], Ap @ np.array(arm.jpos[b][:]))
        setv(arm.jrot[b], np.eye(3).reshape(-1))
        setv(arm.axis[b], A[b] @ np.array(arm.axis[b][:]))
        setv(arm.com[b], A[b] @ np.array(arm.com[b][:]))
        I6 = arm.inertia[b][:]
        I = np.array([[I6[0], I6[1], I6[2]], [I6[1], I6[3], I6[4]], [I6[2], I6[4], I6[5]]])
        I = A[b] @ I @ A[b].T
        setv(arm.inertia[b], [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]])
    for s_ in range(arm.nsub):
        Ab = A[arm.sub_body[s_]]
        setv(arm.sub_com[s_], Ab @ np.array(arm.sub_com[s_][:]))
        setv(arm.sub_rot[s_], (Ab @ np.array(arm.sub_rot[s_][:]).reshape(3, 3)).reshape(-1))
    setv(arm.tcp_pos, A[arm.tcp_body] @ np.array(arm.tcp_pos[:]))
    setv(arm.tcp_rot, (A[arm.tcp_body] @ np.array(arm.tcp_rot[:]).reshape(3, 3)).reshape(-1))
    setv(arm.cam_pos, A[arm.cam_body] @ np.array(arm.cam_pos[:]))
    setv(arm.cam_rot, (A[arm.cam_body] @ np.array(arm.cam_rot[:]).reshape(3, 3)).reshape(-1))
    return A


def default_physics(substeps=24, gravity=(0.0, 0.0, -9.81)):
    """base_tactile_env.py:125-130, base_robot_arm.py:24-25, ur5.py:19-21."""
    p = L.TgPhysics()
    for c in range(3):
        p.gravity[c] = gravity[c]
    p.dt = 1.0 / 240.0
    p.solver_iters = 150
    p.substeps = substeps
    p.lin_damping, p.ang_damping, p.joint_damping = 0.04, 0.04, 0.01
    p.max_force, p.pos_gain, p.vel_gain = 1000.0, 1.0, 1.0
    p.solver_residual_threshold = 1e-7
    p.blocking_force = 100000.0
    p.gravity_comp = 1
    return p


def convex_parts(prims, prim_nv, tol=1e-9):
    """Group the stimulus primitives into CONVEX parts for the scanline raster (csrc/tg_raster_scan.cuh).

    Primitives that share vertices form one component (the edge stimulus and the cube are one box each, the pole a base plate
    and a post); a component is a convex part when it is closed and every vertex of it lies on or inside every face plane.
    Returns (part id per primitive [P] int32, part centroids [n_parts, 3]) - or (None, None) when some component is not
    convex (the general raster kernel renders such scenes; none of the reference's stimuli for the tasks built here is)."""
    prims = np.asarray(prims, dtype=np.float64)
    P = len(prims)
    parent = list(range(P))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    verts = [prims[t][: int(prim_nv[t])] for t in range(P)]
    for a in range(P):
        for b in range(a + 1, P):
            if any(np.allclose(va, vb, atol=tol, rtol=0) for va in verts[a] for vb in verts[b]):
                parent[find(a)] = find(b)
    roots = sorted({find(t) for t in range(P)})
    part = np.array([roots.index(find(t)) for t in range(P)], dtype=np.int32)
    cens = np.zeros((len(roots), 3))
    for k in range(len(roots)):
        pts = np.concatenate([verts[t] for t in range(P) if part[t] == k])
        uniq = np.unique(np.round(pts / tol).astype(np.int64), axis=0) * tol
        cens[k] = uniq.mean(axis=0)
        scale = max(np.abs(pts - cens[k]).max(), 1e-12)
        # every edge of the component must be shared by exactly two faces (closed surface)
        edges = {}
        for t in range(P):
            if part[t] != k:
                continue
            v = verts[t]
            for m in range(len(v)):
                key = tuple(sorted((tuple(np.round(v[m] / tol).astype(np.int64)), tuple(np.round(v[(m + 1) % len(v)] / tol).astype(np.int64)))))
                edges[key] = edges.get(key, 0) + 1
        if any(c != 2 for c in edges.values()):
            return None, None
        for t in range(P):
            if part[t] != k:
                continue
            v = verts[t]
            n = np.cross(v[1] - v[0], v[2] - v[0])
            n /= np.linalg.norm(n)
            if np.dot(n, cens[k] - v[0]) > 0:
                n = -n                                   # outward
            if ((pts - v[0]) @ n).max() > 1e-7 * scale:
                return None, None                        # a vertex outside this face's plane: not convex
    return part, cens

#!/usr/bin/env python3
"""Offline, deterministic asset compiler: reference URDF / OBJ / STL / .npy  ->  flat scene data.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU box):

    python tools/compile_assets.py            # rewrites tactile_gym_b200/assets/*

What it reproduces of PyBullet's URDF loader (all [EXT] items are
restated from the public Bullet3 importer and are NOT verifiable here - no pybullet in this image):

  * link / joint indices = pre-order depth-first walk of the URDF tree, children in file order
    (rest_poses comments, reference rl_envs/exploration/edge_follow/rest_poses.py:8-18,99-110).
  * numeric attributes are parsed with C `atof` semantics: the longest valid float prefix is used, so
    "4.96e-09+0.035" -> 4.96e-09 (reference ur5_with_standard_digit.urdf:279).
  * angles are the file's literals (1.57 / 3.14), never pi.
  * mesh file names resolve against the URDF's directory, then each ancestor directory
    (robot.py:102-110 loads robot_assets/ur5/tactip/*.urdf whose "visual/base.obj" lives in
    robot_assets/ur5/visual/), `package://<name>/` is stripped.
  * [EXT] loadURDF without URDF_USE_INERTIA_FROM_FILE (robot.py:108-110 passes no flags) ignores the
    <inertia> element: the inertial frame is the <inertial><origin>, and the diagonal inertia is that
    of the collision shapes' AABB box (in the inertial frame) grown by the collision margins.  Both
    the URDF tensor and the AABB estimate are written out; the engine uses the AABB one.

The outputs are plain data (json + npz); no reference source code is copied.
"""
import json
import os
import re
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = os.environ.get("TG_REFERENCE", "/root/reference")
ASSETS = os.path.join(REF, "tactile_gym", "assets")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tactile_gym_b200", "assets")

_FLOAT_PREFIX = re.compile(r"^\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)")

URDF_COLLISION_MARGIN = 0.001  # [EXT] gUrdfDefaultCollisionMargin


def atof(tok):
    """C atof: longest valid prefix, 0.0 if none."""
    m = _FLOAT_PREFIX.match(tok)
    return float(m.group(0)) if m else 0.0


def vec(s, n=3, default=0.0):
    if s is None:
        return [default] * n
    toks = s.split()
    out = [atof(t) for t in toks]
    assert len(out) == n, (s, n)
    return out


def rpy_to_mat(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix R = Rz(y) Ry(p) Rx(r)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def resolve_mesh(urdf_path, fname):
    if fname.startswith("package://"):
        fname = fname[len("package://"):]
        # package://mg400/meshes/x.STL -> try with and without the package name
        cands = [fname, fname.split("/", 1)[1] if "/" in fname else fname]
    else:
        cands = [fname]
    d = os.path.dirname(os.path.abspath(urdf_path))
    while True:
        for c in cands:
            p = os.path.normpath(os.path.join(d, c))
            if os.path.isfile(p):
                return p
        nd = os.path.dirname(d)
        if nd == d or not d.startswith(os.path.abspath(REF)):
            break
        d = nd
    raise FileNotFoundError((urdf_path, fname))


def load_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack("<I", data[80:84])[0]
    if 84 + 50 * n == len(data):
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), count=n, offset=84)
        return rec["v"].reshape(n, 3, 3).astype(np.float64)
    # ascii
    verts = [list(map(float, l.split()[1:4])) for l in data.decode("ascii", "ignore").splitlines() if l.strip().startswith("vertex")]
    return np.array(verts, dtype=np.float64).reshape(-1, 3, 3)


def load_obj(path):
    """Triangles [T,3,3]; polygons are fan-triangulated (0,i,i+1) like tinyobjloader."""
    vs, tris = [], []
    with open(path, "r", errors="ignore") as f:
        for line in f:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                idx = []
                for t in line.split()[1:]:
                    i = int(t.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(vs) + i)
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
    vs = np.array(vs, dtype=np.float64)
    return vs[np.array(tris, dtype=np.int64)] if tris else np.zeros((0, 3, 3))


def load_mesh(path):
    return load_stl(path) if path.lower().endswith(".stl") else load_obj(path)


def geom_vertices(urdf_path, elem):
    """Vertices [V,3] (in the geometry's own frame, scaled) of a <visual>/<collision> element, and tris."""
    g = elem.find("geometry")
    origin = elem.find("origin")
    xyz = vec(origin.get("xyz") if origin is not None else None)
    rpy = vec(origin.get("rpy") if origin is not None else None)
    R = rpy_to_mat(rpy)
    mesh, box, sph, cyl = g.find("mesh"), g.find("box"), g.find("sphere"), g.find("cylinder")
    if mesh is not None:
        tris = load_mesh(resolve_mesh(urdf_path, mesh.get("filename")))
        scale = np.array(vec(mesh.get("scale"), 3) if mesh.get("scale") else [1.0, 1.0, 1.0])
        tris = tris * scale
    elif box is not None:
        h = np.array(vec(box.get("size"))) / 2
        c = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
        quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
        tris = np.array([[c[q[0]], c[q[1]], c[q[2]]] for q in quads] + [[c[q[0]], c[q[2]], c[q[3]]] for q in quads])
    elif sph is not None:
        r = atof(sph.get("radius"))
        tris = np.array([[[r, 0, 0], [0, r, 0], [0, 0, r]], [[-r, 0, 0], [0, -r, 0], [0, 0, -r]]], dtype=np.float64)
    elif cyl is not None:
        r, l = atof(cyl.get("radius")), atof(cyl.get("length"))
        tris = np.array([[[r, r, l / 2], [-r, -r, l / 2], [r, -r, -l / 2]], [[-r, r, -l / 2], [r, -r, -l / 2], [-r, -r, l / 2]]], dtype=np.float64)
    else:
        tris = np.zeros((0, 3, 3))
    tris = tris @ R.T + np.array(xyz)
    return tris


def parse_urdf(urdf_path, want_visual=(), want_collision=()):
    root = ET.parse(urdf_path).getroot()
    links, joints = {}, []
    link_order = []
    for l in root.findall("link"):
        name = l.get("name")
        inert = l.find("inertial")
        d = {"name": name, "mass": 0.0, "inertial_xyz": [0.0] * 3, "inertial_rpy": [0.0] * 3, "urdf_inertia": [0.0] * 6}
        if inert is not None:
            o = inert.find("origin")
            if o is not None:
                d["inertial_xyz"] = vec(o.get("xyz"))
                d["inertial_rpy"] = vec(o.get("rpy"))
            m = inert.find("mass")
            d["mass"] = atof(m.get("value")) if m is not None else 0.0
            it = inert.find("inertia")
            if it is not None:
                d["urdf_inertia"] = [atof(it.get(k, "0")) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")]
        d["_elem"] = l
        contact = l.find("contact")
        d["lateral_friction"] = 0.5  # [EXT] bullet default
        d["friction_anchor"] = False
        if contact is not None:
            lf = contact.find("lateral_friction")
            if lf is not None:
                d["lateral_friction"] = atof(lf.get("value"))
            d["friction_anchor"] = contact.find("friction_anchor") is not None
        links[name] = d
        link_order.append(name)
    for j in root.findall("joint"):
        o = j.find("origin")
        ax = j.find("axis")
        joints.append(
            {
                "name": j.get("name"),
                "type": j.get("type"),
                "parent": j.find("parent").get("link"),
                "child": j.find("child").get("link"),
                "xyz": vec(o.get("xyz") if o is not None else None),
                "rpy": vec(o.get("rpy") if o is not None else None),
                "axis": vec(ax.get("xyz")) if ax is not None else [1.0, 0.0, 0.0],
            }
        )
    children = {j["child"] for j in joints}
    roots = [n for n in link_order if n not in children]
    assert len(roots) == 1, roots
    # pre-order DFS, children in joint file order
    order = []

    def walk(lname, pidx):
        for j in joints:
            if j["parent"] == lname:
                idx = len(order)
                order.append((j, pidx))
                walk(j["child"], idx)

    walk(roots[0], -1)

    out_links = []
    for j, pidx in order:
        l = links[j["child"]]
        Rin = rpy_to_mat(l["inertial_rpy"])
        tin = np.array(l["inertial_xyz"])
        # collision AABB in the inertial frame
        cverts = []
        for c in l["_elem"].findall("collision"):
            t = geom_vertices(urdf_path, c).reshape(-1, 3)
            if len(t):
                cverts.append((t - tin) @ Rin)
        if cverts and l["mass"] > 0:
            cv = np.concatenate(cverts)
            h = (cv.max(0) - cv.min(0)) / 2
            ll = 2 * (h + 3 * URDF_COLLISION_MARGIN)
            aabb_inertia = (l["mass"] / 12.0 * np.array([ll[1] ** 2 + ll[2] ** 2, ll[0] ** 2 + ll[2] ** 2, ll[0] ** 2 + ll[1] ** 2])).tolist()
        else:
            aabb_inertia = [0.0, 0.0, 0.0]
        out_links.append(
            {
                "link_name": l["name"],
                "joint_name": j["name"],
                "parent": pidx,
                "joint_type": {"fixed": 0, "revolute": 1, "continuous": 1, "prismatic": 2}[j["type"]],
                "joint_xyz": j["xyz"],
                "joint_rpy": j["rpy"],
                "axis": j["axis"],
                "mass": l["mass"],
                "inertial_xyz": l["inertial_xyz"],
                "inertial_rpy": l["inertial_rpy"],
                "urdf_inertia": l["urdf_inertia"],
                "inertia_diag": aabb_inertia,
                "lateral_friction": l["lateral_friction"],
                "friction_anchor": l["friction_anchor"],
            }
        )
    meshes = {}
    for lname in want_visual:
        ts = [geom_vertices(urdf_path, v) for v in links[lname]["_elem"].findall("visual")]
        meshes["visual:" + lname] = np.concatenate(ts) if ts else np.zeros((0, 3, 3))
    for lname in want_collision:
        ts = [geom_vertices(urdf_path, v) for v in links[lname]["_elem"].findall("collision")]
        meshes["collision:" + lname] = np.concatenate(ts) if ts else np.zeros((0, 3, 3))
    return {"root": roots[0], "links": out_links}, meshes


def box_inertia(mass, size):
    """[EXT] btBoxShape::calculateLocalInertia with the full box size (the URDF margin does not change it)."""
    lx, ly, lz = size
    return [mass / 12.0 * (ly * ly + lz * lz), mass / 12.0 * (lx * lx + lz * lz), mass / 12.0 * (lx * lx + ly * ly)]


def compile_free_object(urdf_path):
    """A floating-base URDF whose joints are all fixed (pole.urdf): per-link mass / inertial frame / inertia
    ([EXT] from the collision box, like loadURDF does without URDF_USE_INERTIA_FROM_FILE), the composite rigid body
    about the composite COM, and the visual triangles in the BASE LINK frame."""
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {j.find("child").get("link") for j in joints}
    base = [n for n in links if n not in children][0]
    frames = {base: (np.eye(3), np.zeros(3))}
    pending = list(joints)
    while pending:
        for j in list(pending):
            par = j.find("parent").get("link")
            if par in frames:
                assert j.get("type") == "fixed", "free objects must be rigid"
                o = j.find("origin")
                Rj = rpy_to_mat(vec(o.get("rpy") if o is not None else None))
                tj = np.array(vec(o.get("xyz") if o is not None else None))
                Rp, tp = frames[par]
                frames[j.find("child").get("link")] = (Rp @ Rj, tp + Rp @ tj)
                pending.remove(j)
    parts, tris = [], []
    for name, l in links.items():
        R, t = frames[name]
        inert = l.find("inertial")
        mass = atof(inert.find("mass").get("value"))
        io = inert.find("origin")
        ixyz = np.array(vec(io.get("xyz") if io is not None else None))
        irpy = vec(io.get("rpy") if io is not None else None)
        col = l.find("collision")
        box = col.find("geometry").find("box")
        assert box is not None, "only box collision shapes are compiled for free objects"
        co = col.find("origin")
        assert np.allclose(vec(co.get("xyz")), ixyz) and np.allclose(vec(co.get("rpy")), irpy), "collision origin must equal the inertial origin"
        Id = box_inertia(mass, vec(box.get("size")))
        Rin = R @ rpy_to_mat(irpy)
        parts.append({"link": name, "mass": mass, "com": (t + R @ ixyz).tolist(), "R": Rin.tolist(), "inertia_diag": Id})
        for v in l.findall("visual"):
            tris.append(geom_vertices(urdf_path, v) @ R.T + t)
    M = sum(p["mass"] for p in parts)
    com = sum(p["mass"] * np.array(p["com"]) for p in parts) / M
    I = np.zeros((3, 3))
    for p in parts:
        Rp = np.array(p["R"]); d = np.array(p["com"]) - com
        I += Rp @ np.diag(p["inertia_diag"]) @ Rp.T + p["mass"] * (d.dot(d) * np.eye(3) - np.outer(d, d))
    assert np.allclose(I, np.diag(np.diag(I)), atol=1e-15), "composite inertia must be diagonal in the base frame"
    base_com = np.array(parts[[p["link"] for p in parts].index(base)]["com"])
    obj = {"base_link": base, "mass": M, "inertia_diag": np.diag(I).tolist(), "base_com": base_com.tolist(),
           "com_off": (com - base_com).tolist(), "parts": parts, "source": os.path.relpath(urdf_path, REF)}
    return obj, np.concatenate(tris)


def convex_hull_vertices(tris):
    from scipy.spatial import ConvexHull

    pts = np.unique(tris.reshape(-1, 3).round(9), axis=0)
    try:
        hull = ConvexHull(pts)
        return pts[hull.vertices]
    except Exception:
        return pts


# (arm, sensor, type) combos that the BASELINE configs + the edge/surface/push/balance defaults need
ROBOTS = [
    ("ur5", "tactip", "standard"),
    ("ur5", "digit", "standard"),
    ("ur5", "digitac", "standard"),
    ("ur5", "tactip", "right_angle"),
    ("ur5", "digit", "right_angle"),
    ("ur5", "digitac", "right_angle"),
    ("ur5", "tactip", "flat"),
    ("mg400", "tactip", "standard"),
    ("mg400", "digit", "standard"),
    ("mg400", "digitac", "standard"),
    ("mg400", "tactip", "right_angle"),
    ("mg400", "tactip", "mini_right_angle"),   # object_push's MG400 + TacTip (object_push_env.py:84-86): the reference's own PPO set-up
    ("mg400", "digit", "right_angle"),
    ("mg400", "digitac", "right_angle"),
    # the `forward` sensor type of surface_follow's vertical surface (noise_mode "vertical_simplex", base_surface_env.py:60-63)
    ("ur5", "tactip", "forward"),
    ("ur5", "digit", "forward"),
    ("ur5", "digitac", "forward"),
    ("mg400", "tactip", "forward"),
    ("mg400", "digit", "forward"),
    ("mg400", "digitac", "forward"),
]

KAT_COMBOS = {("ur5", "tactip", "standard"), ("ur5", "digit", "standard"), ("mg400", "digitac", "right_angle")}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

REST_POSE_FILES = {
    "edge_follow": "rl_envs/exploration/edge_follow/rest_poses.py",
    "surface_follow": "rl_envs/exploration/surface_follow/rest_poses.py",
    "object_push": "rl_envs/nonprehensile_manipulation/object_push/rest_poses.py",
    "object_balance": "rl_envs/nonprehensile_manipulation/object_balance/rest_poses.py",
    "object_roll": "rl_envs/nonprehensile_manipulation/object_roll/rest_poses.py",
}

_PI = float(np.pi)
# camera rig per sensor / type: sensors/tactile_sensor.py:127-148 (fov, focal_dist, near 0.01, far 1.0)
# and :156-182 (cam_pos, cam_rpy in the body link's inertial frame)
SENSOR_CAMERAS = {
    "tactip": {
        "fov": 60.0, "focal_dist": 0.065, "near": 0.01, "far": 1.0,
        "types": {
            "standard": {"cam_pos": [0, 0, 0.03], "cam_rpy": [0, -_PI / 2, _PI]},
            "mini_standard": {"cam_pos": [0, 0, 0.03], "cam_rpy": [0, -_PI / 2, _PI]},
            "flat": {"cam_pos": [0, 0, 0.03], "cam_rpy": [0, -_PI / 2, _PI]},
            "right_angle": {"cam_pos": [0, 0, 0.03], "cam_rpy": [0, -_PI / 2, 140 * _PI / 180]},
            "forward": {"cam_pos": [0, 0, 0.03], "cam_rpy": [0, -_PI / 2, 140 * _PI / 180]},
            "mini_right_angle": {"cam_pos": [0, 0, 0.001], "cam_rpy": [0, -_PI / 2, 140 * _PI / 180]},
        },
    },
    "digit": {
        "fov": 40.0, "focal_dist": 0.0015, "near": 0.01, "far": 1.0,
        "types": {
            "standard": {"cam_pos": [-0.00095, 0.0139, 0.020], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
            "right_angle": {"cam_pos": [-0.00095, 0.0139, 0.005], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
            "forward": {"cam_pos": [-0.00095, 0.0139, 0.005], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
        },
    },
    "digitac": {
        "fov": 40.0, "focal_dist": 0.0015, "near": 0.01, "far": 1.0,
        "types": {
            "standard": {"cam_pos": [-0.00095, 0.0139, 0.020], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
            "right_angle": {"cam_pos": [-0.00095, 0.0139, 0.005], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
            "forward": {"cam_pos": [-0.00095, 0.0139, 0.005], "cam_rpy": [_PI, -_PI / 2, _PI / 2]},
        },
    },
}

OBJECTS = {
    "pole": "rl_env_assets/nonprehensile_manipulation/object_balance/pole/pole.urdf",
    "cube": "rl_env_assets/nonprehensile_manipulation/object_push/cube/cube.urdf",
}

STIMULI = {
    "long_edge": "rl_env_assets/exploration/edge_follow/edge_stimuli/long_edge_flat/long_edge.urdf",
    "short_edge": "rl_env_assets/exploration/edge_follow/edge_stimuli/long_edge_flat/short_edge.urdf",
}


def main():
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "refimg"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "stimuli"), exist_ok=True)

    for arm, sensor, typ in ROBOTS:
        urdf = os.path.join(ASSETS, "robot_assets", arm, sensor, "%s_with_%s_%s.urdf" % (arm, typ, sensor))
        if not os.path.isfile(urdf):
            print("skip (no urdf)", urdf)
            continue
        body, tip = sensor + "_body_link", sensor + "_tip_link"
        wv = [body, tip] + (["tactip_adapter_link"] if (sensor == "tactip" and typ in ("right_angle", "forward")) else [])
        model, meshes = parse_urdf(urdf, want_visual=wv, want_collision=[tip])
        model.update({"arm": arm, "sensor": sensor, "type": typ, "source": os.path.relpath(urdf, REF)})
        # the tip core's collision primitive when it is not a mesh (the flat TacTip's core is a cylinder)
        tip_elem = [l for l in ET.parse(urdf).getroot().findall("link") if l.get("name") == tip][0]
        col = tip_elem.find("collision")
        cyl = col.find("geometry").find("cylinder") if col is not None else None
        if cyl is not None:
            co = col.find("origin")
            model["tip_collision"] = {"type": "cylinder", "length": atof(cyl.get("length")), "radius": atof(cyl.get("radius")),
                                      "xyz": vec(co.get("xyz") if co is not None else None), "rpy": vec(co.get("rpy") if co is not None else None)}
        name = "%s_%s_%s" % (arm, typ, sensor)
        with open(os.path.join(OUT, "models", name + ".json"), "w") as f:
            json.dump(model, f, indent=1)
        # the tip collision hull (needed by the contact envs) and the sensor's own visual meshes
        # (needed only by the fixture-regeneration KAT in tests)
        core = convex_hull_vertices(meshes["collision:" + tip]) if len(meshes["collision:" + tip]) else np.zeros((0, 3))
        np.savez_compressed(os.path.join(OUT, "models", name + "_meshes.npz"), tip_core_hull=core.astype(np.float64))
        if (arm, sensor, typ) in KAT_COMBOS:
            # the sensor's own visual meshes, in their link frames: only the fixture-regeneration
            # known-answer test needs them, so they live with the tests
            os.makedirs(GOLDEN, exist_ok=True)
            np.savez_compressed(
                os.path.join(GOLDEN, "sensor_visual_%s.npz" % name),
                **{k.replace("visual:", ""): v.astype(np.float32) for k, v in meshes.items() if k.startswith("visual:")},
            )
        print("model", name, len(model["links"]), "links; core hull", core.shape)

    # reference images (the only golden data the reference ships, tactile_sensor.py:63-80)
    for sensor in ("tactip", "digit", "digitac"):
        base = os.path.join(ASSETS, "robot_assets", sensor, "reference_images")
        for typ in sorted(os.listdir(base)):
            if not os.path.isdir(os.path.join(base, typ)):
                continue
            for sz in sorted(os.listdir(os.path.join(base, typ))):
                w, h = sz.split("x")
                if w != h or int(w) not in (64, 128, 256):
                    continue
                d = os.path.join(base, typ, sz)
                np.savez_compressed(
                    os.path.join(OUT, "refimg", "%s_%s_%s.npz" % (sensor, typ, w)),
                    nodef_dep=np.load(os.path.join(d, "nodef_dep.npy")),
                    nodef_gray=np.load(os.path.join(d, "nodef_gray.npy")),
                    border_mask=np.load(os.path.join(d, "border_mask.npy")),
                )
    # rest poses (plain numeric tables; the modules only import numpy)
    import importlib.util

    rest = {}
    for env, rel in REST_POSE_FILES.items():
        spec = importlib.util.spec_from_file_location("rp_" + env, os.path.join(REF, "tactile_gym", rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        def conv(x):
            return {k: conv(v) for k, v in x.items()} if isinstance(x, dict) else np.asarray(x).tolist()

        rest[env] = conv(mod.rest_poses_dict)
    with open(os.path.join(OUT, "rest_poses.json"), "w") as f:
        json.dump(rest, f, indent=1)
    with open(os.path.join(OUT, "sensors.json"), "w") as f:
        json.dump(SENSOR_CAMERAS, f, indent=1)

    # free objects (floating-base bodies whose links are all fixed to the base): composite inertia + visual triangles
    os.makedirs(os.path.join(OUT, "objects"), exist_ok=True)
    for name, rel in OBJECTS.items():
        urdf = os.path.join(ASSETS, rel)
        obj, tris = compile_free_object(urdf)
        with open(os.path.join(OUT, "objects", name + ".json"), "w") as f:
            json.dump(obj, f, indent=1)
        np.savez_compressed(os.path.join(OUT, "stimuli", name + ".npz"), tris=tris.astype(np.float64))
        print("object", name, "mass", obj["mass"], "tris", tris.shape)

    for name, rel in STIMULI.items():
        urdf = os.path.join(ASSETS, rel)
        root = ET.parse(urdf).getroot()
        l = root.find("link")
        tris = np.concatenate([geom_vertices(urdf, v) for v in l.findall("visual")])
        np.savez_compressed(os.path.join(OUT, "stimuli", name + ".npz"), tris=tris.astype(np.float64))
        print("stimulus", name, tris.shape)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Golden vectors computed by the REFERENCE'S OWN SOURCE for the pure-numpy pieces of the hot path
(tests/golden/reference_numpy.npz; replayed against the oracle by tests/test_oracle_reference_golden.py).

The reference cannot be imported here (pybullet / gym / opensimplex are not installed), but a good part of the path around its
two native calls is plain numpy: action encoding and scaling, the work-frame transforms, TCP limit handling, the edge and
surface reward geometry, the surface index lookup.  This script reads the reference's .py files, compiles the class bodies as
they stand (ast -> exec; base classes replaced by the compiled ones, `__init__` never run), binds the attributes those methods
read, and records inputs and outputs.  The only stand-in is `_pb`: seven pybullet maths helpers (quaternion / euler / transform
algebra, [EXT] pybullet conventions: xyzw quaternions, getQuaternionFromEuler = fixed-axis roll-pitch-yaw) written in numpy.

Run in the build container only (it needs /root/reference); the GPU box never does.
usage: make_reference_golden.py [reference root] [output.npz]"""
import ast
import os
import sys
import types

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_numpy.npz")
ENVS = os.path.join(REF, "tactile_gym", "rl_envs")


# ------------------------------------------------------------------ pybullet maths stand-ins
class PB:
    @staticmethod
    def getQuaternionFromEuler(rpy):
        r, p, y = [float(v) for v in rpy]
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
        return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)

    @staticmethod
    def getMatrixFromQuaternion(q):
        x, y, z, w = [float(v) for v in q]
        return (1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y))

    @staticmethod
    def getEulerFromQuaternion(q):
        x, y, z, w = [float(v) for v in q]
        sarg = -2 * (x * z - w * y)
        if sarg <= -0.99999:
            return (0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y))
        if sarg >= 0.99999:
            return (0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y))
        sq = [x * x, y * y, z * z, w * w]
        return (np.arctan2(2 * (y * z + w * x), sq[3] - sq[0] - sq[1] + sq[2]), np.arcsin(sarg),
                np.arctan2(2 * (x * y + w * z), sq[3] + sq[0] - sq[1] - sq[2]))

    @staticmethod
    def _qmul(a, b):
        ax, ay, az, aw = a; bx, by, bz, bw = b
        return (aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz)

    @classmethod
    def multiplyTransforms(cls, pa, qa, pb, qb):
        R = np.array(cls.getMatrixFromQuaternion(qa)).reshape(3, 3)
        return tuple(np.asarray(pa, float) + R @ np.asarray(pb, float)), cls._qmul(tuple(float(v) for v in qa), tuple(float(v) for v in qb))

    @classmethod
    def invertTransform(cls, p, q):
        qi = (-float(q[0]), -float(q[1]), -float(q[2]), float(q[3]))
        R = np.array(cls.getMatrixFromQuaternion(qi)).reshape(3, 3)
        return tuple(-(R @ np.asarray(p, float))), qi

    def __getattr__(self, name):   # resetBasePositionAndOrientation, addUserDebugLine ...: scene bookkeeping, no arithmetic
        return lambda *a, **k: None


class _Box:
    def __init__(self, **kw):
        self.__dict__.update(kw)


GYM = types.SimpleNamespace(spaces=types.SimpleNamespace(Box=_Box, Dict=dict), Env=object)


def ref_class(path, name, bases=(), extra=None):
    """the reference's class `name` from `path`, compiled from its source text: methods only, given bases"""
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    ns = {"np": np, "sys": sys, "gym": GYM, "__name__": "reference_source"}
    ns.update(extra or {})
    for k, b in enumerate(bases):
        ns["_Base%d" % k] = b
    cls.bases = [ast.Name(id="_Base%d" % k, ctx=ast.Load()) for k in range(len(bases))]
    cls.keywords = []
    cls.body = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name != "__del__"]   # (its close() wants a live client)
    consts = [n for n in tree.body if isinstance(n, ast.Assign)]     # module-level literals the signatures default to (env_modes_default)
    mod = ast.Module(body=consts + [cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


def bare(cls, **attrs):
    o = object.__new__(cls)
    o.__dict__.update(attrs)
    return o


def main():
    rng = np.random.RandomState(20261017)
    out = {}
    BaseTactileEnv = ref_class(os.path.join(ENVS, "base_tactile_env.py"), "BaseTactileEnv")
    BaseRobotArm = ref_class(os.path.join(REF, "tactile_gym", "robots", "arms", "base_robot_arm.py"), "BaseRobotArm")
    edge_py = os.path.join(ENVS, "exploration", "edge_follow", "edge_follow_env.py")
    surf = os.path.join(ENVS, "exploration", "surface_follow")
    EdgeFollowEnv = ref_class(edge_py, "EdgeFollowEnv", (BaseTactileEnv,))
    BaseSurfaceEnv = ref_class(os.path.join(surf, "base_surface_env.py"), "BaseSurfaceEnv", (BaseTactileEnv,))
    SurfAuto = ref_class(os.path.join(surf, "surface_follow_auto", "surface_follow_auto_env.py"), "SurfaceFollowAutoEnv", (BaseSurfaceEnv,))
    SurfGoal = ref_class(os.path.join(surf, "surface_follow_goal", "surface_follow_goal_env.py"), "SurfaceFollowGoalEnv", (BaseSurfaceEnv,))
    SurfVert = ref_class(os.path.join(surf, "surface_follow_vert", "surface_follow_vert_env.py"), "SurfaceFollowVertEnv", (BaseSurfaceEnv,))
    obj = os.path.join(ENVS, "nonprehensile_manipulation")
    BaseObjectEnv = ref_class(os.path.join(obj, "base_object_env.py"), "BaseObjectEnv", (BaseTactileEnv,))
    Balance = ref_class(os.path.join(obj, "object_balance", "object_balance_env.py"), "ObjectBalanceEnv", (BaseObjectEnv,))
    Roll = ref_class(os.path.join(obj, "object_roll", "object_roll_env.py"), "ObjectRollEnv", (BaseObjectEnv,))

    # ---- A. encode_actions + scale_actions (R2): every movement mode, both control modes where the env defines them
    def actions_case(key, cls, act_dim, **attrs):
        env = bare(cls, **attrs)
        env.get_act_dim = lambda: act_dim
        env.setup_action_space()
        acts = rng.uniform(-0.3, 0.3, (12, act_dim))          # beyond +-0.25 on purpose: scale_actions clips
        res = np.array([env.scale_actions(env.encode_actions(a)) for a in acts])
        out["act_%s_in" % key], out["act_%s_out" % key] = acts, res

    for mode, nd in (("xy", 2), ("xyz", 3), ("xyRz", 3), ("xyzRz", 4)):
        for cm in ("TCP_velocity_control", "TCP_position_control"):
            actions_case("edge_%s_%s" % (mode, cm[4:7]), EdgeFollowEnv, nd, movement_mode=mode, control_mode=cm)
    dirs = np.array([np.cos(0.7), np.sin(0.7), 0.0])
    out["surface_dirs"] = dirs
    for sensor in ("tactip", "digitac", "digit"):
        for mode, nd in (("yz", 1), ("xyz", 1), ("yzRx", 2), ("xyzRxRy", 3)):
            for cm in ("TCP_velocity_control", "TCP_position_control"):
                actions_case("surfauto_%s_%s_%s" % (sensor, mode, cm[4:7]), SurfAuto, nd, movement_mode=mode, control_mode=cm, noise_mode="simplex",
                             t_s_name=sensor, workframe_directions=list(dirs))
    for mode, nd in (("yz", 2), ("xyz", 3), ("yzRx", 3), ("xyzRxRy", 5)):
        actions_case("surfgoal_%s" % mode, SurfGoal, nd, movement_mode=mode, control_mode="TCP_velocity_control", noise_mode="simplex", t_s_name="tactip")
    for sensor in ("tactip", "digitac", "digit"):
        actions_case("surfvert_%s" % sensor, SurfVert, 2, movement_mode="xRz", control_mode="TCP_velocity_control", noise_mode="simplex",
                     t_s_name=sensor, workframe_directions=[0, -1, 0])
    for mode, nd in (("xy", 2), ("xyz", 3), ("RxRy", 2), ("xyRxRy", 4)):
        actions_case("balance_%s" % mode, Balance, nd, movement_mode=mode, control_mode="TCP_velocity_control")
    actions_case("roll_xy", Roll, 2, movement_mode="xy", control_mode="TCP_velocity_control")

    # ---- B. work-frame transforms and TCP limits (R3): BaseRobotArm with the edge / surface work frame and the balance one
    for key, wpos, wrpy in (("flipped", [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2]), ("upright", [0.55, 0.0, 0.35], [0.0, 0.0, 0.0])):
        arm = bare(BaseRobotArm, _pb=PB())
        arm.set_workframe(wpos, wrpy)
        lims = np.array([[-0.1, 0.1], [-0.05, 0.12], [-0.02, 0.03], [-0.3, 0.3], [-0.2, 0.4], [-1.0, 1.0]])
        arm.set_TCP_lims(lims)
        pos, rpy = rng.uniform(-0.3, 0.3, (10, 3)) + np.array(wpos), rng.uniform(-1.2, 1.2, (10, 3))
        vec = rng.uniform(-1, 1, (10, 6))
        w2k = [arm.worldframe_to_workframe(p, r) for p, r in zip(pos, rpy)]
        k2w = [arm.workframe_to_worldframe(p - np.array(wpos), r) for p, r in zip(pos, rpy)]
        out["frame_%s_wpos" % key], out["frame_%s_wrpy" % key], out["frame_%s_lims" % key] = np.array(wpos), np.array(wrpy), lims
        out["frame_%s_pos" % key], out["frame_%s_rpy" % key], out["frame_%s_vec" % key] = pos, rpy, vec
        out["frame_%s_w2k_pos" % key], out["frame_%s_w2k_rpy" % key] = np.array([a for a, _ in w2k]), np.array([b for _, b in w2k])
        out["frame_%s_k2w_pos" % key], out["frame_%s_k2w_rpy" % key] = np.array([a for a, _ in k2w]), np.array([b for _, b in k2w])
        out["frame_%s_vec_w2k" % key] = np.array([arm.worldvec_to_workvec(v[:3]) for v in vec])
        out["frame_%s_vec_k2w" % key] = np.array([arm.workvec_to_worldvec(v[:3]) for v in vec])
        out["frame_%s_vel_w2k" % key] = np.array([np.concatenate(arm.worldvel_to_workvel(v[:3], v[3:])) for v in vec])
        cur = rng.uniform(-0.15, 0.15, (10, 6))
        capped = []
        for c, v in zip(cur, vec):
            arm.get_current_TCP_pos_vel_workframe = lambda c=c: (c[:3], c[3:], None, None, None)
            capped.append(arm.check_TCP_vel_lims(v * 0.01))
        out["frame_%s_cur" % key], out["frame_%s_vel_capped" % key] = cur, np.array(capped)
        out["frame_%s_pos_clipped" % key] = np.array([np.concatenate(arm.check_TCP_pos_lims(c[:3], c[3:])) for c in cur])

    # ---- C. edge_follow reward / termination geometry (R7): update_edge as written, then get_step_data's pieces
    arm = bare(BaseRobotArm, _pb=PB())
    arm.set_workframe([0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2])
    angs, tcps, steps = rng.uniform(-np.pi, np.pi, 10), rng.uniform(-0.2, 0.2, (10, 3)) + np.array([0.65, 0.0, 0.035]), rng.randint(0, 260, 10)
    tcps[3] = None   # filled below: a TCP inside the termination radius
    rows = []
    for k in range(10):
        env = bare(EdgeFollowEnv, _pb=PB(), robot=types.SimpleNamespace(arm=arm), edge_pos=[0.65, 0.0, 0.0], edge_len=0.175, edge_height=0.035,
                   edge_stim_id=0, goal_indicator=1, termination_dist=0.01, _max_steps=250, _env_step_counter=int(steps[k]),
                   np_random=types.SimpleNamespace(uniform=lambda lo, hi, k=k: angs[k]))
        env.update_edge()
        if k == 3:
            tcps[3] = np.array(env.goal_pos_worldframe) + np.array([0.004, -0.003, 0.01])
        env.cur_tcp_pos_worldframe = tcps[k]
        rows.append([env.xy_dist_to_goal(), env.dist_to_center_edge(), env.dense_reward(), env.sparse_reward(), float(env.termination()),
                     *env.goal_pos_workframe, *env.goal_pos_worldframe])
    out["edge_ang"], out["edge_tcp"], out["edge_steps"], out["edge_rows"] = angs, tcps, steps, np.array(rows, dtype=np.float64)

    # ---- D. surface_follow: setup_surface bins, index lookup, distances and the three envs' rewards on a given heightfield
    h = rng.uniform(-0.02, 0.02, (64, 64))
    for k in range(3):     # smooth it a little so that the normals are not degenerate
        h = 0.25 * (np.roll(h, 1, 0) + np.roll(h, -1, 0) + np.roll(h, 1, 1) + np.roll(h, -1, 1))
    sv = bare(SurfAuto, _pb=PB(), noise_mode="simplex", movement_mode="xyzRxRy", well_designed_pos=[0.65, 0.0, 0.0], embed_dist=0.0025,
              termination_dist=0.01, _max_steps=200, _env_step_counter=10, reward_mode="dense")
    sv.setup_surface()
    out["surf_x_bins"], out["surf_y_bins"], out["surf_h"] = sv.x_bins, sv.y_bins, h
    pts = rng.uniform(-0.21, 0.21, (40, 2)) + np.array([0.65, 0.0])
    pts[:4] = [[sv.x_bins[0], sv.y_bins[0]], [sv.x_bins[-1], sv.y_bins[-1]], [sv.x_bins[5], sv.y_bins[7]], [0.65, 0.0]]
    out["surf_pts"], out["surf_idx"] = pts, np.array([sv.xy_to_surface_idx(p[0], p[1]) for p in pts])
    # surface_array / surface_normals exactly as update_surface builds them (:480-499), with this heightfield
    X, Y = np.meshgrid(sv.x_bins, sv.y_bins)
    surface_array = np.dstack((X, Y, h + sv.surface_pos[2]))
    gy, gx = np.gradient(h, sv.heightfield_grid_scale)
    nrm = np.dstack((-gx, -gy, np.ones_like(h)))
    nn = np.linalg.norm(nrm, axis=2)
    for c in range(3):
        nrm[:, :, c] /= nn
    tcp_pos = np.column_stack([rng.uniform(0.55, 0.75, 12), rng.uniform(-0.1, 0.1, 12), rng.uniform(0.0, 0.05, 12)])
    tcp_rpy = np.column_stack([np.pi + rng.uniform(-0.5, 0.5, 12), rng.uniform(-0.5, 0.5, 12), rng.uniform(-0.3, 0.3, 12)])
    goal = np.array([0.70, 0.05, 0.03])
    rows = []
    for p, r in zip(tcp_pos, tcp_rpy):
        vals = []
        for cls, mode in ((SurfAuto, "xyzRxRy"), (SurfAuto, "xyz"), (SurfGoal, "xyzRxRy"), (SurfVert, "xRz")):
            e = bare(cls, _pb=PB(), noise_mode="simplex", movement_mode=mode, embed_dist=0.0025, surface_array=surface_array, surface_normals=nrm,
                     x_bins=sv.x_bins, y_bins=sv.y_bins, num_heightfield_rows=64, num_heightfield_cols=64, goal_pos_worldframe=goal,
                     cur_tcp_pos_worldframe=p, cur_tcp_orn_worldframe=PB.getQuaternionFromEuler(r), termination_dist=0.01, _max_steps=200,
                     _env_step_counter=10)
            e.tip_i, e.tip_j = e.xy_to_surface_idx(p[0], p[1])
            vals += [e.z_dist_to_surface(), e.cos_dist_to_surface_normal(), e.dense_reward()]
        rows.append(vals + [e.xyz_dist_to_goal(), e.xy_dist_to_goal(), float(e.termination())])
    out["surf_tcp_pos"], out["surf_tcp_rpy"], out["surf_goal"], out["surf_rows"] = tcp_pos, tcp_rpy, goal, np.array(rows)

    np.savez_compressed(OUT, **out)
    print("wrote %s: %d arrays" % (OUT, len(out)))


if __name__ == "__main__":
    main()

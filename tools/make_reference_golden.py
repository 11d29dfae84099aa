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

    def __init__(self):
        self.bodies = {}

    def resetBasePositionAndOrientation(self, uid, pos, orn):
        self.bodies[uid] = (tuple(float(v) for v in pos), tuple(float(v) for v in orn))

    def getBasePositionAndOrientation(self, uid):
        return self.bodies[uid]

    def __getattr__(self, name):   # changeVisualShape, addUserDebugLine ...: scene bookkeeping, no arithmetic
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
    noise = lambda x, y: 0.6 * np.sin(2.3 * x + 0.4) * np.cos(0.7 * y) + 0.2 * np.sin(5.1 * x)   # stand-in for OpenSimplex.noise2
    FakeSimplex = type("OpenSimplex", (), {"__init__": lambda self, seed: None, "noise2": lambda self, x, y: noise(x, y)})
    Push = ref_class(os.path.join(obj, "object_push", "object_push_env.py"), "ObjectPushEnv", (BaseObjectEnv,), extra={"OpenSimplex": FakeSimplex, "os": os})

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

    # ---- E. object_push (R2, R7, R10): TCP-frame / work-frame action encodings, the trajectory of goals around a stand-in noise
    # function (the replay injects the same one into the oracle), rewards, goal advancing, extended feature, oracle observation
    def arm_with_state(wpos, wrpy, tcp_pos, push_rpy, lin, ang):
        arm = bare(BaseRobotArm, _pb=PB())
        arm.set_workframe(wpos, wrpy)
        q = PB.getQuaternionFromEuler(push_rpy)
        arm.get_current_TCP_pos_vel_worldframe = lambda: (np.array(tcp_pos), np.array(PB.getEulerFromQuaternion(q)), np.array(q), np.array(lin), np.array(ang))
        return arm

    wpos, wrpy = np.array([0.55, -0.20, 0.04]), np.array([-np.pi, 0.0, np.pi / 2])      # ur5: well_designed_pos, obj_height / 2 (:96-101)
    out["push_wpos"], out["push_wrpy"] = wpos, wrpy
    push_rpy = np.array([-np.pi, 0.0, np.pi / 2 + 0.31])
    out["push_tcp_rpy_for_actions"] = push_rpy
    for mode, nd in (("y", 1), ("yRz", 2), ("xyRz", 3), ("TyRz", 2), ("TxTyRz", 3)):
        arm = arm_with_state(wpos, wrpy, wpos, push_rpy, np.zeros(3), np.zeros(3))
        actions_case("push_%s" % mode, Push, nd, movement_mode=mode, control_mode="TCP_velocity_control", _pb=PB(),
                     robot=types.SimpleNamespace(arm=arm), cur_tcp_orn_worldframe=PB.getQuaternionFromEuler(push_rpy))
    for traj_type, third in (("simplex", 4242.0), ("straight", 0.23)):
        pb = PB()
        arm = arm_with_state(wpos, wrpy, wpos + np.array([0.01, 0.02, 0.0]), push_rpy, [0.004, -0.003, 0.001], [0.01, 0.02, -0.2])
        env = bare(Push, _pb=pb, robot=types.SimpleNamespace(arm=arm), traj_type=traj_type, traj_n_points=10, traj_spacing=0.025, traj_max_perturb=0.1,
                   traj_ids=list(range(100, 110)), obj_width=0.08, termination_pos_dist=0.025, _max_steps=1000, _env_step_counter=5, reward_mode="dense",
                   obj_id=7, np_random=types.SimpleNamespace(randint=lambda hi: int(third), uniform=lambda lo, hi: third))
        env.make_goal()
        out["push_%s_traj_pos_work" % traj_type], out["push_%s_traj_rpy_work" % traj_type] = env.traj_pos_workframe.copy(), env.traj_rpy_workframe.copy()
        out["push_%s_traj_pos_world" % traj_type] = np.array([pb.bodies[i][0] for i in env.traj_ids])
        out["push_%s_traj_orn_world" % traj_type] = np.array([pb.bodies[i][1] for i in env.traj_ids])
        # the cube walks along the trajectory: rewards before the goal advances, goal index after
        rows, objs = [], []
        for k in range(24):
            g = min(env.targ_traj_list_id, 9)
            near = k % 3 != 1
            op = np.array(pb.bodies[env.traj_ids[g]][0]) + (np.array([0.004, -0.006, 0.0]) if near else np.array([0.05, 0.03, 0.0]))
            oq = PB.getQuaternionFromEuler([-np.pi, 0.0, np.pi / 2 + 0.1 * k - 0.4])
            pb.bodies[7] = (tuple(op), oq)
            objs.append(np.concatenate([op, oq]))
            env.reward_mode = "dense"
            (env.cur_tcp_pos_worldframe, env.cur_push_rpy_worldframe, env.cur_tcp_orn_worldframe, _, _) = arm.get_current_TCP_pos_vel_worldframe()
            env.cur_obj_pos_worldframe, env.cur_obj_orn_worldframe = env.get_obj_pos_worldframe()
            dense, sparse = env.dense_reward(), env.sparse_reward()
            rew, done = env.get_step_data()
            feat = env.get_extended_feature_array()
            rows.append([dense, sparse, rew, float(done), env.targ_traj_list_id, *feat])
            if done:
                break
        out["push_%s_obj" % traj_type], out["push_%s_rows" % traj_type] = np.array(objs), np.array(rows, dtype=np.float64)
        if traj_type == "simplex":
            pb.bodies[7] = (tuple(wpos + np.array([0.05, 0.01, 0.0])), PB.getQuaternionFromEuler([-np.pi, 0.0, np.pi / 2 + 0.2]))
            pb.getBaseVelocity = lambda uid: ((0.01, -0.02, 0.003), (0.1, 0.2, -0.3))
            out["push_oracle_obj"] = np.concatenate([pb.bodies[7][0], pb.bodies[7][1], [0.01, -0.02, 0.003], [0.1, 0.2, -0.3]])
            out["push_oracle_tcp"] = np.concatenate([wpos + np.array([0.01, 0.02, 0.0]), push_rpy, [0.004, -0.003, 0.001], [0.01, 0.02, -0.2]])
            out["push_oracle_goal_index"] = np.array([env.targ_traj_list_id])
            out["push_oracle_obs"] = np.array(env.get_oracle_obs(), dtype=np.float64)

    # ---- F. object_balance: check_obj_fall / rewards (R7) and its oracle observation (R10)
    bw, brpy = np.array([0.55, 0.0, 0.35]), np.zeros(3)
    init_pos, init_rpy = bw + np.array([0.0, 0.0, 0.00125 - 0.0035]), np.array([0.0, 0.0, -np.pi / 2])
    rows, objs = [], []
    for k in range(12):
        pb = PB()
        rpy = init_rpy + np.array([0.25, -0.2, 0.4]) * rng.uniform(-3.5, 3.5, 3)
        pos = init_pos + rng.uniform(-0.09, 0.09, 3) * (1.0 if k % 2 else 0.3)
        pb.bodies[3] = (tuple(pos), PB.getQuaternionFromEuler(rpy))
        e = bare(Balance, _pb=pb, obj_id=3, init_obj_rpy=init_rpy, init_obj_pos=init_pos, termination_dist_deg=35, termination_dist_pos=0.1,
                 _max_steps=250, _env_step_counter=int(rng.randint(0, 260)))
        objs.append(np.concatenate([pos, pb.bodies[3][1]]))
        rows.append([float(e.check_obj_fall()), float(e.termination()), e.sparse_reward(), e.dense_reward(), e._env_step_counter])
    out["balance_obj"], out["balance_rows"], out["balance_init_pos"] = np.array(objs), np.array(rows), init_pos
    pb = PB()
    pb.bodies[3] = (tuple(init_pos + np.array([0.01, -0.02, 0.003])), PB.getQuaternionFromEuler(init_rpy + np.array([0.1, -0.05, 0.2])))
    pb.getBaseVelocity = lambda uid: ((0.01, -0.02, 0.003), (0.1, 0.2, -0.3))
    arm = arm_with_state(bw, brpy, bw + np.array([0.002, 0.001, -0.0003]), [0.05, -0.02, 0.01], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2])
    arm._pb = pb
    e = bare(Balance, _pb=pb, obj_id=3, robot=types.SimpleNamespace(arm=arm))
    out["balance_oracle_obj"] = np.concatenate([pb.bodies[3][0], pb.bodies[3][1], [0.01, -0.02, 0.003], [0.1, 0.2, -0.3]])
    out["balance_oracle_tcp"] = np.concatenate([bw + np.array([0.002, 0.001, -0.0003]), [0.05, -0.02, 0.01], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2]])
    out["balance_oracle_obs"] = np.array(e.get_oracle_obs(), dtype=np.float64)

    # ---- G. object_roll: the goal fixed in the TCP frame, rewards / termination, feature and oracle observation
    radius, embed = 0.0025 * 1.4, 0.0025
    rw_, rrpy = np.array([0.65, 0.0, 2 * radius - embed]), np.array([-np.pi, 0.0, np.pi / 2])
    rows = []
    tcp_p, tcp_r = rw_ + np.array([0.003, -0.002, 0.0002]), np.array([-np.pi, 0.0, np.pi / 2])
    for k in range(8):
        pb = PB()
        arm = arm_with_state(rw_, rrpy, tcp_p, tcp_r, [0.004, -0.003, 0.0], [0.0, 0.0, 0.0])
        goal_tcp = np.array([0.008 * np.cos(0.9 * k), 0.008 * np.sin(0.9 * k), 0.0])
        e = bare(Roll, _pb=pb, obj_id=5, robot=types.SimpleNamespace(arm=arm), goal_pos_tcp=goal_tcp, goal_rpy_tcp=[0.0, 0.0, 0.0],
                 goal_orn_tcp=PB.getQuaternionFromEuler([0.0, 0.0, 0.0]), visualise_goal=False, termination_pos_dist=0.001, _max_steps=250,
                 _env_step_counter=3 if k < 7 else 250, reward_mode="dense", scaled_obj_radius=radius)
        e.update_goal()
        obj_p = np.array(e.goal_pos_worldframe) + (np.array([0.0004, -0.0003, 0.0]) if k % 2 == 0 else np.array([0.004, 0.003, 0.0]))
        obj_p[2] = radius
        pb.bodies[5] = (tuple(obj_p), (0.0, 0.0, 0.0, 1.0))
        rew_d, done = e.get_step_data()
        e.reward_mode = "sparse"
        rew_s, _ = e.get_step_data()
        rows.append([*goal_tcp, *obj_p, *e.goal_pos_worldframe, rew_d, rew_s, float(done), e._env_step_counter, *e.get_extended_feature_array()])
    out["roll_rows"], out["roll_tcp"], out["roll_wpos"], out["roll_radius"] = np.array(rows), np.concatenate([tcp_p, tcp_r]), rw_, np.array([radius, embed])
    pb.getBaseVelocity = lambda uid: ((0.01, -0.02, 0.0), (0.1, 0.2, -0.3))
    arm._pb = pb
    out["roll_oracle_obj"] = np.concatenate([pb.bodies[5][0], pb.bodies[5][1], [0.01, -0.02, 0.0], [0.1, 0.2, -0.3]])
    out["roll_oracle_obs"] = np.array(e.get_oracle_obs(), dtype=np.float64)

    # ---- H. edge / surface oracle observations (R10) from a given world TCP state
    arm = arm_with_state([0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], [0.66, 0.01, 0.032], [-np.pi + 0.02, 0.03, np.pi / 2 - 0.4], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2])
    e = bare(EdgeFollowEnv, _pb=PB(), robot=types.SimpleNamespace(arm=arm), edge_pos=[0.65, 0.0, 0.0], edge_len=0.175, edge_height=0.035, edge_stim_id=0,
             goal_indicator=1, np_random=types.SimpleNamespace(uniform=lambda lo, hi: 1.1))
    e.update_edge()
    out["edge_oracle_tcp"] = np.concatenate([[0.66, 0.01, 0.032], [-np.pi + 0.02, 0.03, np.pi / 2 - 0.4], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2]])
    out["edge_oracle_obs"] = np.array(e.get_oracle_obs(), dtype=np.float64)
    arm = arm_with_state([0.65, 0.0, 0.025], [-np.pi, 0.0, np.pi / 2], out["surf_tcp_pos"][0], out["surf_tcp_rpy"][0], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2])
    e = bare(SurfAuto, _pb=PB(), robot=types.SimpleNamespace(arm=arm), surface_array=surface_array, surface_normals=nrm, x_bins=sv.x_bins, y_bins=sv.y_bins,
             num_heightfield_rows=64, num_heightfield_cols=64, goal_pos_workframe=arm.worldframe_to_workframe(goal, [0, 0, 0])[0])
    e.tip_i, e.tip_j = e.xy_to_surface_idx(out["surf_tcp_pos"][0][0], out["surf_tcp_pos"][0][1])
    out["surf_oracle_tcp"] = np.concatenate([out["surf_tcp_pos"][0], out["surf_tcp_rpy"][0], [0.004, -0.003, 0.001], [0.01, 0.02, -0.2]])
    out["surf_oracle_obs"] = np.array(e.get_oracle_obs(), dtype=np.float64)

    # ---- I. TactileSensor (R8, R9): the camera rig relative to the sensor body (update_cam_frame + get_imgs's vectors, with the
    # body at the identity pose) and t_s_camera on a synthetic depth image: the reference's own fixture images, load_reference_images
    # as written, a depth = nodef_dep with dents / noise that exercise the 1e-4 dead band, the 0.05 clip and the uint8 truncation
    assets = os.path.join(REF, "tactile_gym", "assets")
    TactileSensor = ref_class(os.path.join(REF, "tactile_gym", "sensors", "tactile_sensor.py"), "TactileSensor",
                              extra={"os": os, "add_assets_path": lambda p: os.path.join(assets, p)})
    for name, typ, S, border_off in (("tactip", "standard", 64, False), ("tactip", "standard", 128, False), ("tactip", "flat", 128, False),
                                     ("digit", "standard", 128, False), ("digitac", "right_angle", 128, True), ("tactip", "standard", 256, False)):
        pb = PB()
        captured = {}
        pb.getLinkState = lambda *a, **k: ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), None, None, None, None)
        pb.computeProjectionMatrixFOV = lambda fov, aspect, near, far: (fov, aspect, near, far)
        pb.computeViewMatrix = lambda eye, target, up: captured.update(eye=np.array(eye), target=np.array(target), up=np.array(up)) or "view"
        pb.ER_SEGMENTATION_MASK_OBJECT_AND_LINKINDEX, pb.ER_BULLET_HARDWARE_OPENGL = 1, 2
        ts = bare(TactileSensor, _pb=pb, robot_id=4, tactile_link_ids={"body": 9, "tip": 10}, t_s_name=name, t_s_type=typ, image_size=[S, S],
                  turn_off_border=border_off)
        ts.load_reference_images()
        ts.setup_camera_info()
        nd = ts.no_deformation_dep.astype(np.float32)
        cur = nd.copy()
        yy, xx = np.mgrid[0:S, 0:S]
        cur -= (0.06 * np.exp(-((xx - 0.4 * S) ** 2 + (yy - 0.55 * S) ** 2) / (0.02 * S * S))).astype(np.float32)     # a dent deeper than the clip
        cur += rng.uniform(-2.5e-4, 2.5e-4, (S, S)).astype(np.float32)                                               # around the 1e-4 dead band
        cur[: S // 8] += np.float32(0.013)                                                                           # behind the skin: |diff| counts too
        seg = np.full((S, S), -1, dtype=np.int64)
        seg[-(S // 16):, :] = 4 + ((9 + 1) << 24)                                                                     # rows where the sensor body is seen
        pb.getCameraImage = lambda w, h, view, proj, renderer=None, flags=None: (w, h, np.zeros((h, w, 4), np.uint8), cur.copy(), seg.copy())
        img = ts.t_s_camera()
        key = "sensor_%s_%s_%d" % (name, typ, S)
        out[key + "_cam"] = np.concatenate([captured["eye"], captured["target"], captured["up"], [ts.fov, ts.focal_dist, ts.nearplane, ts.farplane]])
        out[key + "_cur"], out[key + "_seg_body"], out[key + "_img"] = cur, (seg >= 0), img
        assert img.dtype == np.uint8 and img.shape == (S, S)

    # ---- J. tcp_velocity_control (R3) run from the reference source (BaseRobotArm for the ur5, MG400's override): check_TCP_vel_lims,
    # work -> world twist, [jac_t; jac_r] stacking, matrix_rank -> inv / pinv, the MG400's slaved joints.  The kinematic INPUTS
    # (TCP pose, Jacobian at the TCP link's inertial frame) come from the oracle for the arm at hand and are stored, so the replay
    # first checks that they still are what the oracle computes, then compares the joint-velocity targets.
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import oracle as O
    O.build()
    MG400 = ref_class(os.path.join(REF, "tactile_gym", "robots", "arms", "mg400", "mg400.py"), "MG400", (BaseRobotArm,))
    for arm_name, sensor, typ, cls in (("ur5", "tactip", "standard", BaseRobotArm), ("mg400", "digitac", "standard", MG400)):
        wpos_, wrpy_ = ([0.33, 0.0, 0.035] if arm_name == "mg400" else [0.65, 0.0, 0.035]), [-np.pi, 0.0, np.pi / 2]
        lims = np.array([[-0.01, 0.01], [-0.02, 0.02], [-0.1, 0.1], [0.0, 0.0], [0.0, 0.0], [-0.5, 0.5]])
        m = O.load_model(arm_name, sensor, typ, wpos_, wrpy_, lims)
        rest = O.rest_pose("edge_follow", arm_name, sensor, typ, m)
        n = m.ndof
        qs, vels, Js, poses, targets = [], [], [], [], []
        for k in range(8):
            q = np.array(rest[:n]) + rng.uniform(-0.15, 0.15, n) * (1.0 if arm_name == "ur5" else 0.3)
            if arm_name == "mg400":      # keep the parallelogram closed (mg400.py:111-120)
                q[n - 3], q[n - 2], q[n - 1] = q[1], -q[1], q[1] + q[2]
            P, Q = O.link_states(m, q)
            J = O.jacobian(m, q, m.tcp_link)
            v = rng.uniform(-0.01, 0.01, 6) * np.array([1, 1, 1, 5, 5, 5])
            pb = PB()
            sent = {}
            pb.calculateJacobian = lambda *a, J=J: (J[:3].tolist(), J[3:].tolist())
            pb.setJointMotorControlArray = lambda *a, **kw: sent.update(kw)
            pb.VELOCITY_CONTROL = 0
            arm = bare(cls, _pb=pb, robot_id=0, TCP_link_id=0, num_control_dofs=n, control_joint_ids=list(range(n)), vel_gain=1.0, max_force=1000.0,
                       robot_type="MG400")
            arm.set_workframe(wpos_, wrpy_)
            arm.set_TCP_lims(lims)
            arm.get_current_TCP_pos_vel_worldframe = lambda P=P, Q=Q: (P[m.tcp_link], np.array(PB.getEulerFromQuaternion(Q[m.tcp_link])), Q[m.tcp_link], np.zeros(3), np.zeros(3))
            arm.get_current_joint_pos_vel = lambda q=q: (list(q), [0.0] * n)
            arm.tcp_velocity_control(v)
            qs.append(q); vels.append(v); Js.append(J); poses.append(np.concatenate([P[m.tcp_link], Q[m.tcp_link]])); targets.append(np.array(sent["targetVelocities"], dtype=np.float64))
            assert sent["forces"] == [1000.0] * n and sent["velocityGains"] == [1.0] * n
        out["velctl_%s_lims" % arm_name] = lims
        out["velctl_%s_q" % arm_name], out["velctl_%s_v" % arm_name], out["velctl_%s_J" % arm_name] = np.array(qs), np.array(vels), np.array(Js)
        out["velctl_%s_pose" % arm_name], out["velctl_%s_target" % arm_name] = np.array(poses), np.array(targets)

    # ---- K. tcp_position_control (base_robot_arm.py:228-279) from the reference source, ur5: pose delta + check_TCP_pos_lims +
    # workframe_to_worldframe -> the IK target pose (calculateInverseKinematics itself is pybullet's: the stub records its
    # arguments and returns the current joints), position-motor settings
    lims = np.array([[-0.004, 0.004], [-0.02, 0.02], [0.002, 0.1], [0.0, 0.0], [0.0, 0.0], [-0.05, 0.05]])
    m = O.load_model("ur5", "tactip", "standard", [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], lims)
    rest = O.rest_pose("edge_follow", "ur5", "tactip", "standard", m)
    qs, deltas, ik_targets = [], [], []
    for k in range(8):
        q = np.array(rest[:6]) + rng.uniform(-0.02, 0.02, 6)
        P, Q = O.link_states(m, q)
        d = rng.uniform(-0.001, 0.001, 6) * np.array([1, 1, 1, 17, 17, 17])
        pb = PB()
        got = {}
        pb.calculateInverseKinematics = lambda rid, link, pos, orn, **kw: got.update(pos=np.array(pos), orn=np.array(orn), kw=kw) or tuple(q)
        pb.setJointMotorControlArray = lambda *a, **kw: got.update(motor=kw)
        pb.POSITION_CONTROL = 2
        arm = bare(BaseRobotArm, _pb=pb, robot_id=0, TCP_link_id=0, num_control_dofs=6, control_joint_ids=list(range(6)), vel_gain=1.0, pos_gain=1.0,
                   max_force=1000.0, rest_poses=rest)
        arm.set_workframe([0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2])
        arm.set_TCP_lims(lims)
        arm.get_current_TCP_pos_vel_worldframe = lambda P=P, Q=Q: (P[m.tcp_link], np.array(PB.getEulerFromQuaternion(Q[m.tcp_link])), Q[m.tcp_link], np.zeros(3), np.zeros(3))
        arm.tcp_position_control(d)
        assert got["kw"]["maxNumIterations"] == 100 and got["kw"]["residualThreshold"] == 1e-8
        assert got["motor"]["forces"] == [1000.0] * 6 and got["motor"]["positionGains"] == [1.0] * 6 and got["motor"]["targetVelocities"] == [0] * 6
        qs.append(q); deltas.append(d); ik_targets.append(np.concatenate([got["pos"], got["orn"]]))
    out["posctl_lims"], out["posctl_q"], out["posctl_delta"], out["posctl_ik_target"] = lims, np.array(qs), np.array(deltas), np.array(ik_targets)

    # ---- L. the random draws of reset() (R11), in the reference's own call order: each env's reset() run from its source with a
    # recording proxy around a gym <= 0.21 RandomState (tactile_gym_b200.seeding: RandomState seeded from sha512(str(seed)));
    # arm, scene and observation calls are no-ops.  The replay compares the engine's host-side draw functions, which must consume
    # the same stream in the same order for seeded runs to match the reference's.
    from tactile_gym_b200 import seeding

    class RecRNG:
        def __init__(self, seed):
            self.rs, self.log = seeding.np_random(seed)[0], []

        def __getattr__(self, name):
            f = getattr(self.rs, name)

            def call(*a, **k):
                r = f(*a, **k)
                self.log.append(float(r))
                return r
            return call

    class ScenePB(PB):
        def getNumJoints(self, uid):
            return 1

        def loadURDF(self, *a, **k):
            return 11

        def createCollisionShape(self, **k):
            return 12

        def createMultiBody(self, **k):
            return 13

    def run_resets(key, cls, seed, n_resets=3, extra_globals=None, **attrs):
        arm = bare(BaseRobotArm, _pb=PB())
        arm.set_workframe(attrs.get("workframe_pos", [0.65, 0.0, 0.035]), attrs.get("workframe_rpy", [-np.pi, 0.0, np.pi / 2]))
        arm.get_current_TCP_pos_vel_worldframe = lambda: (np.zeros(3), np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3), np.zeros(3))
        robot = types.SimpleNamespace(arm=arm, reset=lambda **k: None)
        rec = RecRNG(seed)
        env = bare(cls, _pb=ScenePB(), robot=robot, np_random=rec, reset_counter=0, reset_limit=10 ** 9, _env_step_counter=0, **attrs)
        env.get_step_data = lambda: (0.0, False)
        env.get_observation = lambda: None
        rows = []
        for k in range(n_resets):
            rec.log = []
            env.reset()
            rows.append(list(rec.log))
        assert len({len(r) for r in rows}) == 1
        out["draws_%s" % key], out["draws_%s_seed" % key] = np.array(rows, dtype=np.float64), np.array([seed])

    for sensor in ("tactip", "digit", "digitac"):
        run_resets("edge_%s" % sensor, EdgeFollowEnv, 101, noise_mode="rand_height", t_s_name=sensor, edge_pos=[0.65, 0.0, 0.0], edge_len=0.175,
                   edge_height=0.035, edge_stim_id=0, goal_indicator=1, embed_dist=0.0035)
    run_resets("edge_fixed", EdgeFollowEnv, 102, noise_mode="fixed_height", t_s_name="tactip", edge_pos=[0.65, 0.0, 0.0], edge_len=0.175,
               edge_height=0.035, edge_stim_id=0, goal_indicator=1, embed_dist=0.0035)
    surf_ns = {"OpenSimplex": FakeSimplex}
    BaseSurfaceEnvR = ref_class(os.path.join(surf, "base_surface_env.py"), "BaseSurfaceEnv", (BaseTactileEnv,), extra=surf_ns)
    SurfAutoR = ref_class(os.path.join(surf, "surface_follow_auto", "surface_follow_auto_env.py"), "SurfaceFollowAutoEnv", (BaseSurfaceEnvR,))
    SurfVertR = ref_class(os.path.join(surf, "surface_follow_vert", "surface_follow_vert_env.py"), "SurfaceFollowVertEnv", (BaseSurfaceEnvR,))
    for key, cls, mode, nmode in (("surface_xyzRxRy", SurfAutoR, "xyzRxRy", "simplex"), ("surface_yzRx", SurfAutoR, "yzRx", "simplex"),
                                  ("surface_xyz_none", SurfAutoR, "xyz", "none"), ("surface_vert_xRz", SurfVertR, "xRz", "simplex")):
        e0 = bare(cls, _pb=PB(), noise_mode=nmode, movement_mode=mode, well_designed_pos=[0.65, 0.0, 0.0])
        e0.setup_surface()
        run_resets(key, cls, 103, noise_mode=nmode, movement_mode=mode, reward_mode="dense", embed_dist=0.0025, surface_id=1, goal_indicator=2,
                   workframe_pos=[0.65, 0.0, 0.025], heightfield_data=np.zeros((64, 64)),      # init_surface_and_goal's zeros (:381)
                   **{k: v for k, v in e0.__dict__.items() if k not in ("_pb", "noise_mode", "movement_mode")})
    BalanceR = ref_class(os.path.join(obj, "object_balance", "object_balance_env.py"), "ObjectBalanceEnv", (BaseObjectEnv,),
                         extra={"plot_vector": lambda *a, **k: None})
    for key, rg, re_ in (("balance_rand", True, True), ("balance_fixed", False, False), ("balance_gravity_only", True, False)):
        run_resets(key, BalanceR, 104, rand_gravity=rg, rand_embed_dist=re_, t_s_name="tactip", object_mode="pole", obj_id=3, obj_base_width=0.1,
                   obj_base_height=0.0025, embed_dist=0.0035, init_obj_pos=[0.55, 0.0, 0.35], init_obj_orn=(0, 0, 0, 1), workframe_pos=np.array([0.55, 0.0, 0.35]),
                   workframe_rpy=np.array([0.0, 0.0, 0.0]), obj_tip_constraint_id=1, update_constraints=lambda: None)
    PushR = ref_class(os.path.join(obj, "object_push", "object_push_env.py"), "ObjectPushEnv", (BaseObjectEnv,),
                      extra={"OpenSimplex": FakeSimplex, "os": os})
    for key, ro, rm, tt in (("push_rand_simplex", True, True, "simplex"), ("push_fixed_simplex", False, False, "simplex"), ("push_rand_straight", True, True, "straight")):
        run_resets(key, PushR, 105, rand_init_orn=ro, rand_obj_mass=rm, traj_type=tt, obj_id=7, init_obj_pos=[0.55, -0.16, 0.04], traj_n_points=10,
                   traj_spacing=0.025, traj_max_perturb=0.1, traj_ids=list(range(100, 110)), obj_width=0.08, workframe_pos=np.array([0.55, -0.2, 0.04]))
    RollR = ref_class(os.path.join(obj, "object_roll", "object_roll_env.py"), "ObjectRollEnv", (BaseObjectEnv,), extra={"add_assets_path": lambda p_: p_})
    for key, a_, b_, c_ in (("roll_rand", True, True, True), ("roll_fixed", False, False, False), ("roll_pos_only", False, False, True)):
        run_resets(key, RollR, 106, rand_obj_size=a_, rand_embed_dist=b_, rand_init_obj_pos=c_, default_obj_radius=0.0025, embed_dist=0.0015, obj_id=5,
                   object_path="sphere.urdf", workframe_rpy=np.array([-np.pi, 0.0, np.pi / 2]), visualise_goal=False)

    # ---- M. Robot.blocking_move (robot.py:188-260, R11) from the reference source, driven by an ideal position servo instead of
    # pybullet (step_sim: every joint lands on its commanded target): the constant-velocity retargeting, the halving rule and the
    # exit test on the PRE-step pose / speeds.  Kinematics (TCP pose of a joint vector) from the oracle, as in section J.
    Robot = ref_class(os.path.join(REF, "tactile_gym", "robots", "arms", "robot.py"), "Robot")
    m = O.load_model("ur5", "tactip", "standard", [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], np.zeros((6, 2)))
    rest = np.array(O.rest_pose("edge_follow", "ur5", "tactip", "standard", m)[:6])
    for case, (cv, max_steps, dq) in enumerate(((0.001, 1000, np.array([0.004, -0.003, 0.0025, 0.001, -0.0015, 0.002])),
                                               (None, 10, np.array([0.0004, -0.0003, 0.00025, 0.0001, -0.00015, 0.0002])))):
        targ_j = rest + dq
        Pt, Qt = O.link_states(m, targ_j)
        state = {"q": rest.copy(), "qd": np.zeros(6), "cmd": targ_j.copy()}
        hist = []

        def tcp_world():
            P, Q = O.link_states(m, state["q"])
            return P[m.tcp_link], np.array(PB.getEulerFromQuaternion(Q[m.tcp_link])), Q[m.tcp_link], np.zeros(3), np.zeros(3)

        def set_motors(rid, ids, mode, targetPositions=None, **kw):
            state["cmd"] = np.array(targetPositions, dtype=np.float64)

        def step_sim():
            hist.append(np.concatenate([state["q"], state["qd"], state["cmd"]]))
            new = state["cmd"].copy()
            state["qd"], state["q"] = (new - state["q"]) * 240.0, new

        pb = PB(); pb.setJointMotorControlArray = set_motors; pb.POSITION_CONTROL = 2
        arm = types.SimpleNamespace(target_pos_worldframe=Pt[m.tcp_link], target_orn_worldframe=Qt[m.tcp_link], target_joints=targ_j,
                                    get_current_TCP_pos_vel_worldframe=tcp_world, get_current_joint_pos_vel=lambda: (state["q"].copy(), state["qd"].copy()),
                                    control_joint_ids=list(range(6)), num_control_dofs=6, pos_gain=1.0, vel_gain=1.0, robot_id=0)
        rb = bare(Robot, _pb=pb, arm=arm, robot_id=0)
        rb.step_sim = step_sim
        rb.blocking_move(max_steps=max_steps, constant_vel=cv)
        out["blocking_%d_hist" % case] = np.array(hist)               # per iteration: q, qd before the step, commanded joint target
        out["blocking_%d_target" % case] = np.concatenate([targ_j, Pt[m.tcp_link], Qt[m.tcp_link], [cv if cv is not None else -1.0, max_steps]])

    # ---- N. surface_follow-v2's own surface: noise_mode "vertical_simplex" (base_surface_env.py:60-63, 83-107, 183-194, 248-259,
    # 359-379, 459-506, 523-537, 556-563, 708-754) - the upright heightfield and everything derived from it, from the reference
    # source: setup_surface, reset_task (update_surface with the stand-in noise, make_goal), update_init_pose, the action ranges,
    # and the reward terms at given TCP poses.  (CPU oracle only so far: the CUDA path does not build this mode yet.)
    for arm_name, wd in (("mg400", [0.33, 0.0, 0.0]), ("ur5", [0.65, 0.0, 0.0])):
        for dirsign in (1, -1):
            arm = bare(BaseRobotArm, _pb=PB())
            wp = [wd[0], wd[1], 0.15 + 0.025]
            arm.set_workframe(wp, [-np.pi, 0.0, 0.0])
            ev = bare(SurfVertR, _pb=ScenePB(), noise_mode="vertical_simplex", movement_mode="xRz", reward_mode="dense", well_designed_pos=wd,
                      embed_dist=0.0025, robot=types.SimpleNamespace(arm=arm), surface_id=1, goal_indicator=2, termination_dist=0.01,
                      _max_steps=200, _env_step_counter=7, t_s_name="tactip", control_mode="TCP_velocity_control",
                      np_random=types.SimpleNamespace(randint=lambda hi: 77, choice=lambda opts, dirsign=dirsign: dirsign))
            ev.setup_surface()
            ev.heightfield_data = np.zeros((64, 64))
            ev.reset_task()
            init_pos, init_rpy = ev.update_init_pose()
            key = "vert_%s_%s" % (arm_name, "p" if dirsign > 0 else "m")
            out[key + "_h"], out[key + "_array"], out[key + "_normals"] = ev.heightfield_data.copy(), ev.surface_array.copy(), ev.surface_normals.copy()
            out[key + "_goal"] = np.concatenate([ev.goal_pos_worldframe, ev.goal_pos_workframe])
            out[key + "_init"] = np.concatenate([init_pos, init_rpy])
            out[key + "_bins"] = np.stack([ev.x_bins, ev.y_bins])
            out[key + "_surface_pos"] = np.array(ev.surface_pos, dtype=np.float64)
            poses, rows = [], []
            for k in range(8):
                p_ = np.array([wd[0] + rng.uniform(-0.01, 0.02), rng.uniform(-0.12, 0.12), 0.175 + rng.uniform(-0.002, 0.002)])
                r_ = np.array([-np.pi + rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(-0.5, 0.5)])
                if k == 0:
                    p_ = np.array(ev.goal_pos_worldframe) + np.array([0.002, -0.003, 0.001])      # inside the termination radius
                ev.cur_tcp_pos_worldframe, ev.cur_tcp_orn_worldframe = p_, PB.getQuaternionFromEuler(r_)
                ev.tip_i, ev.tip_j = ev.xy_to_surface_idx(p_[0], p_[1])
                rows.append([ev.z_dist_to_surface(), ev.cos_dist_to_surface_normal(), ev.dense_reward(), float(ev.termination()), ev.tip_i, ev.tip_j])
                poses.append(np.concatenate([p_, r_]))
            out[key + "_poses"], out[key + "_rows"] = np.array(poses), np.array(rows, dtype=np.float64)
    ev.get_act_dim = lambda: 2
    ev.workframe_directions = [0, -1, 0]
    ev.setup_action_space()
    acts = rng.uniform(-0.3, 0.3, (12, 2))
    out["act_surfvert_vertical_in"], out["act_surfvert_vertical_out"] = acts, np.array([ev.scale_actions(ev.encode_actions(a)) for a in acts])

    # ---- O. object_balance's reset as the reference runs it (reset_task + reset_object, object_balance_env.py:295-381): the episode's
    # gravity, the constraint pivot handed to changeConstraint, the pole's start pose and the one-off external force (point, vector)
    rows = []
    for seed in (201, 202, 203):
        pb = ScenePB()
        rec = {}
        pb.setGravity = lambda x, y, z: rec.update(gravity=(x, y, z))
        pb.changeConstraint = lambda cid, **kw: rec.update(pivot=tuple(kw["jointChildPivot"]))
        pb.applyExternalForce = lambda uid, link, force, pos, flags=None: rec.update(force=tuple(float(v) for v in force), fpos=tuple(float(v) for v in pos))
        pb.WORLD_FRAME = 1
        rng_ = RecRNG(seed)
        e = bare(BalanceR, _pb=pb, np_random=rng_, rand_gravity=True, rand_embed_dist=True, t_s_name="tactip", object_mode="pole", obj_id=3,
                 obj_base_width=0.1, obj_base_height=0.0025, embed_dist=0.0035, init_obj_pos=[0.55, 0.0, 0.35], init_obj_orn=(0, 0, 0, 1),
                 workframe_pos=np.array([0.55, 0.0, 0.35]), obj_tip_constraint_id=1)
        e.reset_task()
        e.reset_object()
        rows.append([seed, *rng_.log, rec["gravity"][2], *rec["pivot"], *e.init_obj_pos, *rec["force"], *rec["fpos"]])
    out["balance_reset_rows"] = np.array(rows, dtype=np.float64)

    # ---- P. object_push / object_roll reset scenes from the reference source: the cube's start pose and mass (reset_object,
    # object_push_env.py:196-229), the marble's radius, start position, work frame and TCP-frame goal (object_roll_env.py:176-256)
    rows = []
    for seed in (301, 302):
        pb = ScenePB(); rec = {}
        pb.changeDynamics = lambda uid, link, **kw: rec.update({k: v for k, v in kw.items() if k == "mass"})
        rng_ = RecRNG(seed)
        e = bare(PushR, _pb=pb, np_random=rng_, rand_init_orn=True, rand_obj_mass=True, obj_id=7, init_obj_pos=[0.55, -0.16, 0.04])
        e.reset_object()
        rows.append([seed, *rng_.log, *pb.bodies[7][0], *pb.bodies[7][1], rec["mass"]])
    out["push_reset_rows"] = np.array(rows, dtype=np.float64)
    rows = []
    for seed in (303, 304):
        pb = ScenePB()
        arm = bare(BaseRobotArm, _pb=PB())
        arm.set_workframe([0.65, 0.0, 0.0035], [-np.pi, 0.0, np.pi / 2])
        arm.get_current_TCP_pos_vel_worldframe = lambda: (np.array([0.65, 0.0, 0.004]), np.zeros(3), np.array(PB.getQuaternionFromEuler([-np.pi, 0.0, np.pi / 2])), np.zeros(3), np.zeros(3))
        rng_ = RecRNG(seed)
        e = bare(RollR, _pb=pb, np_random=rng_, robot=types.SimpleNamespace(arm=arm), rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True,
                 default_obj_radius=0.0025, embed_dist=0.0015, obj_id=5, object_path="sphere.urdf", workframe_rpy=np.array([-np.pi, 0.0, np.pi / 2]),
                 visualise_goal=False)
        pb.loadURDF = lambda path, pos, orn, **kw: rec2.update(pos=tuple(pos), scaling=kw.get("globalScaling")) or 5
        rec2 = {}
        e.reset_task(); e.update_workframe(); e.reset_object(); e.make_goal()
        rows.append([seed, *rng_.log, e.scaled_obj_radius, *e.workframe_pos, *rec2["pos"], rec2["scaling"], *e.goal_pos_tcp, *e.goal_pos_worldframe])
    out["roll_reset_rows"] = np.array(rows, dtype=np.float64)

    np.savez_compressed(OUT, **out)
    print("wrote %s: %d arrays" % (OUT, len(out)))


if __name__ == "__main__":
    main()

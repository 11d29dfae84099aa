"""ctypes front-end of the CPU oracle (oracle/tg_oracle.c) plus the task-level restatement.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

Parity status: raster + kinematics pinned by the reference's fixtures; the reference's own Python around its native calls pinned
by vectors computed by running its source (tests/golden/reference_numpy.npz); pybullet's dynamics / IK / contacts PARITY UNPINNED
(pybullet not available) - see oracle/tg_oracle.h.
"""
import ctypes as C
import hashlib
import json
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(HERE, "..", "tactile_gym_b200", "assets")  # compiled DATA only (json/npz)

MAXL, MAXD = 16, 8


class OrModel(C.Structure):
    _fields_ = [
        ("nlinks", C.c_int), ("ndof", C.c_int),
        ("parent", C.c_int * MAXL), ("jtype", C.c_int * MAXL), ("dof_of_link", C.c_int * MAXL), ("link_of_dof", C.c_int * MAXD),
        ("joint_xyz", (C.c_double * 3) * MAXL), ("joint_rpy", (C.c_double * 3) * MAXL), ("axis", (C.c_double * 3) * MAXL),
        ("inertial_xyz", (C.c_double * 3) * MAXL), ("inertial_rpy", (C.c_double * 3) * MAXL),
        ("mass", C.c_double * MAXL), ("inertia", (C.c_double * 3) * MAXL),
        ("tcp_link", C.c_int), ("body_link", C.c_int),
        ("gravity", C.c_double * 3), ("dt", C.c_double), ("solver_iters", C.c_int), ("solver_residual_threshold", C.c_double),
        ("lin_damping", C.c_double), ("ang_damping", C.c_double), ("joint_damping", C.c_double),
        ("workframe_pos", C.c_double * 3), ("workframe_rpy", C.c_double * 3), ("tcp_lims", (C.c_double * 2) * 6),
        ("max_force", C.c_double), ("pos_gain", C.c_double), ("vel_gain", C.c_double), ("mg400_slave", C.c_int),
        ("cam_pos", C.c_double * 3), ("cam_rpy", C.c_double * 3), ("fov_deg", C.c_double), ("focal_dist", C.c_double),
        ("near_", C.c_double), ("far_", C.c_double),
    ]


class OrState(C.Structure):
    _fields_ = [
        ("q", C.c_double * MAXD), ("qd", C.c_double * MAXD), ("motor_mode", C.c_int * MAXD),
        ("target_pos", C.c_double * MAXD), ("target_vel", C.c_double * MAXD), ("kp", C.c_double * MAXD),
        ("kd", C.c_double * MAXD), ("max_force", C.c_double * MAXD),
    ]


_lib = None


def build(force=False):
    so = os.path.join(HERE, "libtg_oracle.so")
    src = os.path.join(HERE, "tg_oracle.c")
    if force or not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libtg_oracle.so")
        if not os.path.isfile(so):
            build()
        _lib = C.CDLL(so)
        _lib.or_robot_reset.restype = C.c_int
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def load_model(arm, sensor, typ, workframe_pos, workframe_rpy, tcp_lims, gravity=(0, 0, -9.81), dt=1.0 / 240.0):
    """Scene constants: base_tactile_env.py:125-130 (gravity, 150 iterations), base_robot_arm.py:24-25
    (damping 0.04 / 0.04 / 0.01), ur5.py:19-21 & mg400.py:27-29 (max_force 1000, gains 1)."""
    with open(os.path.join(ASSETS, "models", "%s_%s_%s.json" % (arm, typ, sensor))) as f:
        mj = json.load(f)
    with open(os.path.join(ASSETS, "sensors.json")) as f:
        sj = json.load(f)[sensor]
    m = OrModel()
    links = mj["links"]
    m.nlinks = len(links)
    names = [l["link_name"] for l in links]
    nd = 0
    for i, l in enumerate(links):
        m.parent[i] = l["parent"]
        m.jtype[i] = 1 if l["joint_type"] == 1 else 0
        m.dof_of_link[i] = -1
        if l["joint_type"] == 1:
            m.dof_of_link[i] = nd
            m.link_of_dof[nd] = i
            nd += 1
        for c in range(3):
            m.joint_xyz[i][c] = l["joint_xyz"][c]
            m.joint_rpy[i][c] = l["joint_rpy"][c]
            m.axis[i][c] = l["axis"][c]
            m.inertial_xyz[i][c] = l["inertial_xyz"][c]
            m.inertial_rpy[i][c] = l["inertial_rpy"][c]
            m.inertia[i][c] = l["inertia_diag"][c]
        m.mass[i] = l["mass"]
    m.ndof = nd
    m.tcp_link = names.index("tcp_link")
    m.body_link = names.index(sensor + "_body_link")
    for c in range(3):
        m.gravity[c] = gravity[c]
        m.workframe_pos[c] = workframe_pos[c]
        m.workframe_rpy[c] = workframe_rpy[c]
        m.cam_pos[c] = sj["types"][typ]["cam_pos"][c]
        m.cam_rpy[c] = sj["types"][typ]["cam_rpy"][c]
    for i in range(6):
        m.tcp_lims[i][0] = tcp_lims[i][0]
        m.tcp_lims[i][1] = tcp_lims[i][1]
    m.dt = dt
    m.solver_iters = 150
    m.solver_residual_threshold = 1e-7
    m.lin_damping = 0.04
    m.ang_damping = 0.04
    m.joint_damping = 0.01
    m.max_force, m.pos_gain, m.vel_gain = 1000.0, 1.0, 1.0
    m.mg400_slave = 1 if arm == "mg400" else 0
    m.fov_deg, m.focal_dist, m.near_, m.far_ = sj["fov"], sj["focal_dist"], sj["near"], sj["far"]
    m._names = names
    m._control_links = [i for i, l in enumerate(links) if l["joint_type"] == 1]
    return m


def load_refimg(sensor, typ, S):
    d = np.load(os.path.join(ASSETS, "refimg", "%s_%s_%d.npz" % (sensor, typ, S)))
    return (np.ascontiguousarray(d["nodef_dep"], dtype=np.float32), np.ascontiguousarray(d["nodef_gray"], dtype=np.float32),
            np.ascontiguousarray(d["border_mask"], dtype=np.uint8))


def rest_pose(env, arm, sensor, typ, model):
    with open(os.path.join(ASSETS, "rest_poses.json")) as f:
        rp = json.load(f)[env][arm]
    rp = rp[sensor][typ] if sensor in rp else rp[typ]
    rp = np.asarray(rp, dtype=np.float64)
    return rp[model._control_links].copy()


# ---------------------------------------------------------------- thin wrappers
def link_states(m, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    P = np.zeros((MAXL, 3)); Q = np.zeros((MAXL, 4))
    lib().or_link_states(C.byref(m), _dptr(q), _dptr(P), _dptr(Q))
    return P[: m.nlinks], Q[: m.nlinks]


def link_frames(m, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    P = np.zeros((MAXL, 3)); R = np.zeros((MAXL, 9))
    lib().or_link_frames(C.byref(m), _dptr(q), _dptr(P), _dptr(R))
    return P[: m.nlinks], R[: m.nlinks].reshape(-1, 3, 3)


def jacobian(m, q, link):
    q = np.ascontiguousarray(q, dtype=np.float64)
    J = np.zeros((6, MAXD))
    lib().or_jacobian(C.byref(m), _dptr(q), C.c_int(link), _dptr(J))
    return J[:, : m.ndof]


def inverse_dynamics(m, q, qd, qdd=None):
    q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
    tau = np.zeros(MAXD)
    qa = np.ascontiguousarray(qdd, dtype=np.float64) if qdd is not None else None
    lib().or_inverse_dynamics(C.byref(m), _dptr(q), _dptr(qd), _dptr(qa) if qa is not None else None, _dptr(tau))
    return tau[: m.ndof]


def forward_dynamics(m, q, qd, tau, with_damping=False):
    q = np.ascontiguousarray(q, dtype=np.float64); qd = np.ascontiguousarray(qd, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    out = np.zeros(MAXD)
    lib().or_forward_dynamics(C.byref(m), _dptr(q), _dptr(qd), _dptr(tau), C.c_int(int(with_damping)), _dptr(out))
    return out[: m.ndof]


def mass_matrix_inverse(m, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    A = np.zeros((MAXD, MAXD))
    lib().or_mass_matrix_inverse(C.byref(m), _dptr(q), _dptr(A))
    return A[: m.ndof, : m.ndof]


def tcp_pose_workframe(m, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    p = np.zeros(3); r = np.zeros(3)
    lib().or_tcp_pose_workframe(C.byref(m), _dptr(q), _dptr(p), _dptr(r))
    return p, r


def quat_from_euler(rpy):
    rpy = np.ascontiguousarray(rpy, dtype=np.float64); q = np.zeros(4)
    lib().or_quat_from_euler(_dptr(rpy), _dptr(q))
    return q


def euler_from_quat(q):
    q = np.ascontiguousarray(q, dtype=np.float64); r = np.zeros(3)
    lib().or_euler_from_quat(_dptr(q), _dptr(r))
    return r


def camera_frame(m, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    e, f, u, r = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
    lib().or_camera_frame(C.byref(m), _dptr(q), _dptr(e), _dptr(f), _dptr(u), _dptr(r))
    return e, f, u, r


def depth_image(eye, fwd, up, right, fov, near, far, S, tris):
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    out = np.zeros((S, S), dtype=np.float32)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (eye, fwd, up, right)]
    lib().or_depth_image(_dptr(a[0]), _dptr(a[1]), _dptr(a[2]), _dptr(a[3]), C.c_double(fov), C.c_double(near), C.c_double(far),
                         C.c_int(S), tris.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(len(tris)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def postprocess(cur_depth, refimg, border_on=True):
    """t_s_camera's arithmetic (tactile_sensor.py:268-292) on a given float32 depth image"""
    dep, gray, mask = refimg
    S = dep.shape[0]
    cur = np.ascontiguousarray(cur_depth, dtype=np.float32)
    img = np.zeros((S, S), dtype=np.uint8)
    lib().or_postprocess(C.c_int(S), cur.ctypes.data_as(C.POINTER(C.c_float)), dep.ctypes.data_as(C.POINTER(C.c_float)),
                         gray.ctypes.data_as(C.POINTER(C.c_float)), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(int(border_on)),
                         img.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return img


def tactile_image(m, q, S, tris_world, refimg, border_on=True, want_depth=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    tris = np.ascontiguousarray(tris_world, dtype=np.float64).reshape(-1, 9)
    dep, gray, mask = refimg
    img = np.zeros((S, S), dtype=np.uint8)
    dout = np.zeros((S, S), dtype=np.float32) if want_depth else None
    lib().or_tactile_image(C.byref(m), _dptr(q), C.c_int(S), _dptr(tris), C.c_int(len(tris)),
                           dep.ctypes.data_as(C.POINTER(C.c_float)), gray.ctypes.data_as(C.POINTER(C.c_float)),
                           mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(int(border_on)),
                           img.ctypes.data_as(C.POINTER(C.c_ubyte)), dout.ctypes.data_as(C.POINTER(C.c_float)) if want_depth else None)
    return (img, dout) if want_depth else img


def mul_transforms(pa, qa, pb, qb):
    po, qo = np.zeros(3), np.zeros(4)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pa, qa, pb, qb)]
    lib().or_mul_transforms(_dptr(a[0]), _dptr(a[1]), _dptr(a[2]), _dptr(a[3]), _dptr(po), _dptr(qo))
    return po, qo


def mat_from_quat(q):
    R = np.zeros(9)
    lib().or_mat_from_quat(_dptr(np.ascontiguousarray(q, dtype=np.float64)), _dptr(R))
    return R.reshape(3, 3)


def world_to_work(m, pos, quat):
    """worldframe_to_workframe (robots/arms/base_robot_arm.py:62-74): pose -> rpy -> quaternion -> inverse workframe -> rpy"""
    wq = quat_from_euler(np.array(m.workframe_rpy[:]))
    iq = np.array([-wq[0], -wq[1], -wq[2], wq[3]])                 # invertTransform
    ip = -(mat_from_quat(iq) @ np.array(m.workframe_pos[:]))
    q2 = quat_from_euler(euler_from_quat(quat))
    po, qo = mul_transforms(ip, iq, pos, q2)
    return po, euler_from_quat(qo)


def world_to_work_vec(m, v):
    """worldvec_to_workvec / worldvel_to_workvel (base_robot_arm.py:88-118): the matrix of the inverted workframe quaternion"""
    wq = quat_from_euler(np.array(m.workframe_rpy[:]))
    return mat_from_quat(np.array([-wq[0], -wq[1], -wq[2], wq[3]])) @ np.asarray(v, dtype=np.float64)


def tcp_state_workframe(m, s):
    """get_current_TCP_pos_vel_workframe (base_robot_arm.py:136-172): pos, rpy, orn, lin vel, ang vel of the TCP link's
    inertial frame (getLinkState(computeLinkVelocity=True) items 0, 1, 6, 7) in the work frame"""
    q = np.array(s.q[: m.ndof]); qd = np.array(s.qd[: m.ndof])
    P, Q = link_states(m, q)
    lin, ang = np.zeros(3), np.zeros(3)
    lib().or_link_velocity(C.byref(m), _dptr(np.ascontiguousarray(q)), _dptr(np.ascontiguousarray(qd)), C.c_int(m.tcp_link), _dptr(lin), _dptr(ang))
    pos, rpy = world_to_work(m, P[m.tcp_link], Q[m.tcp_link])
    return pos, rpy, quat_from_euler(rpy), world_to_work_vec(m, lin), world_to_work_vec(m, ang)


def object_state_workframe(m, o):
    """get_obj_pos_workframe / get_obj_vel_workframe (rl_envs/nonprehensile_manipulation/base_object_env.py:118-139)"""
    pos, rpy = world_to_work(m, np.array(o.pos[:]), np.array(o.quat[:]))
    return pos, rpy, quat_from_euler(rpy), world_to_work_vec(m, np.array(o.vel[:])), world_to_work_vec(m, np.array(o.omg[:]))


# ---------------------------------------------------------------- gym <= 0.21 seeding
def gym_np_random(seed):
    """gym.utils.seeding.np_random of gym <= 0.21 (base_tactile_env.py:61-64): a numpy RandomState seeded
    with the 32-bit words of sha512(str(seed))[:8].  [EXT: gym is unpinned in requirements.txt; the reference
    calls RandomState-only methods (randint), so the RandomState era is the one restated.]"""
    if seed is None:
        seed = int.from_bytes(os.urandom(4), "little")
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    h += b"\0" * 4  # gym's _bigint_from_bytes pads len%4==0 input with 4 zero bytes
    words = struct.unpack("%dI" % (len(h) // 4), h)
    big = sum(2 ** (32 * i) * v for i, v in enumerate(words))
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    rng = np.random.RandomState()
    rng.seed(ints or [0])
    return rng


# ---------------------------------------------------------------- edge_follow task restatement
class EdgeFollowOracle:
    """Restates EdgeFollowEnv (rl_envs/exploration/edge_follow/edge_follow_env.py) + BaseTactileEnv.step
    (rl_envs/base_tactile_env.py:166-185) on top of the C oracle.  One env instance."""

    def __init__(self, image_size=128, arm="ur5", sensor="tactip", max_steps=200, movement_mode="xy",
                 noise_mode="rand_height", seed=None, reward_mode="dense", control_mode="TCP_velocity_control"):
        self.reward_mode, self.control_mode = reward_mode, control_mode
        self.S, self.arm, self.sensor, self.typ = image_size, arm, sensor, "standard"
        self.max_steps, self.movement_mode, self.noise_mode = max_steps, movement_mode, noise_mode
        lims = np.zeros((6, 2))
        if arm == "mg400":  # edge_follow_env.py:73-80
            self.edge_pos = [0.33, 0.0, 0.0]; self.edge_len = 0.105
            lims[0], lims[1], lims[2], lims[5] = (-0.15, 0.15), (-0.11, 0.11), (-0.1, 0.1), (-np.pi, np.pi)
            stim = "short_edge"
        else:  # :81-88
            self.edge_pos = [0.65, 0.0, 0.0]; self.edge_len = 0.175
            lims[0], lims[1], lims[2], lims[5] = (-0.175, 0.175), (-0.175, 0.175), (-0.1, 0.1), (-np.pi, np.pi)
            stim = "long_edge"
        self.edge_height = 0.035
        self.embed_dist = 0.0035
        self.workframe_pos = np.array([self.edge_pos[0], self.edge_pos[1], self.edge_height])
        self.workframe_rpy = np.array([-np.pi, 0.0, np.pi / 2])
        self.m = load_model(arm, sensor, self.typ, self.workframe_pos, self.workframe_rpy, lims)
        self.rest = rest_pose("edge_follow", arm, sensor, self.typ, self.m)
        self.ref = load_refimg(sensor, self.typ, image_size)
        self.tris_local = np.load(os.path.join(ASSETS, "stimuli", stim + ".npz"))["tris"]
        self.s = OrState()
        self.repeat = int(np.floor((1.0 / 10.0) / (1.0 / 240.0)))
        self.termination_dist = 0.01
        self.np_random = gym_np_random(seed)
        self.max_pos_vel, self.max_ang_vel = 0.01, 5.0 * (np.pi / 180)
        self.steps = 0
        self.last_reset_substeps = 0

    def seed(self, seed):
        self.np_random = gym_np_random(seed)

    # edge_follow_env.py:285-299
    def draw(self):
        embed = self.embed_dist
        if self.noise_mode == "rand_height":
            lo, hi = {"tactip": (0.0015, 0.0065), "digit": (0.0011, 0.0028), "digitac": (0.0015, 0.0045)}[self.sensor]
            embed = self.np_random.uniform(lo, hi)
        ang = self.np_random.uniform(-np.pi, np.pi)
        return embed, ang

    def reset(self, draws=None):
        self.steps = 0
        self.embed_dist, self.edge_ang = self.draw() if draws is None else draws
        # update_edge :237-283
        c, s_ = np.cos(self.edge_ang), np.sin(self.edge_ang)
        self.goal_pos = np.array([self.edge_pos[0] + self.edge_len * c, self.edge_pos[1] + self.edge_len * s_, self.edge_pos[2] + self.edge_height])
        self.edge_end_points = np.array([
            [self.edge_pos[0] - self.edge_len * c, self.edge_pos[1] - self.edge_len * s_, self.edge_pos[2] + self.edge_height],
            [self.edge_pos[0] + self.edge_len * c, self.edge_pos[1] + self.edge_len * s_, self.edge_pos[2] + self.edge_height]])
        pos = np.array([0.0, 0.0, self.embed_dist]); rpy = np.zeros(3)
        self.last_reset_substeps = lib().or_robot_reset(C.byref(self.m), C.byref(self.s), _dptr(self.rest), _dptr(pos), _dptr(rpy))
        self.reward, self.done = self.step_data()
        return self.observation()

    def stimulus_world(self):
        q = quat_from_euler([0.0, 0.0, self.edge_ang])
        R = np.zeros(9); lib().or_mat_from_quat(_dptr(q), _dptr(R)); R = R.reshape(3, 3)
        return self.tris_local @ R.T + np.array(self.edge_pos)

    def observation(self):
        q = np.array(self.s.q[: self.m.ndof])
        return tactile_image(self.m, q, self.S, self.stimulus_world(), self.ref, border_on=True)[..., None]

    def tcp_world(self):
        P, Q = link_states(self.m, np.array(self.s.q[: self.m.ndof]))
        return P[self.m.tcp_link], Q[self.m.tcp_link]

    def step_data(self):
        p, _ = self.tcp_world()
        goal_dist = np.linalg.norm(p[:2] - self.goal_pos[:2])
        p1, p2 = self.edge_end_points[0, :2], self.edge_end_points[1, :2]
        a_, b_ = p2 - p1, p1 - p[:2]
        edge_dist = np.abs(a_[0] * b_[1] - a_[1] * b_[0]) / np.linalg.norm(p2 - p1)
        done = bool(goal_dist < self.termination_dist or self.steps >= self.max_steps)
        if self.reward_mode == "sparse":   # sparse_reward :430-438
            return (1.0 if goal_dist < self.termination_dist else 0.0), done
        return -(1.0 * goal_dist + 10.0 * edge_dist), done

    def oracle_obs(self):   # get_oracle_obs, edge_follow_env.py:454-476
        pos, _, _, lin, _ = tcp_state_workframe(self.m, self.s)
        goal, _ = world_to_work(self.m, self.goal_pos, np.array([0.0, 0.0, 0.0, 1.0]))
        return np.hstack([pos, lin, goal, self.edge_ang])

    def encode_scale(self, action):
        enc = np.zeros(6)
        a = np.asarray(action, dtype=np.float64)
        idx = {"xy": [0, 1], "xyz": [0, 1, 2], "xyRz": [0, 1, 5], "xyzRz": [0, 1, 2, 5]}[self.movement_mode]
        enc[idx] = a
        enc = np.clip(enc, -0.25, 0.25)
        amax = np.array([self.max_pos_vel] * 3 + [0.0, 0.0, self.max_ang_vel])
        if self.control_mode == "TCP_position_control":   # edge_follow_env.py:143-153
            amax = np.array([0.001] * 3 + [0.0, 0.0, 1 * (np.pi / 180)])
        amin = -amax
        return (((enc - (-0.25)) * (amax - amin)) / 0.5) + amin

    def step(self, action):
        v = np.ascontiguousarray(self.encode_scale(action), dtype=np.float64)
        self.steps += 1
        if self.control_mode == "TCP_position_control":   # robot.py:156-186, _max_blocking_pos_move_steps = 10 (edge_follow_env.py:38)
            self.last_move_substeps = lib().or_tcp_position_control(C.byref(self.m), C.byref(self.s), _dptr(v), C.c_int(10))
        else:
            lib().or_apply_action(C.byref(self.m), C.byref(self.s), _dptr(v), C.c_int(self.repeat))
        self.reward, self.done = self.step_data()
        return self.observation(), self.reward, self.done, {}


# ---------------------------------------------------------------- object_balance task restatement
class OrObject(C.Structure):
    _fields_ = [
        ("enabled", C.c_int), ("mass", C.c_double), ("inertia", C.c_double * 3), ("com_off", C.c_double * 3),
        ("pos", C.c_double * 3), ("quat", C.c_double * 4), ("vel", C.c_double * 3), ("omg", C.c_double * 3),
        ("ext_force", C.c_double * 3), ("ext_pos", C.c_double * 3), ("ext_pending", C.c_int),
        ("p2p_enabled", C.c_int), ("pivot_b", C.c_double * 3), ("erp", C.c_double), ("max_impulse", C.c_double),
    ]


def balance_draws(rng, sensor="tactip", rand_gravity=True, rand_embed_dist=True):
    """Random draws of one ObjectBalanceEnv.reset in the reference's order: reset_task (gravity, embed_dist;
    object_balance_env.py:300-313) then apply_random_force_base (choice, rand, choice, rand; :366-371)."""
    g = rng.uniform(-1.0, -0.1) if rand_gravity else -0.1
    embed_default = {"tactip": 0.0035, "digitac": 0.0015, "digit": 0.0015}[sensor]
    lo, hi = {"tactip": (0.003, 0.006), "digitac": (0.001, 0.0025), "digit": (0.0015, 0.0025)}[sensor]
    embed = rng.uniform(lo, hi) if rand_embed_dist else embed_default
    fx = rng.choice([-1, 1]) * rng.rand()
    fy = rng.choice([-1, 1]) * rng.rand()
    return np.array([g, embed, fx, fy])


class ObjectBalanceOracle:
    """Restates ObjectBalanceEnv (rl_envs/nonprehensile_manipulation/object_balance/object_balance_env.py, object_mode
    "pole") + BaseObjectEnv.reset (base_object_env.py) on top of the C oracle.  One env instance.

    Deviation from the reference, shared with the product: Robot.reset() is run WITHOUT the pole in the world.  In the
    reference the (fallen) pole of the previous episode still hangs on the constraint while the arm is repositioned and
    is only teleported back afterwards (base_object_env.py reset order); its influence on the arm is bounded by the
    solver residual (<= 3e-4 rad/s over <= 10 substeps, < 1e-5 rad) and is what makes resets independent of history."""

    def __init__(self, image_size=256, sensor="tactip", max_steps=250, movement_mode="xy", rand_gravity=True,
                 rand_embed_dist=True, seed=None, control_mode="TCP_velocity_control"):
        self.S, self.sensor, self.max_steps, self.movement_mode = image_size, sensor, max_steps, movement_mode
        self.control_mode = control_mode
        self.rand_gravity, self.rand_embed_dist = rand_gravity, rand_embed_dist
        self.workframe_pos = np.array([0.55, 0.0, 0.35]); self.workframe_rpy = np.zeros(3)
        lims = np.zeros((6, 2))
        lims[0], lims[1], lims[2] = (-0.1, 0.1), (-0.1, 0.1), (-0.1, 0.1)
        a45 = 45 * np.pi / 180
        lims[3], lims[4], lims[5] = (-a45, a45), (-a45, a45), (-a45, a45)
        self.m = load_model("ur5", sensor, "standard", self.workframe_pos, self.workframe_rpy, lims)
        self.rest = rest_pose("object_balance", "ur5", sensor, "standard", self.m)
        self.ref = load_refimg(sensor, "standard", image_size)
        self.tris_local = np.load(os.path.join(ASSETS, "stimuli", "pole.npz"))["tris"]
        with open(os.path.join(ASSETS, "objects", "pole.json")) as f:
            self.pole = json.load(f)
        self.s = OrState(); self.o = OrObject()
        self.repeat = int(np.floor((1.0 / 20.0) / (1.0 / 240.0)))
        self.base_w, self.base_h = 0.1, 0.0025
        self.init_rpy = np.array([0.0, 0.0, -np.pi / 2])
        self.np_random = gym_np_random(seed)
        self.steps = 0

    def seed(self, seed):
        self.np_random = gym_np_random(seed)

    def reset(self, draws=None):
        self.steps = 0
        d = balance_draws(self.np_random, self.sensor, self.rand_gravity, self.rand_embed_dist) if draws is None else np.asarray(draws, dtype=np.float64)
        g, self.embed_dist, fx, fy = d
        self.m.gravity[2] = g
        self.init_obj_pos = np.array([self.workframe_pos[0], self.workframe_pos[1], self.workframe_pos[2] + self.base_h / 2 - self.embed_dist])
        pos = np.zeros(3); rpy = np.zeros(3)
        self.last_reset_substeps = lib().or_robot_reset(C.byref(self.m), C.byref(self.s), _dptr(self.rest), _dptr(pos), _dptr(rpy))
        o = self.o
        o.enabled = 1; o.mass = self.pole["mass"]
        q0 = quat_from_euler(self.init_rpy)
        for c in range(3):
            o.inertia[c] = self.pole["inertia_diag"][c]; o.com_off[c] = self.pole["com_off"][c]
            o.pos[c] = self.init_obj_pos[c]; o.vel[c] = 0; o.omg[c] = 0
        for c in range(4):
            o.quat[c] = q0[c]
        # apply_random_force_base(force_mag=0.1) (:360-381)
        fpos = self.init_obj_pos + np.array([fx * self.base_w / 2, fy * self.base_w / 2, 0.0])
        for c in range(3):
            o.ext_force[c] = [0.0, 0.0, -0.1][c]; o.ext_pos[c] = fpos[c]
        o.ext_pending = 1
        o.p2p_enabled = 1; o.erp = 0.2; o.max_impulse = 500.0
        o.pivot_b[0], o.pivot_b[1], o.pivot_b[2] = 0.0, 0.0, -self.base_h / 2 + self.embed_dist
        self.reward, self.done = self.step_data()
        return self.observation()

    def stimulus_world(self):
        R = np.zeros(9); q = np.array(self.o.quat[:])
        lib().or_mat_from_quat(_dptr(q), _dptr(R)); R = R.reshape(3, 3)
        t = np.array(self.o.pos[:]) - R @ np.array(self.pole["base_com"])   # base LINK frame from the base COM frame
        return self.tris_local @ R.T + t

    def observation(self):
        q = np.array(self.s.q[: self.m.ndof])
        return tactile_image(self.m, q, self.S, self.stimulus_world(), self.ref, border_on=True)[..., None]

    def step_data(self):
        rpy_deg = euler_from_quat(np.array(self.o.quat[:])) * 180 / np.pi
        init_deg = self.init_rpy * 180 / np.pi
        rpy_dist = np.abs(((rpy_deg - init_deg) + 180) % 360 - 180)
        fall = bool(rpy_dist[0] > 35 or rpy_dist[1] > 35 or np.linalg.norm(np.array(self.o.pos[:]) - self.init_obj_pos) > 0.1)
        reward = (-1.0 if fall else 0.0) if getattr(self, "reward_mode", "dense") == "sparse" else 1.0   # :508-526
        return reward, bool(fall or self.steps >= self.max_steps)

    def oracle_obs(self):   # get_oracle_obs, object_balance_env.py:528-563
        pos, _, orn, lin, ang = tcp_state_workframe(self.m, self.s)
        op, _, oo, ol, oa = object_state_workframe(self.m, self.o)
        return np.hstack([pos, orn, lin, ang, op, oo, ol, oa])

    def encode_scale(self, action):
        enc = np.zeros(6); a = np.asarray(action, dtype=np.float64)
        idx = {"xy": [0, 1], "xyz": [0, 1, 2], "RxRy": [3, 4], "xyRxRy": [0, 1, 3, 4]}[self.movement_mode]
        enc[idx] = a
        enc = np.clip(enc, -0.25, 0.25)
        mv, ma = 0.01, 5.0 * (np.pi / 180)
        if self.control_mode == "TCP_position_control":   # object_balance_env.py:129-139: m / rad per step
            mv, ma = 0.001, 1 * (np.pi / 180)
        amax = np.array([mv, mv, mv, ma, ma, 0.0]); amin = -amax
        return (((enc - (-0.25)) * (amax - amin)) / 0.5) + amin

    def step(self, action):
        v = np.ascontiguousarray(self.encode_scale(action), dtype=np.float64)
        self.steps += 1
        if self.control_mode == "TCP_position_control":   # robot.py:156-186, _max_blocking_pos_move_steps = 10 (object_balance_env.py:36)
            self.last_move_substeps = lib().or_tcp_position_control_world(C.byref(self.m), C.byref(self.s), C.byref(self.o), None, _dptr(v), C.c_int(10))
        else:
            lib().or_tcp_velocity_control(C.byref(self.m), C.byref(self.s), _dptr(v))
            for _ in range(self.repeat):
                lib().or_step_sim_obj(C.byref(self.m), C.byref(self.s), C.byref(self.o))
        self.reward, self.done = self.step_data()
        return self.observation(), self.reward, self.done, {}


# ---------------------------------------------------------------- surface_follow task restatement
def surface_heights(seed_int, rows=64, cols=64, interp=0.05, rng=0.025):
    """gen_heigtfield_simplex_2d (base_surface_env.py:311-327) for OpenSimplex(seed=seed_int)"""
    h = np.zeros((rows, cols))
    lib().or_surface_heights(C.c_longlong(int(seed_int)), rows, cols, C.c_double(interp), C.c_double(rng), _dptr(h))
    return h


def heightfield_local_vertices(h, grid=0.006):
    """[EXT] what pybullet draws for createCollisionShape(GEOM_HEIGHTFIELD) (base_surface_env.py:402-424): the data is
    stored as float32, the shape is centred on the middle of its min/max height (btHeightfieldTerrainShape local origin)
    and on the grid centre, x = column index, y = row index (data index y * width + x); mesh vertices are float32."""
    hf = h.astype(np.float32)
    zc = 0.5 * (float(hf.min()) + float(hf.max()))
    rows, cols = h.shape
    x = ((np.arange(cols) - (cols - 1) / 2.0) * grid).astype(np.float32).astype(np.float64)
    y = ((np.arange(rows) - (rows - 1) / 2.0) * grid).astype(np.float32).astype(np.float64)
    z = (hf.astype(np.float64) - zc).astype(np.float32).astype(np.float64)
    V = np.zeros((rows, cols, 3))
    V[..., 0] = x[None, :]; V[..., 1] = y[:, None]; V[..., 2] = z
    return V


def heightfield_tris(V, i0, i1, j0, j1):
    """[EXT] btHeightfieldTerrainShape::processAllTriangles with flipQuadEdges / diamond / zigzag off: cell (x=j, y=i) ->
    (x,y),(x,y+1),(x+1,y) and (x+1,y),(x,y+1),(x+1,y+1).  Cells i0 <= i < i1, j0 <= j < j1."""
    out = []
    for i in range(i0, i1):
        for j in range(j0, j1):
            out.append([V[i, j], V[i + 1, j], V[i, j + 1]])
            out.append([V[i, j + 1], V[i + 1, j], V[i + 1, j + 1]])
    return np.array(out).reshape(-1, 3, 3)


class SurfaceFollowOracle:
    """Restates the three surface envs (rl_envs/exploration/surface_follow/surface_follow_{auto,goal,vert}/*_env.py) on
    BaseSurfaceEnv (rl_envs/exploration/surface_follow/base_surface_env.py).  One env instance.  The tip core <-> table contact (only reachable in the deepest valleys,
    SURVEY.md 8a R5) is not modelled."""

    def __init__(self, image_size=128, arm="ur5", sensor="digit", max_steps=200, movement_mode="xyzRxRy", seed=None, variant="auto",
                 noise_mode="simplex", reward_mode="dense", render=True, control_mode="TCP_velocity_control"):
        """variant "auto": SurfaceFollowAutoEnv (surface_follow-v0); "goal": SurfaceFollowGoalEnv (surface_follow-v1,
        surface_follow_goal/surface_follow_goal_env.py: the policy steers x / y, the reward adds the goal distance); "vert":
        SurfaceFollowVertEnv (surface_follow-v2, surface_follow_vert/surface_follow_vert_env.py: x steered, y driven, 10 / 3 weights).
        noise_mode "simplex" | "none" | "random" | "vertical_simplex" (the upright surface of -v2, `forward` sensors); movement modes
        yz / xyz / yzRx / xyzRxRy (+ xRz for "vert"); reward_mode dense | sparse; render=False skips the images."""
        self.control_mode = control_mode
        self.S, self.arm, self.sensor, self.typ = image_size, arm, sensor, "standard"
        self.max_steps, self.movement_mode, self.variant = max_steps, movement_mode, variant
        self.noise_mode, self.reward_mode, self.render = noise_mode, reward_mode, render
        self.one_d = movement_mode in ("yz", "yzRx", "xRz")
        self.grid, self.hrange, self.rows, self.cols, self.interp, self.extent = 0.006, 0.025, 64, 64, 0.05, 0.15
        wd = [0.33, 0.0, 0.0] if arm == "mg400" else [0.65, 0.0, 0.0]                  # base_surface_env.py:54-57
        self.embed_dist = {"tactip": 0.0025, "digitac": 0.0015, "digit": 0.0015}[sensor]  # :67-75
        self.surface_pos = np.array([wd[0], wd[1], self.hrange])                          # :261
        self.workframe_pos = np.array([wd[0], wd[1], self.hrange]); self.workframe_rpy = np.array([-np.pi, 0.0, np.pi / 2])  # :111-114
        lims = np.zeros((6, 2))
        lims[0], lims[1], lims[2] = (-self.extent, self.extent), (-self.extent, self.extent), (-self.hrange, self.hrange)
        lims[3], lims[4] = (-np.pi / 4, np.pi / 4), (-np.pi / 4, np.pi / 4)
        # noise_mode "vertical_simplex" (surface_follow-v2's own surface; CPU oracle only so far, the CUDA path does not build it):
        # the heightfield stands upright 0.15 m above the table, the `forward` sensor type faces it (:60-63, :83-107, :248-259)
        self.vertical = noise_mode == "vertical_simplex"
        self.R_flip = np.eye(3)
        self.original_surface_pos = self.surface_pos.copy()
        if self.vertical:
            self.typ = "forward"
            self.surface_pos = np.array([wd[0], wd[1], 0.15 + self.hrange])
            self.R_flip = mat_from_quat(quat_from_euler([0.0, -np.pi / 2, 0.0]))
            self.workframe_pos, self.workframe_rpy = self.surface_pos.copy(), np.array([-np.pi, 0.0, 0.0])
            lims = np.zeros((6, 2))
            lims[0], lims[1], lims[5] = (-self.hrange, self.hrange), (-self.extent, self.extent), (-np.pi / 4, np.pi / 4)
        self.m = load_model(arm, sensor, self.typ, self.workframe_pos, self.workframe_rpy, lims)
        self.rest = rest_pose("surface_follow", arm, sensor, self.typ, self.m)
        self.ref = load_refimg(sensor, self.typ, image_size)
        # :264-271 x/y bins
        sp = self.original_surface_pos      # (the vertical set-up builds the bins around the un-flipped surface position, :248-259)
        self.x_bins = np.linspace(sp[0] - (self.rows / 2) * self.grid, sp[0] + (self.rows / 2) * self.grid, self.rows)
        self.y_bins = np.linspace(sp[1] - (self.cols / 2) * self.grid, sp[1] + (self.cols / 2) * self.grid, self.cols)
        self.s = OrState()
        self.repeat = int(np.floor((1.0 / 10.0) / (1.0 / 240.0)))
        self.termination_dist = 0.01
        self.np_random = gym_np_random(seed)
        self.steps = 0
        self.last_reset_substeps = 0
        R = np.zeros(9); lib().or_mat_from_quat(_dptr(quat_from_euler(self.workframe_rpy)), _dptr(R)); self.Rw = R.reshape(3, 3)

    def seed(self, seed):
        self.np_random = gym_np_random(seed)

    def draw(self):
        """reset_task order (:539-547): update_surface's randint(1e8) (:448), then make_goal's uniform(-pi, pi) (:508)"""
        seed_int = self.np_random.randint(1e8) if self.noise_mode in ("simplex", "vertical_simplex") else 0      # :436-458
        if self.noise_mode == "random":
            # gen_heigtfield_noisey (:302-318): 32 x 32 uniform draws, columns outer, rows inner, each filling a 2 x 2 block -
            # drawn by update_surface BEFORE make_goal's draw; the heights ride in the first slot instead of a seed
            h = np.zeros((self.rows, self.cols))
            for j in range(self.cols // 2):
                for i in range(self.rows // 2):
                    height = self.np_random.uniform(0, self.hrange * 0.2)
                    h[2 * i, 2 * j] = h[2 * i + 1, 2 * j] = h[2 * i, 2 * j + 1] = h[2 * i + 1, 2 * j + 1] = height
            seed_int = h
        ang = float(self.np_random.choice([-1, 1])) if self.one_d else self.np_random.uniform(-np.pi, np.pi)   # :512-520
        return (seed_int if self.noise_mode == "random" else float(seed_int)), ang

    def xy_to_surface_idx(self, x, y):   # :273-288
        i = int(np.digitize(y, self.y_bins)); j = int(np.digitize(x, self.x_bins))
        if i == self.cols: i -= 1
        if j == self.rows: j -= 1
        return i, j

    def reset(self, draws=None):
        self.steps = 0
        seed_int, ang = self.draw() if draws is None else draws
        if self.vertical:                                                               # gen_heigtfield_simplex_1d_vertical :359-379
            col = np.array([opensimplex_noise2(int(seed_int), x * self.interp, 1 * self.interp) * self.hrange for x in range(self.rows)])
            self.h = np.tile(col[:, None], (1, self.cols))
        elif self.noise_mode == "random":                                               # :438-439
            self.h = np.array(seed_int, dtype=np.float64)
        elif self.noise_mode == "none" or self.movement_mode == "xRz":                 # :436-437; "xRz" is in neither list of :450-455
            self.h = np.zeros((self.rows, self.cols))
        elif self.one_d:                                                                # gen_heigtfield_simplex_1d :339-357
            row = np.array([opensimplex_noise2(int(seed_int), 1 * self.interp, y * self.interp) * self.hrange for y in range(self.cols)])
            self.h = np.tile(row, (self.rows, 1))
        else:
            self.h = surface_heights(int(seed_int), self.rows, self.cols, self.interp, self.hrange)
        self.accum_rew = 0.0                                                           # reset_task :590-591
        # update_surface :480-499: surface_array / normals
        X, Y = np.meshgrid(self.x_bins, self.y_bins)
        self.surface_array = np.dstack((X, Y, self.h + self.surface_pos[2]))
        gy, gx = np.gradient(self.h, self.grid)
        nrm = np.dstack((-gx, -gy, np.ones_like(self.h)))
        self.surface_normals = nrm / np.linalg.norm(nrm, axis=2)[..., None]
        self.V = heightfield_local_vertices(self.h, self.grid) + self.surface_pos
        if self.vertical:
            # :472-489 every surface point goes world -> surface frame (translation only), is turned by the surface's orientation
            # and comes back; :499-506 the normals are turned the same way.  The drawn mesh is the body at surface_pos / surface_orn.
            self.surface_array = (self.surface_array - self.surface_pos) @ self.R_flip.T + self.surface_pos
            self.surface_normals = self.surface_normals @ self.R_flip.T
            self.V = heightfield_local_vertices(self.h, self.grid) @ self.R_flip.T + self.surface_pos
        # make_goal :501-537
        self.dirs = np.array([0.0, ang, 0.0]) if self.one_d else np.array([np.cos(ang), np.sin(ang), 0.0])
        wdir = self.Rw @ self.dirs
        g = [self.original_surface_pos[0] + self.extent * wdir[0], self.original_surface_pos[1] + self.extent * wdir[1]]
        gi, gj = self.xy_to_surface_idx(g[0], g[1])
        self.goal_pos = self.surface_array[gi, gj].copy() if self.vertical else np.array([g[0], g[1], self.surface_array[gi, gj, 2]])   # :523-537
        # update_init_pose :549-573 + Robot.reset
        ch = self.h[self.rows // 2, self.cols // 2]
        init_world = np.array([self.surface_pos[0], self.surface_pos[1], self.surface_pos[2] + ch - self.embed_dist])
        if self.vertical:
            init_world = np.array([self.surface_pos[0] - (ch - self.embed_dist), self.surface_pos[1], self.surface_pos[2]])
        pos = self.Rw.T @ (init_world - self.workframe_pos); rpy = np.zeros(3)
        self.last_reset_substeps = lib().or_robot_reset(C.byref(self.m), C.byref(self.s), _dptr(self.rest), _dptr(np.ascontiguousarray(pos)), _dptr(rpy))
        self.reward, self.done = self.step_data()
        return self.observation() if self.render else None

    def tcp_world(self):
        P, Q = link_states(self.m, np.array(self.s.q[: self.m.ndof]))
        return P[self.m.tcp_link], Q[self.m.tcp_link]

    def stimulus_world(self, radius=0.06):
        """the heightfield cells within `radius` (m) of the TCP in x/y (everything the tactile camera can see)"""
        p, _ = self.tcp_world()
        if self.vertical:
            # the upright surface: grid columns (local x) run along world z, rows (local y) along world y
            local = (p - self.surface_pos) @ self.R_flip          # = R_flip^T (p - surface_pos): the TCP in the heightfield's own frame
            cx, cy = local[0] / self.grid + (self.cols - 1) / 2.0, local[1] / self.grid + (self.rows - 1) / 2.0
        else:
            cx = (p[0] - self.surface_pos[0]) / self.grid + (self.cols - 1) / 2.0
            cy = (p[1] - self.surface_pos[1]) / self.grid + (self.rows - 1) / 2.0
        r = radius / self.grid
        j0, j1 = int(max(0, np.floor(cx - r))), int(min(self.cols - 1, np.ceil(cx + r)))
        i0, i1 = int(max(0, np.floor(cy - r))), int(min(self.rows - 1, np.ceil(cy + r)))
        if i1 <= i0 or j1 <= j0:
            return np.zeros((0, 3, 3))
        return heightfield_tris(self.V, i0, i1, j0, j1)

    def observation(self):
        q = np.array(self.s.q[: self.m.ndof])
        return tactile_image(self.m, q, self.S, self.stimulus_world(), self.ref, border_on=True)[..., None]

    def step_data(self):   # :631-775 + surface_follow_auto_env.py:76-94
        p, qt = self.tcp_world()
        self.tip_i, self.tip_j = self.xy_to_surface_idx(p[0], p[1])
        R = np.zeros(9); lib().or_mat_from_quat(_dptr(np.ascontiguousarray(qt)), _dptr(R)); R = R.reshape(3, 3)
        done = bool(np.linalg.norm(p - self.goal_pos) < self.termination_dist or self.steps >= self.max_steps)
        emb = p + R @ np.array([0.0, 0.0, -self.embed_dist])
        surf_dist = abs(emb[2] - self.surface_array[self.tip_i, self.tip_j, 2])
        n = self.surface_normals[self.tip_i, self.tip_j]
        v = R @ np.array([0.0, 0.0, -1.0])
        if self.vertical:   # :708-711, :733-735, :752-754: the forward sensor's axis is its -x, the distance is measured along world x
            emb = p + R @ np.array([-self.embed_dist, 0.0, 0.0])
            surf_dist = abs(emb[0] - self.surface_array[self.tip_i, self.tip_j, 0])
            v = R @ np.array([-1.0, 0.0, 0.0])
        cos_dist = 1 - np.dot(n, v) / (np.linalg.norm(n) * np.linalg.norm(v))
        w_norm = 0.0 if self.movement_mode in ("yz", "xyz") else 1.0
        if self.variant == "vert":   # surface_follow_vert_env.py:63-79
            dense = -(10.0 * surf_dist + 3.0 * cos_dist)
        elif self.variant == "goal":   # surface_follow_goal_env.py:62-81
            goal_xy = np.linalg.norm(p[:2] - self.goal_pos[:2])
            dense = -(1.0 * goal_xy + 10.0 * surf_dist + w_norm * cos_dist)
        else:
            dense = -(1.0 * surf_dist + w_norm * cos_dist)
        if self.reward_mode == "sparse":   # sparse_reward (surface_follow_auto_env.py:59-73, surface_follow_goal_env.py:53-67)
            self.accum_rew += dense
            return (self.accum_rew if np.linalg.norm(p - self.goal_pos) < self.termination_dist else 0.0), done
        return dense, done

    def oracle_obs(self):   # get_oracle_obs, base_surface_env.py:789-819 (tip_i / tip_j from the last get_step_data)
        pos, _, orn, lin, ang = tcp_state_workframe(self.m, self.s)
        goal, _ = world_to_work(self.m, self.goal_pos, np.array([0.0, 0.0, 0.0, 1.0]))
        n = world_to_work_vec(self.m, self.surface_normals[self.tip_i, self.tip_j])
        return np.hstack([pos, orn, lin, ang, goal, self.surface_array[self.tip_i, self.tip_j, 2], n])

    def features(self):   # SurfaceFollowGoalEnv.get_extended_feature_array (surface_follow_goal_env.py:83-97)
        tp = tcp_state_workframe(self.m, self.s)[0]
        gp = self.Rw.T @ (self.goal_pos - self.workframe_pos)
        return np.concatenate([tp, gp])

    def encode_scale(self, action):   # surface_follow_auto_env.py:27-57, base_tactile_env.py:141-164
        enc = np.zeros(6); a = np.asarray(action, dtype=np.float64)
        k = {"tactip": 1.0, "digitac": 0.9, "digit": 0.7}[self.sensor]
        if self.variant == "vert":   # surface_follow_vert_env.py:30-45
            enc[1] = self.dirs[1] * 0.25 * k
            enc[0], enc[5] = a[0], a[1]
        elif self.variant == "goal":   # surface_follow_goal_env.py:27-52
            enc[{"yz": [1, 2], "xyz": [0, 1, 2], "yzRx": [1, 2, 3], "xyzRxRy": [0, 1, 2, 3, 4]}[self.movement_mode]] = a
        else:
            enc[0] = self.dirs[0] * 0.25 * k; enc[1] = self.dirs[1] * 0.25 * k
            enc[{"yz": [2], "xyz": [2], "yzRx": [2, 3], "xyzRxRy": [2, 3, 4]}[self.movement_mode]] = a
        enc = np.clip(enc, -0.25, 0.25)
        mv, ma = 0.01, 5.0 * (np.pi / 180)
        if self.control_mode == "TCP_position_control":   # base_surface_env.py:172-181
            mv, ma = 0.001, 1 * (np.pi / 180)
        amax = np.array([mv, mv, mv, ma, ma, 0.0]); amin = -amax
        if self.vertical:   # :183-194
            amax = np.array([mv, mv, 0.0, 0.0, 0.0, ma]); amin = -amax
        return (((enc - (-0.25)) * (amax - amin)) / 0.5) + amin

    def step(self, action):
        v = np.ascontiguousarray(self.encode_scale(action), dtype=np.float64)
        self.steps += 1
        if self.control_mode == "TCP_position_control":   # robot.py:156-186, _max_blocking_pos_move_steps = 10 (base_surface_env.py:28)
            self.last_move_substeps = lib().or_tcp_position_control(C.byref(self.m), C.byref(self.s), _dptr(v), C.c_int(10))
        else:
            lib().or_apply_action(C.byref(self.m), C.byref(self.s), _dptr(v), C.c_int(self.repeat))
        self.reward, self.done = self.step_data()
        return (self.observation() if self.render else None), self.reward, self.done, {}


# ---------------------------------------------------------------- object_push task restatement
OR_MAXC = 8


class OrPush(C.Structure):
    _fields_ = [
        ("half", C.c_double * 3), ("table_z", C.c_double), ("mu_table", C.c_double), ("mu_tip", C.c_double),
        ("tip_k", C.c_double), ("tip_d", C.c_double), ("erp", C.c_double), ("slop", C.c_double),
        ("lin_damping", C.c_double), ("ang_damping", C.c_double), ("tip_link", C.c_int), ("n_hull", C.c_int),
        ("hull", C.POINTER(C.c_double)), ("shape", C.c_int), ("radius", C.c_double), ("cyl_pos", C.c_double * 3),
        ("cyl_axis", C.c_double * 3), ("cyl_half_len", C.c_double), ("cyl_radius", C.c_double), ("warmstart", C.c_double), ("ws_n", C.c_int), ("ws_feature", C.c_int * OR_MAXC),
        ("ws_impulse", (C.c_double * 3) * OR_MAXC), ("n_contacts", C.c_int), ("n_iters", C.c_int),
        ("normal_impulse", C.c_double * OR_MAXC), ("contact_pos", (C.c_double * 3) * OR_MAXC),
    ]


def opensimplex_noise2(seed_int, x, y):
    perm = (C.c_short * 256)()
    lib().or_opensimplex_init(C.c_longlong(int(seed_int)), perm)
    lib().or_opensimplex_noise2.restype = C.c_double
    return lib().or_opensimplex_noise2(perm, C.c_double(x), C.c_double(y))


def push_draws(rng, rand_init_orn=False, rand_obj_mass=False, traj_type="simplex"):
    """Random draws of one ObjectPushEnv.reset in the reference's order: reset_object (init_obj_ang, obj_mass;
    object_push_env.py:204-229) then make_goal -> update_trajectory (OpenSimplex seed :289 / traj_ang :310)."""
    ang = rng.uniform(-np.pi / 32, np.pi / 32) if rand_init_orn else 0.0
    mass = rng.uniform(0.4, 0.8) if rand_obj_mass else 0.491
    third = float(rng.randint(1e8)) if traj_type == "simplex" else rng.uniform(-np.pi / 8, np.pi / 8)
    return np.array([ang, mass, third])


class ObjectPushOracle:
    """Restates ObjectPushEnv (rl_envs/nonprehensile_manipulation/object_push/object_push_env.py) + BaseObjectEnv
    (base_object_env.py) on top of the C oracle.  One env instance.  As for object_balance, Robot.reset() is run without
    the cube in the world (the reference repositions the arm with the previous episode's cube still lying around)."""

    def __init__(self, image_size=128, arm="mg400", sensor="digitac", max_steps=1000, movement_mode="TyRz", traj_type="simplex",
                 rand_init_orn=False, rand_obj_mass=False, reward_mode="dense", seed=None, control_mode="TCP_velocity_control"):
        self.S, self.arm, self.sensor, self.max_steps = image_size, arm, sensor, max_steps
        self.movement_mode, self.traj_type, self.reward_mode = movement_mode, traj_type, reward_mode
        self.control_mode = control_mode
        self.rand_init_orn, self.rand_obj_mass = rand_init_orn, rand_obj_mass
        self.typ = "right_angle"
        self.obj_w = self.obj_h = 0.08
        lims = np.zeros((6, 2))
        a45 = 45 * np.pi / 180
        if arm == "mg400":   # object_push_env.py:72-88
            if sensor == "tactip":
                self.typ = "mini_right_angle"                                          # :84-86
            lims[0], lims[1], lims[5] = (0.0, 0.3), (-0.1, 0.08), (-a45, a45)
            wd = np.array([0.30 if sensor == "tactip" else 0.25, -0.1, self.obj_h / 2])   # :82-88
        else:                # :89-101
            lims[0], lims[1], lims[5] = (0.0, 0.3), (-0.1, 0.1), (-a45, a45)
            wd = np.array([0.55, -0.20, self.obj_h / 2])
        self.workframe_pos, self.workframe_rpy = wd, np.array([-np.pi, 0.0, np.pi / 2])
        self.m = load_model(arm, sensor, self.typ, self.workframe_pos, self.workframe_rpy, lims)
        self.rest = rest_pose("object_push", arm, sensor, self.typ, self.m)
        self.ref = load_refimg(sensor, self.typ, image_size)
        self.tris_local = np.load(os.path.join(ASSETS, "stimuli", "cube.npz"))["tris"]
        with open(os.path.join(ASSETS, "objects", "cube.json")) as f:
            self.cube = json.load(f)
        self.hull = np.ascontiguousarray(np.load(os.path.join(ASSETS, "models", "%s_%s_%s_meshes.npz" % (arm, self.typ, sensor)))["tip_core_hull"], dtype=np.float64)
        self.init_obj_pos = np.array([wd[0], wd[1] + self.obj_w / 2, self.obj_h / 2])       # :193
        self.s, self.o, self.p = OrState(), OrObject(), OrPush()
        p = self.p
        dyn = {"tactip": (50, 100, 10.0), "digitac": (300, 100, 10.0), "digit": (50, 200, 10.0)}[sensor]   # :61-66
        cube_mu, table_mu = 0.065, 1.0                                                       # :218, table.urdf
        for c in range(3):
            p.half[c] = 0.04
        p.table_z = 0.0
        p.mu_table, p.mu_tip = cube_mu * table_mu, min(cube_mu * dyn[2], 10.0)
        # [EXT] combined stiffness 1/(1/kA + 1/kB) with the cube at bullet's default 1e18; combined damping dA + dB, default 0.1
        p.tip_k, p.tip_d = 1.0 / (1.0 / dyn[0] + 1.0 / 1e18), dyn[1] + 0.1
        p.erp, p.slop = 0.2, 1e-4
        p.lin_damping, p.ang_damping = 0.04, 0.04
        p.warmstart, p.ws_n = 0.0, 0   # warm starting off (the device path has none yet)
        p.tip_link = self.m._names.index(sensor + "_tip_link")
        p.n_hull = len(self.hull)
        p.hull = self.hull.ctypes.data_as(C.POINTER(C.c_double))
        self.repeat = int(np.floor((1.0 / 10.0) / (1.0 / 240.0)))
        self.termination_pos_dist = 0.025
        self.traj_n, self.traj_spacing, self.traj_max_perturb = 10, 0.025, 0.1
        self.np_random = gym_np_random(seed)
        self.steps = 0
        R = np.zeros(9); lib().or_mat_from_quat(_dptr(quat_from_euler(self.workframe_rpy)), _dptr(R)); self.Rw = R.reshape(3, 3)

    def seed(self, seed):
        self.np_random = gym_np_random(seed)

    def work_to_world(self, pos, rpy):   # base_robot_arm.py:47-60
        po, qo = np.zeros(3), np.zeros(4)
        lib().or_mul_transforms(_dptr(np.ascontiguousarray(self.workframe_pos)), _dptr(quat_from_euler(self.workframe_rpy)),
                                _dptr(np.ascontiguousarray(pos, dtype=np.float64)), _dptr(quat_from_euler(rpy)), _dptr(po), _dptr(qo))
        return po, euler_from_quat(qo)

    def update_trajectory(self, third):   # :255-320
        n = self.traj_n
        self.traj_pos_work = np.zeros((n, 3)); self.traj_rpy_work = np.zeros((n, 3))
        init_offset = self.obj_w / 2 + self.traj_spacing
        if self.traj_type == "simplex":
            first = None
            for i in range(n):
                noise = opensimplex_noise2(int(third), i * 0.1, 1) * self.traj_max_perturb
                if first is None:
                    first = -noise
                self.traj_pos_work[i] = [init_offset + i * self.traj_spacing, first + noise, 0.0]
        else:
            for i in range(n):
                dist = i * self.traj_spacing
                self.traj_pos_work[i] = [init_offset + dist * np.cos(third), dist * np.sin(third), 0.0]
        self.traj_rpy_work[:, 2] = np.gradient(self.traj_pos_work[:, 1], self.traj_spacing)
        self.traj_pos_world = np.zeros((n, 3)); self.traj_orn_world = np.zeros((n, 4))
        for i in range(n):
            pw, rw = self.work_to_world(self.traj_pos_work[i], self.traj_rpy_work[i])
            self.traj_pos_world[i] = pw; self.traj_orn_world[i] = quat_from_euler(rw)

    def update_goal(self):   # :335-367
        self.targ += 1
        if self.targ >= self.traj_n:
            return False
        self.goal_pos_world, self.goal_orn_world = self.traj_pos_world[self.targ], self.traj_orn_world[self.targ]
        self.goal_pos_work, self.goal_rpy_work = self.traj_pos_work[self.targ], self.traj_rpy_work[self.targ]
        return True

    def reset(self, draws=None):
        self.steps = 0
        d = push_draws(self.np_random, self.rand_init_orn, self.rand_obj_mass, self.traj_type) if draws is None else np.asarray(draws, dtype=np.float64)
        ang, mass, third = d
        pos = np.zeros(3); rpy = np.zeros(3)
        self.last_reset_substeps = lib().or_robot_reset(C.byref(self.m), C.byref(self.s), _dptr(self.rest), _dptr(pos), _dptr(rpy))
        o = self.o
        o.enabled = 1; o.mass = mass; o.p2p_enabled = 0; o.ext_pending = 0
        q0 = quat_from_euler([-np.pi, 0.0, np.pi / 2 + ang])                              # :210
        # [EXT] changeDynamics(mass=) recomputes the box inertia from the collision shape
        for c in range(3):
            o.inertia[c] = self.cube["inertia_diag"][c] * (mass / self.cube["mass"]); o.com_off[c] = 0.0
            o.pos[c] = self.init_obj_pos[c]; o.vel[c] = 0; o.omg[c] = 0
        for c in range(4):
            o.quat[c] = q0[c]
        self.p.ws_n = 0            # resetBasePositionAndOrientation: the manifolds start empty
        self.update_trajectory(third)
        self.targ = -1
        self.update_goal()
        self.reward, self.done = self.step_data()
        return self.observation()

    def tcp_world(self):
        P, Q = link_states(self.m, np.array(self.s.q[: self.m.ndof]))
        return P[self.m.tcp_link], Q[self.m.tcp_link]

    def stimulus_world(self):
        R = np.zeros(9); q = np.array(self.o.quat[:])
        lib().or_mat_from_quat(_dptr(q), _dptr(R)); R = R.reshape(3, 3)
        return self.tris_local @ R.T + np.array(self.o.pos[:])

    def oracle_obs(self):   # get_oracle_obs :571-609
        pos, rpy, _, lin, ang = tcp_state_workframe(self.m, self.s)
        op, orpy, _, ol, oa = object_state_workframe(self.m, self.o)
        return np.hstack([pos, rpy, lin, ang, op, orpy, ol, oa, self.goal_pos_work, self.goal_rpy_work])

    def features(self):   # get_extended_feature_array :611-629
        p, r = tcp_state_workframe(self.m, self.s)[:2]
        return np.concatenate([p, r, self.goal_pos_work, self.goal_rpy_work])

    def observation(self):
        q = np.array(self.s.q[: self.m.ndof])
        return {"tactile": tactile_image(self.m, q, self.S, self.stimulus_world(), self.ref, border_on=True)[..., None],
                "extended_feature": self.features()}

    def step_data(self):   # get_step_data :456-569
        tp, tq = self.tcp_world()
        op, oq = np.array(self.o.pos[:]), np.array(self.o.quat[:])
        pos_dist = np.linalg.norm(op - self.goal_pos_world)
        if self.reward_mode == "sparse":
            reward = 1.0 if pos_dist < self.termination_pos_dist else 0.0
        else:
            orn_dist = np.arccos(np.clip(2 * (np.inner(self.goal_orn_world, oq) ** 2) - 1, -1, 1))
            Ro, Rt = np.zeros(9), np.zeros(9)
            lib().or_mat_from_quat(_dptr(np.ascontiguousarray(oq)), _dptr(Ro)); lib().or_mat_from_quat(_dptr(np.ascontiguousarray(tq)), _dptr(Rt))
            ov, tv = Ro.reshape(3, 3)[:, 0], Rt.reshape(3, 3)[:, 0]
            cos_dist = 1 - np.dot(ov, tv) / (np.linalg.norm(ov) * np.linalg.norm(tv))
            reward = -(pos_dist + orn_dist + cos_dist)
        done = False
        if pos_dist < self.termination_pos_dist:
            if not self.update_goal():
                done = True
        if self.steps >= self.max_steps:
            done = True
        return reward, done

    def encode_scale(self, action):   # encode_actions :369-454, scale_actions base_tactile_env.py:141-164
        a = np.asarray(action, dtype=np.float64)
        enc = np.zeros(6)
        mm = self.movement_mode
        if mm == "y":
            enc[0], enc[1] = 0.25, a[0]
        elif mm == "yRz":
            enc[0], enc[1], enc[5] = 0.25, a[0], a[1]
        elif mm == "xyRz":
            enc[0], enc[1], enc[5] = a[0], a[1], a[2]
        else:
            _, tq = self.tcp_world()
            Rt = np.zeros(9); lib().or_mat_from_quat(_dptr(np.ascontiguousarray(tq)), _dptr(Rt)); Rt = Rt.reshape(3, 3)
            par = self.Rw.T @ (Rt @ np.array([1.0, 0.0, 0.0])); perp = self.Rw.T @ (Rt @ np.array([0.0, -1.0, 0.0]))
            if mm == "TyRz":
                pa, pe = par * 0.25, perp * a[0]
                enc[0] += pe[0] + pa[0]; enc[1] += pe[1] + pa[1]; enc[5] += a[1]
            else:   # TxTyRz
                pa, pe = par * a[0], perp * a[1]
                enc[0] += pe[0] + pa[0]; enc[1] += pe[1] + pa[1]; enc[5] += a[2]
        enc = np.clip(enc, -0.25, 0.25)
        mv, ma = 0.01, 5.0 * (np.pi / 180)
        if self.control_mode == "TCP_position_control":   # object_push_env.py:137-147: m / rad per step
            mv, ma = 0.001, 1 * (np.pi / 180)
        amax = np.array([mv, mv, 0.0, 0.0, 0.0, ma]); amin = -amax
        return (((enc - (-0.25)) * (amax - amin)) / 0.5) + amin

    def step(self, action):
        v = np.ascontiguousarray(self.encode_scale(action), dtype=np.float64)
        self.steps += 1
        if self.control_mode == "TCP_position_control":   # robot.py:156-186, _max_blocking_pos_move_steps = 10 (object_push_env.py:39)
            self.last_move_substeps = lib().or_tcp_position_control_world(C.byref(self.m), C.byref(self.s), C.byref(self.o), C.byref(self.p), _dptr(v), C.c_int(10))
        else:
            lib().or_tcp_velocity_control(C.byref(self.m), C.byref(self.s), _dptr(v))
            for _ in range(self.repeat):
                lib().or_step_sim_push(C.byref(self.m), C.byref(self.s), C.byref(self.o), C.byref(self.p))
        self.reward, self.done = self.step_data()
        return self.observation(), self.reward, self.done, {}


# ---------------------------------------------------------------- object_roll task restatement
def roll_draws(rng, rand_obj_size=False, rand_embed_dist=False, rand_init_obj_pos=False):
    """Random draws of one ObjectRollEnv.reset in the reference's order: reset_task (scaling_factor, embed_dist;
    object_roll_env.py:182-195), reset_object (x, y; :209-216), make_goal (goal_ang, goal_dist; :244-250)."""
    scale = rng.uniform(1.0, 2.0) if rand_obj_size else 1.0
    embed = rng.uniform(0.0015, 0.003) if rand_embed_dist else 0.0015
    dx = rng.uniform(-0.009, 0.009) if rand_init_obj_pos else 0.0
    dy = rng.uniform(-0.009, 0.009) if rand_init_obj_pos else 0.0
    ang = rng.uniform(-np.pi, np.pi)
    dist = rng.uniform(0.0, 0.015) if rand_init_obj_pos else rng.uniform(0.005, 0.015)
    return np.array([scale, embed, dx, dy, ang, dist])


def sphere_tactile_image(m, q, S, centre, radius, refimg, border_on=True):
    q = np.ascontiguousarray(q, dtype=np.float64)
    dep, gray, mask = refimg
    img = np.zeros((S, S), dtype=np.uint8)
    lib().or_tactile_image_sphere(C.byref(m), _dptr(q), C.c_int(S), _dptr(np.ascontiguousarray(centre, dtype=np.float64)), C.c_double(radius),
                                  dep.ctypes.data_as(C.POINTER(C.c_float)), gray.ctypes.data_as(C.POINTER(C.c_float)),
                                  mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(int(border_on)), img.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return img


class ObjectRollOracle:
    """Restates ObjectRollEnv (rl_envs/nonprehensile_manipulation/object_roll/object_roll_env.py) + BaseObjectEnv on top of
    the C oracle: a marble between the table and the flat TacTip.  One env instance.  Robot.reset() runs without the
    marble in the world; the semi-transparent goal indicator is not drawn (as for the other tasks)."""

    def __init__(self, image_size=128, sensor="tactip", max_steps=250, movement_mode="xy", rand_obj_size=False, rand_embed_dist=False,
                 rand_init_obj_pos=False, reward_mode="dense", seed=None, control_mode="TCP_velocity_control"):
        self.S, self.sensor, self.max_steps, self.movement_mode, self.reward_mode = image_size, sensor, max_steps, movement_mode, reward_mode
        self.control_mode = control_mode
        self.rand_obj_size, self.rand_embed_dist, self.rand_init_obj_pos = rand_obj_size, rand_embed_dist, rand_init_obj_pos
        self.typ = "flat"
        self.default_obj_radius = 0.0025
        self.workframe_rpy = np.array([-np.pi, 0.0, np.pi / 2])
        lims = np.zeros((6, 2))
        lims[0], lims[1], lims[2] = (-0.05, 0.05), (-0.05, 0.05), (-0.01, 0.01)      # :74-81
        self.m = load_model("ur5", sensor, self.typ, [0.65, 0.0, 2 * 0.0025 - 0.0015], self.workframe_rpy, lims)
        self.rest = rest_pose("object_roll", "ur5", sensor, self.typ, self.m)
        self.ref = load_refimg(sensor, self.typ, image_size)
        with open(os.path.join(ASSETS, "models", "ur5_flat_%s.json" % sensor)) as f:
            cyl = json.load(f)["tip_collision"]
        self.s, self.o, self.p = OrState(), OrObject(), OrPush()
        p = self.p
        p.shape = 1
        p.table_z = 0.0
        # sphere lateralFriction 10 (:226) x table 1.0, x tip 10 (:62); [EXT] products, clamped at MAX_FRICTION 10
        p.mu_table, p.mu_tip = min(10.0 * 1.0, 10.0), min(10.0 * 10.0, 10.0)
        p.tip_k, p.tip_d = 1.0 / (1.0 / 10.0 + 1.0 / 1e18), 100 + 0.1                # t_s_dynamics :62
        p.erp, p.slop = 0.2, 1e-4
        p.lin_damping, p.ang_damping = 0.04, 0.04
        p.warmstart, p.ws_n = 0.0, 0
        p.tip_link = self.m._names.index(sensor + "_tip_link")
        p.n_hull = 0
        R = np.zeros(9); lib().or_mat_from_quat(_dptr(quat_from_euler(cyl["rpy"])), _dptr(R)); R = R.reshape(3, 3)
        ax = R @ np.array([0.0, 0.0, 1.0])
        for c in range(3):
            p.cyl_pos[c] = cyl["xyz"][c]; p.cyl_axis[c] = ax[c]
        p.cyl_half_len, p.cyl_radius = cyl["length"] / 2, cyl["radius"]
        self.mass = 0.05                                                            # sphere.urdf
        self.repeat = int(np.floor((1.0 / 10.0) / (1.0 / 240.0)))
        self.termination_pos_dist = 0.001
        self.np_random = gym_np_random(seed)
        self.steps = 0

    def seed(self, seed):
        self.np_random = gym_np_random(seed)

    def reset(self, draws=None):
        self.steps = 0
        d = roll_draws(self.np_random, self.rand_obj_size, self.rand_embed_dist, self.rand_init_obj_pos) if draws is None else np.asarray(draws, dtype=np.float64)
        scale, self.embed_dist, dx, dy, ang, dist = d
        self.radius = self.default_obj_radius * scale
        # update_workframe (:197-202)
        self.workframe_pos = np.array([0.65, 0.0, 2 * self.radius - self.embed_dist])
        for c in range(3):
            self.m.workframe_pos[c] = self.workframe_pos[c]
        pos = np.zeros(3); rpy = np.zeros(3)
        self.last_reset_substeps = lib().or_robot_reset(C.byref(self.m), C.byref(self.s), _dptr(self.rest), _dptr(pos), _dptr(rpy))
        o = self.o
        o.enabled = 1; o.mass = self.mass; o.p2p_enabled = 0; o.ext_pending = 0
        init = [0.65 + dx, 0.0 + dy, self.radius]                                    # :209-216
        for c in range(3):
            o.inertia[c] = 0.4 * self.mass * self.radius ** 2                         # [EXT] btSphereShape::calculateLocalInertia
            o.com_off[c] = 0.0; o.pos[c] = init[c]; o.vel[c] = 0; o.omg[c] = 0
        o.quat[0] = o.quat[1] = o.quat[2] = 0.0; o.quat[3] = 1.0
        self.p.radius = self.radius
        self.p.ws_n = 0
        self.goal_pos_tcp = np.array([dist * np.cos(ang), dist * np.sin(ang), 0.0])   # make_goal :244-256
        self.reward, self.done = self.step_data()
        return self.observation()

    def tcp_world(self):
        P, Q = link_states(self.m, np.array(self.s.q[: self.m.ndof]))
        return P[self.m.tcp_link], Q[self.m.tcp_link]

    def update_goal(self):   # :258-286: the goal is fixed in the TCP frame
        tp, tq = self.tcp_world()
        po, qo = np.zeros(3), np.zeros(4)
        lib().or_mul_transforms(_dptr(np.ascontiguousarray(tp)), _dptr(np.ascontiguousarray(tq)), _dptr(np.ascontiguousarray(self.goal_pos_tcp)),
                                _dptr(np.array([0.0, 0.0, 0.0, 1.0])), _dptr(po), _dptr(qo))
        self.goal_pos_world = po

    def step_data(self):     # get_step_data :299-322, termination :324-336, rewards :338-360
        self.update_goal()
        d = np.linalg.norm(np.array(self.o.pos[:2]) - self.goal_pos_world[:2])
        done = bool(d < self.termination_pos_dist or self.steps >= self.max_steps)
        reward = (1.0 if d < self.termination_pos_dist else 0.0) if self.reward_mode == "sparse" else -d
        return reward, done

    def oracle_obs(self):    # get_oracle_obs :371-409
        pos, _, orn, lin, ang = tcp_state_workframe(self.m, self.s)
        op, _, oo, ol, oa = object_state_workframe(self.m, self.o)
        return np.hstack([pos, orn, lin, ang, op, oo, ol, oa, self.goal_pos_tcp, [0.0, 0.0, 0.0, 1.0], self.radius])

    def features(self):      # get_extended_feature_array :402-408
        return self.goal_pos_tcp.copy()

    def observation(self):
        q = np.array(self.s.q[: self.m.ndof])
        return {"tactile": sphere_tactile_image(self.m, q, self.S, np.array(self.o.pos[:]), self.radius, self.ref)[..., None],
                "extended_feature": self.features()}

    def encode_scale(self, action):
        enc = np.zeros(6); a = np.asarray(action, dtype=np.float64)
        enc[0], enc[1] = a[0], a[1]
        enc = np.clip(enc, -0.25, 0.25)
        mv = 0.001 if self.control_mode == "TCP_position_control" else 0.01   # object_roll_env.py:112-139
        amax = np.array([mv, mv, 0.0, 0.0, 0.0, 0.0]); amin = -amax
        return (((enc - (-0.25)) * (amax - amin)) / 0.5) + amin

    def step(self, action):
        v = np.ascontiguousarray(self.encode_scale(action), dtype=np.float64)
        self.steps += 1
        if self.control_mode == "TCP_position_control":   # robot.py:156-186, _max_blocking_pos_move_steps = 10 (object_roll_env.py:37)
            self.last_move_substeps = lib().or_tcp_position_control_world(C.byref(self.m), C.byref(self.s), C.byref(self.o), C.byref(self.p), _dptr(v), C.c_int(10))
        else:
            lib().or_tcp_velocity_control(C.byref(self.m), C.byref(self.s), _dptr(v))
            for _ in range(self.repeat):
                lib().or_step_sim_push(C.byref(self.m), C.byref(self.s), C.byref(self.o), C.byref(self.p))
        self.reward, self.done = self.step_data()
        return self.observation(), self.reward, self.done, {}

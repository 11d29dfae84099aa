"""Live-PyBullet probe: the oracle (oracle/, test infrastructure) against the REAL pybullet calls the reference makes on the hot
path, and against the unmodified reference env.  SURVEY.md 8(c) / VERDICT r1: every dynamics constant of the oracle tagged [EXT]
(AABB-box inertias, per-link damping, DLS IK, PGS exit) is a recollection of Bullet3 that only a live pybullet can confirm.

pybullet is NOT installed in this image or on the GPU pool and there is no wheel in /opt/wheelhouse, so these tests SKIP here;
they have therefore never executed (parity of the dynamics half stays "unpinned").  They exist so that a
box which does have the wheel reports the comparison instead of silently staying on the port.  Tolerances are north_star's:
pose / reward 1e-3, tactile image L-inf <= 2.
Call sites compared: getLinkState robots/arms/base_robot_arm.py:140, calculateJacobian :300, calculateInverseDynamics :176,
calculateInverseKinematics :201, stepSimulation robots/arms/robot.py:141, getCameraImage sensors/tactile_sensor.py:239-246.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import live_reference as LR  # noqa: E402

_tg, _pb, _why = LR.probe()
pytestmark = pytest.mark.skipif(_tg is None, reason="live reference unavailable: %s" % _why)

EDGE = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}


@pytest.fixture()
def arm(oracle):
    """the UR5 + TacTip URDF loaded the way Robot.load_robot does (robots/arms/robot.py:95-112), at the edge_follow rest pose"""
    pb = _pb
    cid = pb.connect(pb.DIRECT)
    pb.setGravity(0, 0, -9.81, physicsClientId=cid)
    pb.setPhysicsEngineParameter(fixedTimeStep=1.0 / 240, numSolverIterations=150, enableConeFriction=1, contactBreakingThreshold=1e-4,
                                 physicsClientId=cid)                                      # base_tactile_env.py:127-130
    urdf = LR.asset_path("robot_assets", "ur5", "tactip", "ur5_with_standard_tactip.urdf")
    rid = pb.loadURDF(urdf, [0, 0, 0], [0, 0, 0, 1], useFixedBase=True, physicsClientId=cid)
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rest = oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m)
    nj = pb.getNumJoints(rid, physicsClientId=cid)
    ctrl = [j for j in range(nj) if pb.getJointInfo(rid, j, physicsClientId=cid)[2] == pb.JOINT_REVOLUTE]
    assert ctrl == list(m._control_links)[: len(ctrl)] or len(ctrl) == m.ndof
    for j in range(nj):
        pb.changeDynamics(rid, j, linearDamping=0.04, angularDamping=0.04, physicsClientId=cid)   # base_robot_arm.py:24-25
        pb.changeDynamics(rid, j, jointDamping=0.01, physicsClientId=cid)
    yield pb, cid, rid, m, rest, ctrl
    pb.disconnect(cid)


def _set(pb, cid, rid, ctrl, q, qd=None):
    for k, j in enumerate(ctrl):
        pb.resetJointState(rid, j, q[k], 0.0 if qd is None else qd[k], physicsClientId=cid)


def test_link_state_jacobian_inverse_dynamics(oracle, arm):
    pb, cid, rid, m, rest, ctrl = arm
    rng = np.random.RandomState(0)
    for _ in range(8):
        q = rest + rng.uniform(-0.4, 0.4, m.ndof); qd = rng.uniform(-0.5, 0.5, m.ndof)
        _set(pb, cid, rid, ctrl, q, qd)
        P, Q = oracle.link_states(m, q)
        for link in (m.tcp_link, m.body_link):
            ls = pb.getLinkState(rid, link, physicsClientId=cid)
            assert np.allclose(ls[0], P[link], atol=1e-6) and min(np.abs(np.array(ls[1]) - Q[link]).max(), np.abs(np.array(ls[1]) + Q[link]).max()) < 1e-6
        ls = pb.getLinkState(rid, m.tcp_link, physicsClientId=cid)
        jt, jr = pb.calculateJacobian(rid, m.tcp_link, [0, 0, 0], list(q), [0] * m.ndof, [0] * m.ndof, physicsClientId=cid)
        J = oracle.jacobian(m, q, m.tcp_link)
        assert np.allclose(np.vstack([jt, jr]), J, atol=1e-6)
        tau = pb.calculateInverseDynamics(rid, list(q), list(qd), [0] * m.ndof, physicsClientId=cid)
        assert np.allclose(tau, oracle.inverse_dynamics(m, q, qd), atol=1e-5)       # [EXT] AABB-box inertias are on trial here


def test_step_simulation_with_velocity_motors(oracle, arm):
    """24 x stepSimulation under the motors tcp_velocity_control sets (base_robot_arm.py:325-332) vs or_step_sim"""
    import ctypes as C

    pb, cid, rid, m, rest, ctrl = arm
    rng = np.random.RandomState(1)
    q = rest + rng.uniform(-0.2, 0.2, m.ndof); tv = rng.uniform(-0.05, 0.05, m.ndof)
    _set(pb, cid, rid, ctrl, q)
    s = oracle.OrState()
    for k in range(m.ndof):
        s.q[k] = q[k]; s.qd[k] = 0.0; s.motor_mode[k] = 0; s.target_vel[k] = tv[k]; s.kd[k] = 1.0; s.max_force[k] = 1000.0
    pb.setJointMotorControlArray(rid, ctrl, pb.VELOCITY_CONTROL, targetVelocities=list(tv), velocityGains=[1.0] * m.ndof,
                                 forces=[1000.0] * m.ndof, physicsClientId=cid)
    for _ in range(24):
        js = pb.getJointStates(rid, ctrl, physicsClientId=cid)
        tau = pb.calculateInverseDynamics(rid, [x[0] for x in js], [x[1] for x in js], [0] * m.ndof, physicsClientId=cid)
        pb.setJointMotorControlArray(rid, ctrl, pb.TORQUE_CONTROL, forces=list(tau), physicsClientId=cid)   # Robot.step_sim, robot.py:131-141
        pb.stepSimulation(physicsClientId=cid)
        oracle.lib().or_step_sim(C.byref(m), C.byref(s))
    js = pb.getJointStates(rid, ctrl, physicsClientId=cid)
    assert np.allclose([x[0] for x in js], np.array(s.q[: m.ndof]), atol=1e-5)
    assert np.allclose([x[1] for x in js], np.array(s.qd[: m.ndof]), atol=1e-3)


def test_unmodified_reference_env_against_the_oracle(oracle):
    """north_star's parity statement, literally: same seed, same actions -> image L-inf <= 2, pose and reward within 1e-3"""
    env = LR.make_env("edge_follow-v0", EDGE, [128, 128], 200)
    ref = oracle.EdgeFollowOracle(image_size=128, seed=5)
    env.seed(5)
    o = env.reset(); o2 = ref.reset()
    assert np.abs(o["tactile"].astype(int) - o2.astype(int)).max() <= 2
    rng = np.random.RandomState(5)
    for k in range(30):
        a = rng.uniform(-0.25, 0.25, 2).astype(np.float32)
        o, r, d, _ = env.step(a); o2, r2, d2, _ = ref.step(a)
        assert abs(r - r2) < 1e-3 and bool(d) == bool(d2), (k, r, r2)
        assert np.abs(o["tactile"].astype(int) - o2.astype(int)).max() <= 2, k
        tcp = env.robot.arm.get_current_TCP_pos_vel_worldframe()[0]
        assert np.allclose(tcp, ref.tcp_world()[0], atol=1e-3)
    env.close()

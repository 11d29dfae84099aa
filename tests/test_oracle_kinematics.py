"""CPU: the oracle's kinematics against the reference's own design identities:
every rest_poses entry puts the TCP (tcp_link INERTIAL frame) at the env's workframe origin with zero rpy."""
import numpy as np
import pytest

CASES = [
    # env, arm, sensor, type, workframe pos, workframe rpy, expected TCP pos in workframe, tol
    ("edge_follow", "ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], [0, 0, 0.00371], 2e-5),
    ("edge_follow", "ur5", "digit", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], [0, 1e-5, 0.00218], 2e-5),
    ("edge_follow", "ur5", "digitac", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], [0, 0, 0.00146], 2e-5),
    ("object_balance", "ur5", "tactip", "standard", [0.55, 0, 0.35], [0, 0, 0], [0, 0, 0], 2e-5),
    ("surface_follow", "ur5", "digit", "standard", [0.65, 0, 0.025], [-np.pi, 0, np.pi / 2], [0, -7e-4, -0.035], 1e-4),
]


@pytest.mark.parametrize("env,arm,sensor,typ,wpos,wrpy,expect,tol", CASES)
def test_rest_pose_puts_tcp_at_workframe_origin(oracle, env, arm, sensor, typ, wpos, wrpy, expect, tol):
    m = oracle.load_model(arm, sensor, typ, wpos, wrpy, np.zeros((6, 2)))
    q = oracle.rest_pose(env, arm, sensor, typ, m)
    pos, rpy = oracle.tcp_pose_workframe(m, q)
    assert np.allclose(pos, expect, atol=tol), pos
    assert np.allclose(rpy, 0, atol=1e-4), rpy


def test_link_frame_would_be_wrong(oracle):
    """getLinkState()[0:2] is the INERTIAL frame: with the URDF link frame the TCP yaw is off by 1.57."""
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    q = oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m)
    _, Q = oracle.link_states(m, q)
    _, R = oracle.link_frames(m, q)
    Rc = np.zeros(9)
    oracle.lib().or_mat_from_quat(oracle._dptr(np.ascontiguousarray(Q[m.tcp_link])), oracle._dptr(Rc))
    rel = R[m.tcp_link].T @ Rc.reshape(3, 3)
    yaw = np.arctan2(rel[1, 0], rel[0, 0])
    assert abs(yaw - 1.57) < 1e-9


def test_jacobian_matches_finite_differences(oracle):
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rng = np.random.RandomState(0)
    q = oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m) + rng.uniform(-0.3, 0.3, 6)
    J = oracle.jacobian(m, q, m.tcp_link)
    P0, _ = oracle.link_states(m, q)
    for i in range(6):
        dq = q.copy(); dq[i] += 1e-6
        P1, _ = oracle.link_states(m, dq)
        assert np.allclose((P1[m.tcp_link] - P0[m.tcp_link]) / 1e-6, J[:3, i], atol=1e-5)


def test_euler_quat_roundtrip(oracle):
    rng = np.random.RandomState(1)
    for _ in range(100):
        rpy = rng.uniform([-np.pi, -1.5, -np.pi], [np.pi, 1.5, np.pi])
        assert np.allclose(oracle.euler_from_quat(oracle.quat_from_euler(rpy)), rpy, atol=1e-9)


def test_inverse_kinematics_reaches_the_target(oracle):
    """or_inverse_kinematics (the damped-least-squares restatement of pb.calculateInverseKinematics, base_robot_arm.py:201-209):
    from the rest pose, targets a few centimetres / degrees away are met by forward kinematics to the solver's residual, and a
    target at the current pose leaves the joints where they are."""
    import ctypes as C

    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rest = np.array(oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m)[:6])
    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
    P0, Q0 = oracle.link_states(m, rest)
    rng = np.random.RandomState(8)
    for k in range(5):
        dpos = rng.uniform(-0.03, 0.03, 3) if k else np.zeros(3)
        drpy = rng.uniform(-0.2, 0.2, 3) if k else np.zeros(3)
        tpos = P0[m.tcp_link] + dpos
        _, tq = oracle.mul_transforms(np.zeros(3), oracle.quat_from_euler(drpy), np.zeros(3), Q0[m.tcp_link])
        q = np.zeros(8)
        oracle.lib().or_inverse_kinematics(C.byref(m), dp(rest), dp(tpos), dp(tq), q.ctypes.data_as(C.POINTER(C.c_double)))
        P, Q = oracle.link_states(m, q[:6])
        # the damping (0.5 on the diagonal of J^T J) makes the last approach slow: 100 iterations leave ~1e-5 m / rad
        assert np.linalg.norm(P[m.tcp_link] - tpos) < 5e-5, (k, np.linalg.norm(P[m.tcp_link] - tpos))
        assert 1.0 - abs(np.dot(Q[m.tcp_link], tq)) < 1e-8, k
        if k == 0:
            assert np.allclose(q[:6], rest, atol=1e-12)

"""CPU: self-consistency of the oracle's dynamics (the only pybullet-free checks there are):
RNEA(q, qd, ABA(q, qd, tau)) == tau, M^-1 symmetric positive definite, motor-only step reaches the target."""
import ctypes as C

import numpy as np
import pytest


def _model(oracle):
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    return m, oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m)


def test_rnea_aba_identity(oracle):
    m, rest = _model(oracle)
    rng = np.random.RandomState(0)
    for _ in range(20):
        q = rest + rng.uniform(-0.5, 0.5, 6); qd = rng.uniform(-1, 1, 6); tau = rng.uniform(-5, 5, 6)
        qdd = oracle.forward_dynamics(m, q, qd, tau)
        assert np.allclose(oracle.inverse_dynamics(m, q, qd, qdd), tau, atol=1e-10)


def test_minv_spd_and_consistent_with_rnea(oracle):
    m, rest = _model(oracle)
    Minv = oracle.mass_matrix_inverse(m, rest)
    assert np.allclose(Minv, Minv.T, atol=1e-12)
    assert np.linalg.eigvalsh(Minv).min() > 0
    # column j of M = RNEA(q, 0, e_j) - RNEA(q, 0, 0)
    g = oracle.inverse_dynamics(m, rest, np.zeros(6), np.zeros(6))
    M = np.stack([oracle.inverse_dynamics(m, rest, np.zeros(6), np.eye(6)[j]) - g for j in range(6)], axis=1)
    assert np.allclose(M @ Minv, np.eye(6), atol=1e-9)


def test_velocity_motor_reaches_target(oracle):
    m, rest = _model(oracle)
    s = oracle.OrState()
    tv = np.array([0.02, -0.01, 0.015, 0.01, -0.02, 0.03])
    for i in range(6):
        s.q[i] = rest[i]; s.qd[i] = 0; s.motor_mode[i] = 0; s.target_vel[i] = tv[i]; s.kd[i] = 1.0; s.max_force[i] = 1000.0
    oracle.lib().or_step_sim(C.byref(m), C.byref(s))
    # the PGS sweep stops at pybullet's solverResidualThreshold (1e-7 on the squared velocity change of a row),
    # so one substep lands within sqrt(1e-7) ~ 3.2e-4 rad/s of the target and later substeps close the gap
    assert np.abs(np.array(s.qd[:6]) - tv).max() < 3.2e-4
    for _ in range(10):
        oracle.lib().or_step_sim(C.byref(m), C.byref(s))
    assert np.abs(np.array(s.qd[:6]) - tv).max() < 1e-5
    m.solver_residual_threshold = 0.0   # all 150 sweeps: converges to the target exactly
    oracle.lib().or_step_sim(C.byref(m), C.byref(s))
    assert np.allclose(np.array(s.qd[:6]), tv, atol=1e-12)


def test_gym_seeding_is_deterministic(oracle):
    a = oracle.gym_np_random(7).uniform(size=3)
    b = oracle.gym_np_random(7).uniform(size=3)
    c = oracle.gym_np_random(8).uniform(size=3)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_opensimplex_known_answer_and_surface_env(oracle):
    """OpenSimplex restatement: the opensimplex package README's known answer (seed 1234: noise2(10, 10) =
    0.580279369186297), permutation validity, and the surface_follow oracle env's basic invariants."""
    import ctypes as C

    lib = oracle.lib()
    lib.or_opensimplex_noise2.restype = C.c_double
    perm = (C.c_short * 256)()
    lib.or_opensimplex_init(C.c_longlong(1234), perm)
    assert sorted(perm) == list(range(256))
    assert lib.or_opensimplex_noise2(perm, C.c_double(10.0), C.c_double(10.0)) == 0.580279369186297
    h = oracle.surface_heights(1234)
    assert h.shape == (64, 64) and np.abs(h).max() <= 0.025 and np.abs(np.diff(h, axis=0)).max() < 0.004   # coherent noise
    e = oracle.SurfaceFollowOracle(image_size=64, sensor="digit", seed=5)
    e.reset()
    p, _ = e.tcp_world()
    target_z = e.surface_pos[2] + e.h[32, 32] - e.embed_dist
    assert abs(p[2] - target_z) < 3e-4 and abs(p[0] - 0.65) < 2e-4 and abs(p[1]) < 2e-4      # update_init_pose reached
    assert abs(np.linalg.norm(e.goal_pos[:2] - e.surface_pos[:2]) - 0.15) < 1e-12
    _, r, d, _ = e.step(np.array([0.25, 0.0, 0.0]))
    assert r < 0 and not d


def test_free_body_known_answers(oracle):
    """SURVEY 8(c) self-consistency checks for the free rigid body of object_balance (or_step_sim_obj with the point-to-point
    rows switched off): semi-implicit free fall z_n = z_0 + g dt^2 n (n + 1) / 2 with the episode's gravity, a spin about a
    principal axis stays what it is, and a torque-free tumble keeps its angular momentum R I R^T w"""
    import ctypes as C

    b = oracle.ObjectBalanceOracle(image_size=64, rand_gravity=False, rand_embed_dist=False)
    b.reset(draws=np.array([-0.5, 0.0035, 0.0, 0.0]))
    o = b.o
    o.p2p_enabled, o.ext_pending = 0, 0
    dt, g = 1.0 / 240.0, -0.5
    z0 = o.pos[2]
    com0 = np.array(o.pos[:]) + oracle.mat_from_quat(np.array(o.quat[:])) @ np.array(o.com_off[:])
    n = 48
    for _ in range(n):
        oracle.lib().or_step_sim_obj(C.byref(b.m), C.byref(b.s), C.byref(o))
    assert abs((o.pos[2] - z0) - g * dt * dt * n * (n + 1) / 2) < 1e-12
    assert abs(o.vel[2] - g * dt * n) < 1e-12 and abs(o.vel[0]) < 1e-15 and abs(o.omg[0]) < 1e-15
    com1 = np.array(o.pos[:]) + oracle.mat_from_quat(np.array(o.quat[:])) @ np.array(o.com_off[:])
    assert np.allclose(com1[:2], com0[:2], atol=1e-15)
    # spin about the body's z axis (a principal axis of the pole): constant
    b.m.gravity[2] = 0.0
    R = oracle.mat_from_quat(np.array(o.quat[:]))
    for c in range(3):
        o.vel[c] = 0.0; o.omg[c] = 3.0 * R[c, 2]
    for _ in range(240):
        oracle.lib().or_step_sim_obj(C.byref(b.m), C.byref(b.s), C.byref(o))
    R1 = oracle.mat_from_quat(np.array(o.quat[:]))
    assert np.allclose(np.array(o.omg[:]), 3.0 * R1[:, 2], atol=1e-9) and np.allclose(R1[:, 2], R[:, 2], atol=1e-9)
    # torque-free tumble: angular momentum is conserved to the integrator's order, kinetic energy too
    I = np.array(o.inertia[:])
    w0 = R1 @ np.array([1.0, 0.7, 2.0])
    for c in range(3):
        o.omg[c] = w0[c]
    L0 = R1 @ (I * (R1.T @ w0))
    E0 = 0.5 * np.dot(w0, L0)
    for _ in range(240):
        oracle.lib().or_step_sim_obj(C.byref(b.m), C.byref(b.s), C.byref(o))
    R2, w2 = oracle.mat_from_quat(np.array(o.quat[:])), np.array(o.omg[:])
    L2 = R2 @ (I * (R2.T @ w2))
    assert np.linalg.norm(L2 - L0) < 2e-2 * np.linalg.norm(L0)
    assert abs(0.5 * np.dot(w2, L2) - E0) < 2e-2 * E0
    assert np.linalg.norm(R2 - R1) > 0.5          # it did tumble


@pytest.mark.parametrize("arm,sensor", [("ur5", "tactip"), ("mg400", "digitac")])
def test_mass_matrix_from_link_jacobians(oracle, arm, sensor):
    """A third, independent route to the joint-space inertia: M = sum_links J_v^T m J_v + J_w^T (R I R^T) J_w from the
    (finite-difference-checked) link Jacobians at the inertial frames, against the inverse of the oracle's ABA-built M^-1.
    Ties the dynamics restatement to the kinematics one without going through either dynamics algorithm."""
    typ = "standard"
    m = oracle.load_model(arm, sensor, typ, [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rest = np.array(oracle.rest_pose("edge_follow", arm, sensor, typ, m)[: m.ndof])
    rng = np.random.RandomState(4)
    for k in range(3):
        q = rest + rng.uniform(-0.3, 0.3, m.ndof) * (1.0 if arm == "ur5" else 0.2)
        P, Q = oracle.link_states(m, q)
        M = np.zeros((m.ndof, m.ndof))
        for i in range(m.nlinks):
            if m.mass[i] <= 0:
                continue
            J = oracle.jacobian(m, q, i)
            R = oracle.mat_from_quat(Q[i])
            Iw = R @ np.diag(np.array(m.inertia[i][:])) @ R.T
            M += m.mass[i] * J[:3].T @ J[:3] + J[3:].T @ Iw @ J[3:]
        Minv = oracle.mass_matrix_inverse(m, q)
        assert np.allclose(Minv @ M, np.eye(m.ndof), atol=1e-8), np.abs(Minv @ M - np.eye(m.ndof)).max()

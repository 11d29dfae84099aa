"""CPU: the oracle's raster against the reference's golden data - the nodef_dep / border_mask fixtures
(tactile_gym/assets/robot_assets/*/reference_images, loaded at sensors/tactile_sensor.py:63-80).
Known-answer test of SURVEY.md 8(c): z-buffering only the sensor's own visual meshes with the camera model
of tactile_sensor.py:127-229 reproduces the reference's pybullet capture."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [
    ("ur5", "tactip", "standard", 64, 3e-5, 2),
    ("ur5", "tactip", "standard", 128, 3e-5, 4),
    ("ur5", "digit", "standard", 128, 1e-5, 0),
    ("mg400", "digitac", "right_angle", 128, 2e-5, 0),
]


@pytest.mark.parametrize("arm,sensor,typ,S,dtol,mask_tol", CASES)
def test_sensor_render_reproduces_fixture(oracle, arm, sensor, typ, S, dtol, mask_tol):
    m = oracle.load_model(arm, sensor, typ, [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    q = np.zeros(m.ndof)
    meshes = np.load(os.path.join(GOLDEN, "sensor_visual_%s_%s_%s.npz" % (arm, typ, sensor)))
    P, R = oracle.link_frames(m, q)
    e, f, u, r = oracle.camera_frame(m, q)
    dep, gray, mask = oracle.load_refimg(sensor, typ, S)
    body_name = sensor + "_body_link"
    d_body = np.ones((S, S), np.float32)
    d_rest = np.ones((S, S), np.float32)
    for k in meshes.files:
        i = m._names.index(k)
        tris = meshes[k].astype(np.float64) @ R[i].T + P[i]
        d = oracle.depth_image(e, f, u, r, m.fov_deg, m.near_, m.far_, S, tris)
        if k == body_name:
            d_body = np.minimum(d_body, d)
        else:
            d_rest = np.minimum(d_rest, d)
    housing = d_body < d_rest
    skin = (~housing) & (mask == 0)
    assert skin.sum() > 0.5 * S * S
    assert (np.minimum(d_body, d_rest) < 1.0).all()          # sensor meshes cover the whole image
    assert np.abs(d_rest - dep)[skin].max() < dtol           # < 0.16 output LSB
    assert (housing.astype(np.uint8) != mask).sum() <= mask_tol


def test_postprocess_matches_numpy_restatement(oracle):
    """t_s_camera's arithmetic (tactile_sensor.py:268-292) in numpy float32 vs the oracle's C."""
    env = oracle.EdgeFollowOracle(image_size=128, seed=3)
    env.reset()
    q = np.array(env.s.q[:6])
    img, cur = oracle.tactile_image(env.m, q, 128, env.stimulus_world(), env.ref, border_on=True, want_depth=True)
    dep, gray, mask = env.ref
    diff = np.subtract(cur, dep)
    eps = 1e-4
    diff[(diff >= -eps) & (diff <= eps)] = 0
    pen = np.abs(diff)
    pen = ((np.clip(pen, 0, 0.05) / 0.05) * 255).astype(np.uint8)
    pen[mask == 1] = gray[mask == 1]
    assert np.array_equal(pen, img)
    assert (img[mask == 0] > 0).sum() > 100    # the edge is visible


def test_edge_signature(oracle):
    """SURVEY 8(c) mechanism check: box at yaw 0 under the rest pose gives a corner-shaped patch in one quadrant."""
    env = oracle.EdgeFollowOracle(image_size=128, seed=0)
    env.edge_ang, env.embed_dist = 0.0, 0.0035
    for i in range(6):
        env.s.q[i] = env.rest[i]
    img = env.observation()[..., 0]
    dep, gray, mask = env.ref
    rows, cols = np.nonzero((img > 0) & (mask == 0))
    assert rows.min() >= 60 and cols.min() >= 60
    assert 40 < img[mask == 0].max() < 90
    assert np.array_equal(img[mask == 1], gray[mask == 1].astype(np.uint8))


@pytest.mark.parametrize("border_on", [True, False])
def test_body_mask_zeroing_is_an_identity_for_the_composited_depth(oracle, border_on):
    """t_s_camera zeroes every pixel whose SEGMENTATION id is the sensor body (`pen_img[full_mask] = 0`,
    tactile_sensor.py:283-288) before the border is pasted.  The oracle and the kernels have no segmentation image: they
    composite the stimulus into nodef_dep, so wherever the housing is the nearest surface the current depth IS nodef_dep and
    the pixel is already 0; wherever the stimulus is in front of the housing the segmentation id is the stimulus' and the
    reference keeps the value too.  Checked here with a segmentation built from the sensor's own meshes (the KAT's z-buffer)
    plus the stimulus: applying the reference's line to the oracle's image changes nothing, border on or off - including the
    ~2 pixels where `border_mask` and the housing silhouette disagree at 128 x 128."""
    S = 128
    env = oracle.EdgeFollowOracle(image_size=S, seed=3)
    env.reset()
    q = np.array(env.s.q[:6])
    m = env.m
    meshes = np.load(os.path.join(GOLDEN, "sensor_visual_ur5_standard_tactip.npz"))
    P, R = oracle.link_frames(m, q)
    e, f, u, r = oracle.camera_frame(m, q)
    d_body = np.ones((S, S), np.float32)
    d_rest = np.ones((S, S), np.float32)
    for k in meshes.files:
        i = m._names.index(k)
        d = oracle.depth_image(e, f, u, r, m.fov_deg, m.near_, m.far_, S, meshes[k].astype(np.float64) @ R[i].T + P[i])
        if k == "tactip_body_link":
            d_body = np.minimum(d_body, d)
        else:
            d_rest = np.minimum(d_rest, d)
    d_stim = oracle.depth_image(e, f, u, r, m.fov_deg, m.near_, m.far_, S, env.stimulus_world())
    full_mask = (d_body < d_rest) & (d_body <= d_stim)          # the body link wins the depth test = its segmentation id
    img = oracle.tactile_image(m, q, S, env.stimulus_world(), env.ref, border_on=border_on)
    dep, gray, mask = env.ref
    ref_img = img.copy()
    # the reference's order: zero the body pixels, THEN paste the border
    ref_img[full_mask] = 0
    if border_on:
        ref_img[mask == 1] = gray[mask == 1]
    assert full_mask.sum() > 0.3 * S * S and (d_stim < d_rest).sum() > 100   # the housing is seen, and so is the edge through the skin
    assert np.array_equal(ref_img, img)

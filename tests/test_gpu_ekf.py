"""N3: LandmarkEstimatorEKF over the landmarks of a merger pass on the GPU -- pslam_landmarks_ekf_update
against the CPU oracle (fp64 filter inside; the fp32 outputs agree to rounding, the inlier decisions exactly) and through
the scenarios of the reference's filter tests (tests/test_stereo_projective_point_ekf.cpp)."""
import numpy as np
import pytest

import ekf_fixtures as F
import oracle_lib as O

pytestmark = pytest.mark.gpu
KF = F.K.astype(np.float32).reshape(9)


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
    yield c
    c.close()


def scene(n, seed, kind):
    rng = np.random.default_rng(seed)
    truth = np.stack([rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(4, 40, n)], 1)
    cam = np.concatenate([F.rot(1, 0.03) @ F.rot(0, -0.02), np.array([[0.2], [-0.05], [0.4]])], 1)  # sensor_in_world
    M = np.concatenate([F.rot(2, 0.4), np.array([[1.0], [-2.0], [0.5]])], 1)
    sil = np.concatenate([M[:, :3] @ cam[:, :3], (M[:, :3] @ cam[:, 3] + M[:, 3]).reshape(3, 1)], 1)
    pc = (truth - cam[:, 3]) @ cam[:, :3]
    uv = np.stack([F.project(p) for p in pc])
    uvr = np.stack([F.project(p, F.BASELINE) for p in pc])
    meas = {"projective": uv, "projective_depth": np.concatenate([uv, pc[:, 2:3]], 1), "stereo": np.concatenate([uv, uvr], 1)}[kind]
    meas = meas + rng.normal(0, 0.5, meas.shape)
    state = (truth + rng.normal(0, 0.2, (n, 3))).astype(np.float32)
    A = rng.normal(0, 0.2, (n, 3, 3))
    cov = (np.eye(3) * rng.uniform(0.001, 0.5, (n, 1, 1)) + A @ A.transpose(0, 2, 1) * 0.1).astype(np.float32)
    return state, cov.reshape(n, 9), meas.astype(np.float32), cam, sil


@pytest.mark.parametrize("kind", ["projective", "projective_depth", "stereo"])
@pytest.mark.parametrize("n", [1, 33, 1000, 40000])
def test_against_oracle(ctx, kind, n):
    from srrg2_proslam_b200 import capi
    state, cov, meas, cam, sil = scene(n, n + len(kind), kind)
    kw = dict(min_cov=0.01, max_cov_norm2=2.0, max_dist2=0.5)
    g = ctx.landmarks_ekf_update(kind, capi.ekf_cfg(kind, KF, F.BASELINE[:2], cam, sil, **kw), state, cov, meas)
    o = O.landmarks_ekf_update(kind, KF, F.BASELINE[:2], cam, sil, state, cov, meas, **kw)
    assert np.array_equal(g[3], o[3])                      # isInlier
    if n >= 1000:
        assert 0 < g[3].sum() < n                          # both outcomes are exercised
    for a, b in zip(g[:3], o[:3]):
        assert np.allclose(a, b, rtol=2e-6, atol=1e-6)     # fp64 inside, fp32 out: equal up to the last rounding
    rej = ~g[3]
    assert np.array_equal(g[0][rej], state[rej]) and np.array_equal(g[1].reshape(n, 9)[rej], cov[rej])  # rejected: untouched


def test_chain_converges_like_the_reference_scenario(ctx):
    """many landmarks, a moving stereo camera, 12 frames: the same loop as tests/test_oracle_ekf.py on the GPU"""
    from srrg2_proslam_b200 import capi
    rng = np.random.default_rng(3)
    n = 4096
    truth = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(6, 20, n)], 1)
    state = (truth + rng.normal(0, 0.15, (n, 3))).astype(np.float32)
    cov = np.tile(np.eye(3, dtype=np.float32).reshape(9), (n, 1))
    err0 = np.linalg.norm(state - truth, axis=1).mean()
    for step in range(12):
        cam = np.concatenate([F.rot(1, 0.01 * step), np.array([[0.05 * step], [0.0], [0.1 * step]])], 1)
        pc = (truth - cam[:, 3]) @ cam[:, :3]
        meas = np.stack([np.concatenate([F.project(p), F.project(p, F.BASELINE)]) for p in pc]) + rng.normal(0, 0.3, (n, 4))
        cfg = capi.ekf_cfg("stereo", KF, F.BASELINE[:2], cam, cam, max_cov_norm2=4.0, max_dist2=1.0)
        state, cov3, local, inl = ctx.landmarks_ekf_update("stereo", cfg, state, cov, meas)
        cov = cov3.reshape(n, 9)
        assert inl.mean() > 0.9
    assert np.linalg.norm(state - truth, axis=1).mean() < 0.6 * err0


def test_empty_and_invalid(ctx):
    from srrg2_proslam_b200 import capi
    cfg = capi.ekf_cfg("stereo", KF, F.BASELINE[:2], np.eye(3, 4), np.eye(3, 4))
    g = ctx.landmarks_ekf_update("stereo", cfg, np.zeros((0, 3)), np.zeros((0, 9)), np.zeros((0, 4)))
    assert len(g[3]) == 0
    cfg.kind = 7
    with pytest.raises(capi.PslamError) as e:
        ctx.landmarks_ekf_update("stereo", cfg, np.zeros((4, 3)), np.zeros((4, 9)), np.zeros((4, 4)))
    assert e.value.code == capi.PSLAM_E_INVALID


def test_conf_landmark_estimator_module(oracle):
    """kitti.conf "landmark_estimator_ekf" (LandmarkEstimatorStereoProjectiveEKF3D -> StereoProjectivePointEKF3D
    "point_filter") instantiated by class name with the file's thresholds (0.25 / 25 / 0.01), batched compute"""
    import pathlib
    from srrg2_proslam_b200 import plugin as P
    m = P.Manager(pathlib.Path(__file__).resolve().parent / "golden" / "configurations" / "kitti_hotpath.conf")
    est = m.get("landmark_estimator_ekf")
    assert est.class_name == "LandmarkEstimatorStereoProjectiveEKF3D" and not est.is_generic
    flt = est.link("filter")
    assert flt.class_name == "StereoProjectivePointEKF3D" and flt.name == "point_filter"
    flt.filter_set_camera(KF, (F.BASELINE[0], F.BASELINE[1]))
    state, cov, meas, cam, sil = scene(5000, 17, "stereo")
    est.estimator_set_transforms(cam, sil)
    g = est.estimator_compute_batch(state, cov, meas)
    o = O.landmarks_ekf_update("stereo", KF, F.BASELINE[:2], cam, sil, state, cov, meas,
                               min_cov=est.get("minimum_state_element_covariance"),
                               max_cov_norm2=est.get("maximum_covariance_norm_squared"),
                               max_dist2=est.get("maximum_distance_geometry_meters_squared"))
    assert est.get("maximum_covariance_norm_squared") == 0.25 and est.get("maximum_distance_geometry_meters_squared") == 25
    assert np.array_equal(g[3], o[3]) and g[4] == int(o[3].sum()) and 0 < g[4] < 5000
    for a, b in zip(g[:3], o[:3]):
        assert np.allclose(a, b, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("n", [1, 257, 50000])
def test_weighted_mean_bit_exact(ctx, n):
    rng = np.random.default_rng(n)
    truth = rng.uniform(-5, 5, (n, 3)) + np.array([0, 0, 12.0])
    state = (truth + rng.normal(0, 0.5, (n, 3))).astype(np.float32)
    n_opt = rng.integers(0, 30, n).astype(np.int32)
    cam = np.concatenate([F.rot(1, 0.1) @ F.rot(2, -0.05), np.array([[0.3], [0.1], [-0.2]])], 1)
    M = np.concatenate([F.rot(0, 0.2), np.array([[2.0], [0.5], [-1.0]])], 1)
    sil = np.concatenate([M[:, :3] @ cam[:, :3], (M[:, :3] @ cam[:, 3] + M[:, 3]).reshape(3, 1)], 1)
    in_sensor = ((truth - cam[:, 3]) @ cam[:, :3] + rng.normal(0, 0.4, (n, 3))).astype(np.float32)
    g = ctx.landmarks_weighted_mean_update(cam, sil, state, n_opt, in_sensor, max_dist2=0.02)
    o = O.landmarks_weighted_mean_update(cam, sil, state, n_opt, in_sensor, max_dist2=0.02)
    for a, b in zip(g, o):
        assert np.array_equal(a, b)  # fp32, bit exact
    if n > 1000:
        assert 0 < g[2].sum() < n


def test_conf_weighted_mean_module(oracle):
    import pathlib
    from srrg2_proslam_b200 import plugin as P
    m = P.Manager(pathlib.Path(__file__).resolve().parent / "golden" / "configurations" / "kitti_hotpath.conf")
    est = m.get("landmark_estimator_weighted_mean")
    assert est.class_name == "LandmarkEstimatorWeightedMean4D3D" and est.get("maximum_distance_geometry_meters_squared") == 100
    rng = np.random.default_rng(9)
    n = 3000
    state = rng.uniform(-5, 5, (n, 3)).astype(np.float32) + np.array([0, 0, 15], np.float32)
    in_sensor = state + rng.normal(0, 3.0, (n, 3)).astype(np.float32)
    n_opt = rng.integers(0, 5, n).astype(np.int32)
    est.estimator_set_transforms(np.eye(3, 4), np.eye(3, 4))
    g = est.estimator_weighted_mean_batch(state, n_opt, in_sensor)
    o = O.landmarks_weighted_mean_update(np.eye(3, 4), np.eye(3, 4), state, n_opt, in_sensor, max_dist2=100.0)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2], o[2]) and g[3] == int(o[2].sum())


def test_smoother_no_noise_world_bit_exact(ctx):
    """LandmarkWorldNoNoise / LandmarkEstimatorPoseBasedSmoother4D3D (tests/test_landmark_estimators.cpp:210-258): the GPU
    runs the whole scenario with the SAME states as the CPU restatement in every frame (fp32, bit exact), and ends within
    the reference's 1 mm of the true positions"""
    calls = []

    def both(*a):
        g = ctx.landmarks_smoother_update(*a)
        o = O.landmarks_smoother_update(*a)
        for x, y in zip(g, o):
            assert np.array_equal(x, y)
        calls.append(int(g[3].sum()))
        return g

    state, truth, n_opt = F.run_smoother_scenario(both)
    assert len(calls) == 9 and np.linalg.norm(state - truth, axis=1).max() < 1e-3 and n_opt.max() >= 3


@pytest.mark.parametrize("n", [1, 300, 20000])
def test_smoother_random_histories(ctx, n):
    """ragged histories (1 .. 12 measurements), noisy observations, outliers that trip the saturated kernel, landmarks
    behind a camera: every branch of the estimator, bit exact"""
    rng = np.random.default_rng(n)
    F_ = 12
    frames = np.array([np.concatenate([F.rot(1, 0.02 * k) @ F.rot(0, 0.01 * k), np.array([[0.15 * k], [0.02 * k], [0.1 * k]])], 1)
                       for k in range(F_)], np.float32)
    truth = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 25, n)], 1)
    lens = rng.integers(1, 13, n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    hf = np.concatenate([np.sort(rng.choice(F_, L, replace=False)) for L in lens]).astype(np.int32)
    owner = np.repeat(np.arange(n), lens)
    T = frames[hf].astype(np.float64)
    pc = np.einsum("kij,kj->ki", T[:, :, :3].transpose(0, 2, 1), truth[owner] - T[:, :, 3])
    pc += rng.normal(0, 0.02, pc.shape)
    bad = rng.random(len(pc)) < 0.05
    pc[bad] += rng.normal(0, 3.0, (int(bad.sum()), 3))
    K = F.K_WORLD.reshape(3, 3).astype(np.float64)
    h = pc @ K.T
    uv = h[:, :2] / h[:, 2:3] + rng.normal(0, 0.4, (len(pc), 2))
    state = (truth + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    n_opt = rng.integers(0, 6, n).astype(np.int32)
    args = (F.K_WORLD, frames, frames[-1], frames[-1], off, hf, uv.astype(np.float32), pc.astype(np.float32), state, n_opt)
    kw = dict(max_dist2=0.3, max_reproj2=25.0)
    g = ctx.landmarks_smoother_update(*args, **kw)
    o = O.landmarks_smoother_update(*args, **kw)
    for x, y in zip(g, o):
        assert np.array_equal(x, y)
    if n > 1000:
        assert 0 < g[3].sum() < n


def test_conf_smoother_module(oracle):
    """kitti.conf "landmark_estimator_smoother" (LandmarkEstimatorPoseBasedSmoother4D3D, the estimator of merger_triangulation)
    by class name with the file's parameters, driven through the no-noise scenario"""
    import pathlib
    from srrg2_proslam_b200 import plugin as P
    m = P.Manager(pathlib.Path(__file__).resolve().parent / "golden" / "configurations" / "kitti_hotpath.conf")
    est = m.get("landmark_estimator_smoother")
    assert est.class_name == "LandmarkEstimatorPoseBasedSmoother4D3D" and not est.is_generic
    assert est.get("maximum_number_of_iterations") == 100 and est.get("maximum_distance_geometry_meters_squared") == 100
    est.smoother_set_camera_matrix(F.K_WORLD)
    kw = dict(max_iterations=int(est.get("maximum_number_of_iterations")), chi2_delta=est.get("convergence_criterion_minimum_chi2_delta"),
              max_reproj2=est.get("maximum_reprojection_error_pixels_squared"),
              min_measurements=int(est.get("minimum_number_of_measurements_for_optimization")),
              max_dist2=est.get("maximum_distance_geometry_meters_squared"))

    def update(K, frames, siw, sil, off, hf, uv, pic, state, n_opt):
        est.estimator_set_transforms(siw, sil)
        g = est.smoother_compute_batch(frames, off, hf, uv, pic, state, n_opt)
        o = O.landmarks_smoother_update(K, frames, siw, sil, off, hf, uv, pic, state, n_opt, **kw)
        for x, y in zip(g, o):
            assert np.array_equal(x, y)
        return g

    state, truth, n_opt = F.run_smoother_scenario(update, n_points=400)
    assert np.linalg.norm(state - truth, axis=1).max() < 1e-3

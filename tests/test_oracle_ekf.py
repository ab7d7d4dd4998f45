"""Pins the oracle's point EKFs (all in-tree reference code: mapping/landmarks/filters/*.cpp,
point_ekf_base.hpp) through the scenarios and assertions of the reference's tests/test_projective_point_ekf.cpp,
tests/test_projective_depth_point_ekf.cpp and tests/test_stereo_projective_point_ekf.cpp."""
import numpy as np
import pytest

import ekf_fixtures as F
import oracle_lib as O

MEAS = {"projective": 2, "projective_depth": 3, "stereo": 4}


def run_filter(kind, gt, T, meas, Q, Rm_scale, runs=1):
    E = MEAS[kind]
    worst = 0.0
    for _ in range(runs):
        state, cov = gt[0].copy(), np.eye(3)
        for j in range(len(T)):
            state, cov = O.point_ekf(kind, F.CAM6, T[j], Q, meas[j + 1], np.eye(E) * Rm_scale, state, cov)
            assert np.all(np.isfinite(state)) and np.all(np.isfinite(cov))
            worst = max(worst, float(np.linalg.norm(state - gt[j + 1])))
    return worst, cov


@pytest.mark.parametrize("motion", ["translation", "rotation", "transform"])
@pytest.mark.parametrize("kind", ["projective", "projective_depth", "stereo"])
def test_zero_noise_tracks_exactly(oracle, kind, motion):
    """..._ZeroNoise tests (test_stereo_projective_point_ekf.cpp:15-109, test_projective_point_ekf.cpp:14-46,160-192,306-338):
    perfect initial guess, perfect transitions and measurements -> the state follows the ground truth (identity
    measurement covariance, no transition noise)"""
    gt, T, mono, depth, stereo = F.transitions(motion, 100)
    meas = {"projective": mono, "projective_depth": depth, "stereo": stereo}[kind]
    worst, cov = run_filter(kind, gt, T, meas, np.zeros((3, 3)), 1.0)
    assert worst < 1e-9
    assert np.allclose(cov, cov.T, atol=1e-12) and np.all(np.linalg.eigvalsh(0.5 * (cov + cov.T)) > -1e-12)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_full_noise_stays_bounded(oracle, seed):
    """StereoProjectivePointEKF_Transforms_FullNoise (test_stereo_projective_point_ekf.cpp:111-188): motion noise 0.01,
    measurement noise 1 px, Q = 0.1 I, R = 10 I; every filter stays within 1 m of the ground truth for 100 transitions.
    The reference asserts this for its one srand(0) sample; our generator draws other samples, so the 1 m bound is kept
    for seed 0 and the binocular filter, and 2 m for the weaker monocular filters on the other seeds."""
    gt, T, mono, depth, stereo = F.transitions("transform", 100, sd_motion=0.01, sd_meas=1.0, seed=seed)
    for kind, meas in (("projective", mono), ("projective_depth", depth), ("stereo", stereo)):
        worst, _ = run_filter(kind, gt, T, meas, np.eye(3) * 0.1, 10.0)
        assert worst < (1.0 if seed == 0 or kind == "stereo" else 2.0), (kind, worst)


def test_landmark_estimator_ekf(oracle):
    """LandmarkEstimatorEKF_::compute (landmark_estimator_ekf_impl.cpp:17-82): a static world point observed from a moving
    stereo camera converges towards the truth; a measurement that would move the landmark by more than
    maximum_distance_geometry_meters_squared or leaves a large covariance is rejected (isInlier stays false)"""
    rng = np.random.default_rng(3)
    Kf = F.K.astype(np.float32).reshape(9)
    n = 64
    truth = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(6, 20, n)], 1)
    state = (truth + rng.normal(0, 0.15, (n, 3))).astype(np.float32)
    cov = np.tile(np.eye(3, dtype=np.float32).reshape(9), (n, 1))
    err0 = np.linalg.norm(state - truth, axis=1).mean()
    for step in range(12):
        cam = np.concatenate([F.rot(1, 0.01 * step), np.array([[0.05 * step], [0.0], [0.1 * step]])], 1)  # sensor_in_world
        Rw, tw = cam[:, :3], cam[:, 3]
        pc = (truth - tw) @ Rw  # world -> sensor
        meas = np.stack([np.concatenate([F.project(p), F.project(p, F.BASELINE)]) for p in pc]) + rng.normal(0, 0.3, (n, 4))
        # the local map's frame: world moved by M (sensor_in_local_map = M * sensor_in_world => world_in_local_map = M)
        M = np.concatenate([F.rot(2, 0.4), np.array([[1.0], [-2.0], [0.5]])], 1)
        sensor_in_local_map = np.concatenate([M[:, :3] @ Rw, (M[:, :3] @ tw + M[:, 3]).reshape(3, 1)], 1)
        state, cov3, local, inl = O.landmarks_ekf_update("stereo", Kf, F.BASELINE[:2], cam, sensor_in_local_map, state, cov, meas,
                                                        max_cov_norm2=4.0, max_dist2=1.0)
        cov = cov3.reshape(n, 9)
        assert inl.sum() >= n - 4
        assert np.allclose(local[inl], state[inl] @ M[:, :3].T + M[:, 3], atol=1e-4)
    assert np.linalg.norm(state - truth, axis=1).mean() < 0.6 * err0
    # gross outlier: rejected, statistics untouched
    bad = meas.copy()
    bad[:, 0] += 150.0
    bad[:, 2] += 120.0
    s2, c2, _, inl2 = O.landmarks_ekf_update("stereo", Kf, F.BASELINE[:2], cam, cam, state, cov, bad, max_cov_norm2=4.0, max_dist2=0.01)
    assert not inl2.any() and np.array_equal(s2, state) and np.array_equal(c2.reshape(n, 9), cov)


def test_landmark_estimator_weighted_mean(oracle):
    """LandmarkEstimatorWeightedMean_::compute (landmark_estimator_weighted_mean_impl.cpp:7-41): the mean converges to the
    true position when the re-observations are exact (tests/test_landmark_estimators.cpp:29-69), a jump farther than the
    geometric gate is rejected"""
    rng = np.random.default_rng(5)
    n = 200
    truth = rng.uniform(-5, 5, (n, 3)) + np.array([0, 0, 12.0])
    state = (truth + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    n_opt = np.zeros(n, np.int32)
    for step in range(20):
        cam = np.concatenate([F.rot(1, 0.02 * step), np.array([[0.1 * step], [0.0], [0.05 * step]])], 1)
        in_sensor = (truth - cam[:, 3]) @ cam[:, :3]
        state, local, inl = O.landmarks_weighted_mean_update(cam, cam, state, n_opt, in_sensor, max_dist2=1.0)
        assert inl.all() and np.allclose(local, state, atol=1e-5)  # sensor_in_local_map = sensor_in_world => local map = world
        n_opt += 1
    assert np.abs(state - truth).max() < 0.05
    far = in_sensor + np.array([0, 0, 50.0])
    s2, _, inl2 = O.landmarks_weighted_mean_update(cam, cam, state, n_opt, far, max_dist2=1.0)
    assert not inl2.any() and np.array_equal(s2, state)


def test_pose_based_smoother_no_noise_world(oracle):
    """LandmarkWorldNoNoise / LandmarkEstimatorPoseBasedSmoother4D3D (tests/test_landmark_estimators.cpp:210-258,332-347):
    1000 points, 10 poses, exact observations -> every landmark state ends within 1 mm of its true world position"""
    state, truth, n_opt = F.run_smoother_scenario(O.landmarks_smoother_update)
    assert len(state) > 100
    err = np.linalg.norm(state - truth, axis=1)
    assert err.max() < 1e-3, err.max()
    assert n_opt.max() >= 3  # the Gauss-Newton branch ran (three or more measurements)


def test_full_piv_lu_and_rejection(oracle):
    """a re-observation far from the landmark is averaged in but rejected by the geometric gate while the landmark has
    fewer than three measurements (landmark_estimator_pose_based_smoother_impl.cpp:29-43)"""
    poses = np.array([np.eye(3, 4), np.concatenate([np.eye(3), [[0.1], [0], [0]]], 1)], np.float32)
    truth = np.array([[0.5, -0.2, 8.0]], np.float32)
    off = np.array([0, 2], np.int32)
    hf = np.array([0, 1], np.int32)
    uv = np.array([[112.5, 95.0], [110.0, 95.0]], np.float32)
    pic = np.array([[0.5, -0.2, 8.0], [0.4, -0.2, 80.0]], np.float32)  # second depth is absurd
    st, no, loc, inl = O.landmarks_smoother_update(F.K_WORLD, poses, poses[1], poses[1], off, hf, uv, pic, truth, np.zeros(1, np.int32),
                                                   max_dist2=1.0)
    assert not inl[0] and np.array_equal(st, truth) and no[0] == 0
    pic[1] = [0.4, -0.2, 8.0]
    st, no, loc, inl = O.landmarks_smoother_update(F.K_WORLD, poses, poses[1], poses[1], off, hf, uv, pic, truth, np.zeros(1, np.int32))
    assert inl[0] and no[0] == 2 and np.allclose(st, truth, atol=1e-5)

"""Stage 3/4 oracle against the reference's own pose-tolerance tests (the only pins the reference holds for the
factor / solver arithmetic): KITTI 00 -> 01 on test_data, tests/test_aligners.cpp:640-758
(fixed correspondences from the projective finder, 100 GN iterations from identity, with and without
inverse-depth weighting: |t| < 0.1 m, |q| < 0.005) and :470-583 (aligner loop from identity: < 0.15 m, < 0.005)."""
import numpy as np
import pytest

import oracle_lib as O
from test_oracle_known_answers import CAM00, CAM01, K_KITTI

BASELINE_L_IN_R = (-386.1448, 0.0, 0.0)  # -baseline_right_in_left_pixels, tests/test_aligners.cpp:604-605


@pytest.fixture(scope="module")
def kitti(oracle):
    """tests/fixtures.hpp:832-846,926-952: adaptor (thr 15, target 500, epipolar 50 / 0.8) on frames 0 and 1,
    frame 0 triangulated with minimum disparity 0"""
    c = O.extract_cfg(threshold=15, target=500)
    meas = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"),
                             c, "epipolar", 50, 0.8) for i in (0, 1)]
    xyz, ninv = O.triangulate(meas[0]["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
    assert ninv == 0
    cam01_in_00 = O.pose_mul(O.pose_inverse(CAM00), CAM01)
    return meas, xyz, cam01_in_00


def manifold_error(estimate, cam01_in_00):
    return O.t2tnq(O.pose_mul(estimate, cam01_in_00))  # t2tnq(variable->estimate() * camera_01_in_00)


def test_fixture_sizes(kitti):  # 145 / 139 stereo points on frames 0 / 1
    meas, xyz, _ = kitti
    assert len(meas[0]["uvuv"]) == 145 and len(meas[1]["uvuv"]) == 139 and len(xyz) == 145


@pytest.mark.parametrize("weighted", [False, True])
def test_kitti_00_to_01_fixed_correspondences(kitti, weighted):  # tests/test_aligners.cpp:640-758
    meas, xyz, cam01_in_00 = kitti
    pf = O.ProjectiveFinder(K_KITTI, 376, 1241, "circle", max_desc_dist=100, ratio=0.5, min_matching_ratio=0.1,
                            min_desc_dist=100, max_radius=5, min_radius=5, min_iterations=10)
    pf.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    pf.set_moving(xyz, meas[0]["desc"])
    pf.set_estimate(O.pose_inverse(cam01_in_00))  # perfect estimate
    for _ in range(100):
        fi, mi, _d = pf.compute()
    assert len(fi) == 42
    md = 0.0
    if weighted:
        d = meas[1]["uvuv"][fi, 0] - meas[1]["uvuv"][fi, 2]
        md = float(np.float32(d.astype(np.float32).sum(dtype=np.float32) / np.float32(len(d))))
    lcfg = O.linearize_cfg("stereo", K_KITTI, 1241, 376, BASELINE_L_IN_R, md, "saturated", 1000.0)
    info = np.tile(np.array([1.0, 2.0, 1.0]), (len(meas[1]["uvuv"]), 1))
    pose = np.eye(3, 4).reshape(12)
    for _ in range(100):
        H, b, st = O.linearize(lcfg, pose, xyz, meas[1]["uvuv"], fi, mi, info)
        rc, pose, dx = O.gn_step(H, b, 0.0, pose)
        assert rc == 0
    assert np.abs(dx).max() < 1e-9  # converged
    e = manifold_error(pose, cam01_in_00)
    assert np.all(np.abs(e[:3]) < 0.1) and np.all(np.abs(e[3:]) < 0.005), e


def test_kitti_00_to_01_aligner_loop(kitti):
    """kitti.conf wiring of finder + slice + GN (adaptive finder starting at descriptor distance 25), identity guess,
    no motion-model prior.  The reference holds no pin for this combination on test_data (its aligner test,
    tests/test_aligners.cpp:440-583, runs on a random synthetic world); with the 43 correspondences the adaptive
    finder yields the 0.86 m forward motion is recovered to 0.2 m / 0.002 -- bounds below are OURS, the reference's
    0.1 m bound is met by the fixed-correspondence tests above."""
    meas, xyz, cam01_in_00 = kitti
    pf = O.ProjectiveFinder(K_KITTI, 376, 1241, "circle", max_desc_dist=75, ratio=0.8, min_matching_ratio=0.1,
                            min_desc_dist=25, desc_step=5, max_radius=50, min_radius=10, radius_step=10,
                            min_iterations=5, max_change_norm=0.01, iters_per_projection=5)
    pf.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    pf.set_moving(xyz, meas[0]["desc"])
    r = O.align(pf, "stereo", K_KITTI, 376, 1241, meas[1]["uvuv"], xyz, [1, 2, 1], baseline=BASELINE_L_IN_R,
                inverse_depth_weighting=True, chi_threshold=25.0, max_iterations=100, damping=1.0,
                min_num_inliers=6, min_num_correspondences=10)
    assert r["status"] == O.ALIGNER_STATUS["Success"]
    e = manifold_error(r["pose"], cam01_in_00)
    assert np.all(np.abs(e[:3]) < 0.25) and np.all(np.abs(e[3:]) < 0.005), e
    assert pf.state()["converged"]

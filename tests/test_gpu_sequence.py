"""The whole hot path chained over the reference's KITTI snippet (test_data/kitti/city frames 0..4, the frames of
tests/test_trackers.cpp:260-365 "KITTI 00To04_Tracker_ProjectiveBF_NoMerges"): per frame
  stereo adaptor (kitti.conf extractor + epipolar finder) -> rigid-stereo triangulation of the previous frame ->
  conf-driven aligner (projective circle finder + stereo factor + motion-model slice + GN, 100 iterations) -> pose chain.
The aligner's second slice (AlignerSliceMotionModel3D, configurations/kitti.conf:747-772) gets the trajectory chunk a
tracker would hand it: the previous frame-to-frame motion, so that the constant-velocity prediction seeds every
alignment and its pose-prior factor is summed into H, b inside the fused launch.
The tracker / merger / clipper control plane is NOT part of the path; the test chains
frame-to-frame alignments itself, once through the CUDA-backed plugin modules and once through the CPU oracle, and
checks the north_star's pose criteria: every per-frame pose within 1e-6 m / 1e-6 rad of the CPU path, trajectory
error against the KITTI ground truth (tests/fixtures.hpp:884-911) unchanged within 0.1 %, and the final 00 -> 04
error inside the reference tracker test's own tolerances (tests/test_trackers.cpp:359-364)."""
import pathlib

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_known_answers import K_KITTI, kitti_pose

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
BASELINE_M = 0.537166  # tests/fixtures.hpp:811-816

# tests/fixtures.hpp:884-905 (camera_0k_in_world, copied from KITTI 00 gt.txt)
CAM_IN_WORLD = [
    kitti_pose([1.0, 9.043680e-12, 2.326809e-11, 9.043683e-12, 1.0, 2.392370e-10, 2.326810e-11, 2.392370e-10,
                9.999999e-01], [5.551115e-17, 3.330669e-16, -4.440892e-16]),
    kitti_pose([9.999978e-01, 5.272628e-04, -2.066935e-03, -5.296506e-04, 9.999992e-01, -1.154865e-03,
                2.066324e-03, 1.155958e-03, 9.999971e-01], [-4.690294e-02, -2.839928e-02, 8.586941e-01]),
    kitti_pose([9.999910e-01, 1.048972e-03, -4.131348e-03, -1.058514e-03, 9.999968e-01, -2.308104e-03,
                4.128913e-03, 2.312456e-03, 9.999887e-01], [-9.374345e-02, -5.676064e-02, 1.716275e+00]),
    kitti_pose([9.999796e-01, 1.566466e-03, -6.198571e-03, -1.587952e-03, 9.999927e-01, -3.462706e-03,
                6.193102e-03, 3.472479e-03, 9.999747e-01], [-1.406429e-01, -8.515762e-02, 2.574964e+00]),
    kitti_pose([9.999637e-01, 2.078471e-03, -8.263498e-03, -2.116664e-03, 9.999871e-01, -4.615826e-03,
                8.253797e-03, 4.633149e-03, 9.999551e-01], [-1.874858e-01, -1.135202e-01, 3.432648e+00]),
]


def kitti_pair(i):
    return O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png")


def ate_rmse(cams_in_00, gt_in_00):
    """absolute trajectory error: RMSE of the camera positions against the ground truth (both start at identity)"""
    d = [np.asarray(a, np.float64).reshape(3, 4)[:, 3] - np.asarray(b, np.float64).reshape(3, 4)[:, 3]
         for a, b in zip(cams_in_00, gt_in_00)]
    return float(np.sqrt(np.mean([np.dot(x, x) for x in d])))


def test_kitti_00_to_04_odometry(oracle):
    from srrg2_proslam_b200 import plugin as P
    m = P.Manager(GOLDEN / "configurations" / "kitti_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    pr = sl.link("projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    ad = m.get("adaptor_stereo_projective")
    al.aligner_set_left_camera_in_right([-BASELINE_M, 0, 0])
    b_x = np.float32(718.856) * np.float32(BASELINE_M)
    base = (K_KITTI.reshape(3, 3) @ np.array([-BASELINE_M, 0, 0], np.float32)).astype(np.float32)
    ecfg = O.extract_cfg(threshold=15, target=1000)  # kitti.conf extractor (:229-255), epipolar finder (:484-501)

    # ONE finder for the whole sequence, like the plugin module: the adapted search radius / descriptor distance
    # survive from frame to frame (correspondence_finder_projective_base_impl.cpp:113-122, "once per session")
    of = O.ProjectiveFinder(K_KITTI, 376, 1241, "circle", max_desc_dist=75, ratio=0.8, min_matching_ratio=0.1,
                            min_desc_dist=25, desc_step=5, max_radius=50, min_radius=10, radius_step=10,
                            min_iterations=5, max_change_norm=0.01, iters_per_projection=5)
    g_prev = o_prev = None
    chunk = np.eye(3, 4, dtype=np.float32).reshape(1, 12)  # first alignment: one pose, no motion yet
    g_00_in_k = np.eye(3, 4).reshape(12)  # camera 00 expressed in camera k (= moving_in_fixed chained)
    o_00_in_k = np.eye(3, 4).reshape(12)
    g_traj, o_traj = [np.eye(3, 4).reshape(12)], [np.eye(3, 4).reshape(12)]
    for k in range(5):
        L, R = kitti_pair(k)
        g_meas = ad.stereo_adaptor(L, R)
        o_meas = O.stereo_adaptor(L, R, ecfg, "epipolar", 100, 0.5, 100, 0)
        assert len(g_meas["uvuv"]) == len(o_meas["uvuv"]) > 100
        for key in ("uvuv", "intensity", "desc"):
            assert np.array_equal(g_meas[key], o_meas[key]), (k, key)  # stages 1-2: bit exact
        if k > 0:
            xyz, _ = O.triangulate(o_prev["uvuv"], K_KITTI, b_x, 0.0)
            al.aligner_set_fixed(g_meas["uvuv"], g_meas["desc"])
            al.aligner_set_moving(xyz, g_prev["desc"])
            al.aligner_set_moving_in_fixed(np.eye(3, 4, dtype=np.float32))
            # robot poses in the local map (= the previous camera frame), oldest first: camera k-2, then camera k-1 itself
            al.aligner_set_trajectory_chunk(chunk)
            g = al.aligner_compute()
            of.set_fixed(o_meas["uvuv"], o_meas["desc"])
            of.set_moving(xyz, o_prev["desc"])
            pred = O.constant_velocity_prediction(list(chunk))
            o = O.align(of, "stereo", K_KITTI, 376, 1241, o_meas["uvuv"], xyz, [1, 2, 1], baseline=base,
                        inverse_depth_weighting=True, chi_threshold=25.0, max_iterations=100, damping=1.0,
                        min_num_inliers=6, min_num_correspondences=10, init_pose=pred, prior=(pred, np.eye(6)))
            if k > 1:  # the prediction (last motion repeated) starts within a few centimetres of the solution
                assert np.abs(O.t2tnq(O.pose_mul(O.pose_inverse(pred), g["pose"]))[:3]).max() < 0.1
            assert g["status"] == o["status"] == O.ALIGNER_STATUS["Success"], k
            assert np.array_equal(g["stats"][:, :3], o["stats"][:, :3]), k
            assert all(np.array_equal(a, b) for a, b in zip(g["corr"], o["corr"])), k
            d = O.t2tnq(O.pose_mul(O.pose_inverse(o["pose"]), g["pose"]))
            assert np.abs(d[:3]).max() < 1e-6 and np.abs(d[3:]).max() < 1e-6, (k, d)  # north_star: per-frame pose
            g_00_in_k = O.pose_mul(g["pose"], g_00_in_k)
            o_00_in_k = O.pose_mul(o["pose"], o_00_in_k)
            g_traj.append(O.pose_inverse(g_00_in_k))
            o_traj.append(O.pose_inverse(o_00_in_k))
            chunk = np.stack([g["pose"], np.eye(3, 4).reshape(12)]).astype(np.float32)
        g_prev, o_prev = g_meas, o_meas

    w_in_00 = O.pose_inverse(CAM_IN_WORLD[0])
    gt = [O.pose_mul(w_in_00, c) for c in CAM_IN_WORLD]
    # final error, the reference tracker test's expression and tolerances (tests/test_trackers.cpp:357-364):
    # t2tnq(robotInLocalMap^-1 * camera_04_in_00) with robotInLocalMap = camera 04 in 00
    e = O.t2tnq(O.pose_mul(g_00_in_k, gt[4]))
    assert np.all(np.abs(e[:2]) < 0.2) and abs(e[2]) < 0.7 and np.all(np.abs(e[3:]) < 0.01), e
    ate_g, ate_o = ate_rmse(g_traj, gt), ate_rmse(o_traj, gt)
    assert ate_o > 0 and abs(ate_g - ate_o) / ate_o < 1e-3, (ate_g, ate_o)  # north_star: ATE unchanged within 0.1 %
    print(f"KITTI 00->04 odometry: final error {np.round(e, 4)}, ATE gpu {ate_g:.6f} m, cpu {ate_o:.6f} m")

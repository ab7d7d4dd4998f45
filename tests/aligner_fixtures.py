"""The reference's conf-driven aligner scenarios (tests/test_aligners.cpp:883-1337) rebuilt from the committed images:
fixtures of tests/fixtures.hpp (ICL :555-720, KITTI :800-985) and, per scenario, the parameters the reference test
sets on top of the shipped icl.conf / kitti.conf "aligner".  Shared by the CPU pin of the oracle
(tests/test_oracle_solver.py) and the GPU test through the plugin mirror (tests/test_gpu_aligner_pins.py)."""
import functools

import numpy as np

import oracle_lib as O
from merger_fixtures import icl_measurements, quat_to_R
from scene_fixtures import K_ICL, K_KITTI, unproject

BASELINE_M = 0.537166  # tf right_in_left, tests/fixtures.hpp:813-816


def _pose(R, t):
    return np.concatenate([np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3, 1)], 1).reshape(12)


@functools.lru_cache(maxsize=None)
def icl():
    """tests/fixtures.hpp:560-720: adaptor measurements (u, v, depth) of frames 00 / 50, frame 00 unprojected into its
    camera, the 2-D view of frame 50 and the ground-truth motion camera_50_in_00"""
    m0, m50 = icl_measurements(0), icl_measurements(50)
    cam00 = _pose(quat_to_R(1, 0, 0, 0), [0, 0, -2.25])                                   # :597-599
    cam50 = _pose(quat_to_R(0.995539, -0.00521396, 0.0821083, 0.0461804), [0.129723, 0.00959134, -2.25525])  # :604-608
    return dict(K=K_ICL, rows=480, cols=640, points_00=unproject(m0["uvd"], K_ICL), desc_00=m0["desc"],
                meas_50=m50["uvd"], desc_50=m50["desc"], meas_50_2d=np.ascontiguousarray(m50["uvd"][:, :2]),
                cam_50_in_00=O.pose_mul(O.pose_inverse(cam00), cam50))


# KITTI sequence 00 ground truth, tests/fixtures.hpp:884-908
_KITTI_T = [(5.551115e-17, 3.330669e-16, -4.440892e-16), (-4.690294e-02, -2.839928e-02, 8.586941e-01),
            (-9.374345e-02, -5.676064e-02, 1.716275e+00)]
_KITTI_R = [(1.000000e+00, 9.043680e-12, 2.326809e-11, 9.043683e-12, 1.000000e+00, 2.392370e-10, 2.326810e-11, 2.392370e-10, 9.999999e-01),
            (9.999978e-01, 5.272628e-04, -2.066935e-03, -5.296506e-04, 9.999992e-01, -1.154865e-03, 2.066324e-03, 1.155958e-03, 9.999971e-01),
            (9.999910e-01, 1.048972e-03, -4.131348e-03, -1.058514e-03, 9.999968e-01, -2.308104e-03, 4.128913e-03, 2.312456e-03, 9.999887e-01)]


@functools.lru_cache(maxsize=None)
def kitti():
    """tests/fixtures.hpp:832-980: stereo adaptor (thr 15, target 500, epipolar 50 / 0.8) on frames 0..2, frame 0
    triangulated (minimum disparity 0), ground-truth motions"""
    c = O.extract_cfg(threshold=15, target=500)
    meas = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"),
                             c, "epipolar", 50, 0.8) for i in range(3)]
    xyz, ninv = O.triangulate(meas[0]["uvuv"], K_KITTI, np.float32(718.856) * np.float32(BASELINE_M), 0.0)
    assert ninv == 0
    cam = [_pose(np.reshape(R, (3, 3)), t) for R, t in zip(_KITTI_R, _KITTI_T)]
    w_in_0 = O.pose_inverse(cam[0])
    return dict(K=K_KITTI, rows=376, cols=1241, meas=meas, points_00=xyz, desc_00=meas[0]["desc"],
                cam_in_00=[O.pose_mul(w_in_0, cam[i]) for i in range(3)],
                baseline=(K_KITTI.reshape(3, 3) @ np.array([-BASELINE_M, 0, 0], np.float32)).astype(np.float32))


IDENTITY_PRIOR = (np.eye(3, 4).reshape(12), np.eye(6))  # motion-model slice with an EMPTY trajectory chunk: no motion

# icl.conf "aligner" (#6): 100 iterations, min 6 inliers, inlier-only runs + keep-only-inliers on, solver #9 -> GN damping 0.1;
# slice #7: diag [1, 1, 10], saturated chi 10, min_num_correspondences 0; second slice #8 = the motion model
ICL_ALIGNER = dict(max_iterations=100, min_num_inliers=6, damping=0.1, enable_inlier_only_runs=True,
                   keep_only_inlier_correspondences=True, min_num_correspondences=0)
# kitti.conf "aligner" (#29): 100 iterations, min 6 inliers, no inlier-only runs, GN damping 1; slice #22: diag [1, 2, 1],
# inverse-depth weighting on, min_num_correspondences 10; the tests raise the saturated chi threshold to 1000
KITTI_ALIGNER = dict(max_iterations=100, min_num_inliers=6, damping=1.0, min_num_correspondences=10)

# name -> (dataset, scenario).  bounds = the reference test's |error| limits on t2tnq(movingInFixed * camera_b_in_a)
SCENARIOS = {
    # tests/test_aligners.cpp:883-961: new AlignerSliceProcessorProjective (saturated chi 100^2 by default,
    # aligner_slice_processor_projective.cpp:7-20), cf_bruteforce_2d (30 / 0.7), diag (1, 1), identity guess
    "icl_00to50_projective_bruteforce": dict(
        data="icl", kind="mono", fixed="meas_50_2d", finder=("bruteforce", dict(max_dist=30, ratio=0.7)), diag=[1, 1],
        chi=100.0 * 100.0, init="identity", bounds=[0.01] * 6, aligner=dict(ICL_ALIGNER)),
    # :964-1032: the conf's depth slice (diag set to (1, 1, 10), saturated chi 10), cf_bruteforce_3d (35 / 0.7)
    "icl_00to50_depth_bruteforce": dict(
        data="icl", kind="depth", fixed="meas_50", finder=("bruteforce", dict(max_dist=35, ratio=0.7)), diag=[1, 1, 10],
        chi=10.0, init="identity", bounds=[0.01] * 6, aligner=dict(ICL_ALIGNER)),
    # :1035-1103: the conf's cf_projective_circle (icl.conf:319-360) on the fixture's projector.  The test calls
    # setMovingInFixed(camera_50_in_00) (:1083) -- the INVERSE of the solution, ~22 degrees / 0.26 m away from it, a
    # 160-pixel image shift that a 100-pixel search radius cannot bridge -- and still expects < 0.01: the aligner's second
    # slice (AlignerSliceMotionModel3D, empty trajectory chunk) must therefore seed the estimate with its prediction
    # (no motion).  Restated that way ("seed_from_prior", [upstream, inferred from this test]).
    "icl_00to50_depth_projective_circle": dict(
        data="icl", kind="depth", fixed="meas_50", diag=[1, 1, 10], chi=10.0, init="cam_50_in_00", bounds=[0.01] * 6,
        finder=("circle", dict(max_desc_dist=35, ratio=0.9, min_matching_ratio=0.2, min_desc_dist=30, desc_step=5, max_radius=100,
                               min_radius=25, radius_step=5, min_iterations=5, max_change_norm=0.01, iters_per_projection=5)),
        aligner=dict(ICL_ALIGNER)),
    # :1106-1179: Bruteforce4D3D 100 / 0.5, chi 1000
    "kitti_00to01_bruteforce": dict(
        data="kitti", kind="stereo", frame=1, finder=("bruteforce", dict(max_dist=100, ratio=0.5)), diag=[1, 2, 1], chi=1000.0,
        init="identity", bounds=[0.1, 0.1, 0.2, 0.01, 0.01, 0.01], aligner=dict(KITTI_ALIGNER)),
    # :1182-1260: kitti.conf cf_projective_circle with distance 50 -> 100, ratio 0.8, radius 50 -> 10, re-projection every 5th
    "kitti_00to01_projective_circle": dict(
        data="kitti", kind="stereo", frame=1, diag=[1, 2, 1], chi=1000.0, init="identity",
        bounds=[0.05, 0.05, 0.2, 0.01, 0.01, 0.01],
        finder=("circle", dict(max_desc_dist=100, ratio=0.8, min_matching_ratio=0.1, min_desc_dist=50, desc_step=5, max_radius=50,
                               min_radius=10, radius_step=10, min_iterations=5, max_change_norm=0.01, iters_per_projection=5)),
        aligner=dict(KITTI_ALIGNER)),
    # :1264-1337: frame 00 -> 02
    "kitti_00to02_bruteforce": dict(
        data="kitti", kind="stereo", frame=2, finder=("bruteforce", dict(max_dist=100, ratio=0.5)), diag=[1, 2, 1], chi=1000.0,
        init="identity", bounds=[0.1, 0.1, 0.35, 0.01, 0.01, 0.01], aligner=dict(KITTI_ALIGNER)),
}


def scenario_inputs(name):
    """-> (scenario, data dict, fixed coords, fixed desc, moving xyz, moving desc, ground-truth camera_b_in_a, init pose12)"""
    sc = SCENARIOS[name]
    if sc["data"] == "icl":
        d = icl()
        fixed, fdesc, gt = d[sc["fixed"]], d["desc_50"], d["cam_50_in_00"]
    else:
        d = kitti()
        fixed, fdesc, gt = d["meas"][sc["frame"]]["uvuv"], d["meas"][sc["frame"]]["desc"], d["cam_in_00"][sc["frame"]]
    init = np.eye(3, 4).reshape(12) if sc["init"] == "identity" else np.asarray(gt, np.float64)
    return sc, d, fixed, fdesc, d["points_00"], d["desc_00"], gt, init


def oracle_align(name, with_prior=True):
    """runs the scenario through the CPU oracle; returns (result dict, manifold error vs ground truth)"""
    sc, d, fixed, fdesc, xyz, mdesc, gt, init = scenario_inputs(name)
    shape, kw = sc["finder"]
    if shape == "bruteforce":
        pf = O.BruteforceFinder(**kw)
    else:
        pf = O.ProjectiveFinder(d["K"], d["rows"], d["cols"], shape, **kw)
    pf.set_fixed(fixed, fdesc)
    pf.set_moving(xyz, mdesc)
    stereo = sc["kind"] == "stereo"
    if with_prior:
        init = IDENTITY_PRIOR[0]  # the motion-model slice seeds the estimate with its prediction (see SCENARIOS)
    r = O.align(pf, sc["kind"], d["K"], d["rows"], d["cols"], fixed, xyz, sc["diag"], init_pose=init,
                baseline=d["baseline"] if stereo else (0, 0, 0), inverse_depth_weighting=stereo, chi_threshold=sc["chi"],
                prior=IDENTITY_PRIOR if with_prior else None, **sc["aligner"])
    return r, O.t2tnq(O.pose_mul(r["pose"], gt))

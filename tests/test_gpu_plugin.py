"""Drop-in boundary on the GPU: modules instantiated BY CLASS NAME from the (hot-path subsets of the) shipped
.conf files, driven through the reference's module interface (setFixed / setMoving / compute, setRawData-like
calls, PARAM setters), compared with the CPU oracle and with the reference's own known-answer constants.
These tests read like tests/test_measurement_adaptors.cpp, tests/test_correspondence_finders.cpp and
tests/test_aligners.cpp of the reference."""
import pathlib

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_known_answers import CAM00, CAM01, K_KITTI

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
K_ICL = np.array([481.2, 0, 319.5, 0, -481, 239.5, 0, 0, 1], np.float32)  # tests/fixtures.hpp:577
BASELINE_M = 0.537166  # tests/fixtures.hpp:811-816


@pytest.fixture(scope="module")
def P(oracle):
    from srrg2_proslam_b200 import plugin
    return plugin


def kitti_pair(i):
    return O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png")


def same_cloud(g, o, coords):
    assert len(g[coords]) == len(o[coords])
    for k in (coords, "intensity", "desc"):
        assert np.array_equal(g[k], o[k]), k


def test_kitti_conf_stereo_adaptor(P):
    """configurations/kitti.conf as shipped: "adaptor_stereo_projective" -> extractors #8/#9 + epipolar finder #7"""
    m = P.Manager(GOLDEN / "configurations" / "kitti_hotpath.conf")
    a = m.get("adaptor_stereo_projective")
    for i in range(5):
        L, R = kitti_pair(i)
        g = a.stereo_adaptor(L, R)
        o = O.stereo_adaptor(L, R, O.extract_cfg(15, 1, 1000), "epipolar", 100, 0.5, 100, 0)
        assert g["status"] == 2 and len(g["uvuv"]) > 100
        same_cloud(g, {"uvuv": o["uvuv"], "intensity": o["intensity"], "desc": o["desc"]}, "uvuv")


def test_euroc_conf_stereo_adaptor(P):
    m = P.Manager(GOLDEN / "configurations" / "euroc_hotpath.conf")
    a = m.get("adaptor_stereo_projective")
    L, R = O.load_gray("scene_flow_image_left.png"), O.load_gray("scene_flow_image_right.png")
    g = a.stereo_adaptor(L, R)
    o = O.stereo_adaptor(L, R, O.extract_cfg(10, 1, 1000), "epipolar", 75, 0.5, 200, 0)
    same_cloud(g, o, "uvuv")


def test_adaptor_known_answers(P):  # tests/test_measurement_adaptors.cpp:26-56,110,130
    m = P.Manager()
    a = m.create("RawDataPreprocessorStereoProjective", "adaptor")
    ex = m.create("IntensityFeatureExtractorBinned3D").set("target_number_of_keypoints", 500).set("detector_threshold", 5)
    a.set("feature_extractor", ex).set("feature_extractor_right", ex)
    epi = m.create("CorrespondenceFinderDescriptorBasedEpipolar3D3D")
    bf = m.create("CorrespondenceFinderDescriptorBasedBruteforce3D3D")
    for f in (epi, bf):
        f.set("maximum_descriptor_distance", 100).set("maximum_distance_ratio_to_second_best", 0.8)
    L, R = kitti_pair(0)
    a.set("correspondence_finder", epi)
    assert len(a.stereo_adaptor(L, R)["uvuv"]) == 177  # :130
    a.set("correspondence_finder", bf)
    g = a.stereo_adaptor(L, R)
    assert len(g["uvuv"]) == 213  # :110
    same_cloud(g, O.stereo_adaptor(L, R, O.extract_cfg(5, 1, 500), "bruteforce", 100, 0.8), "uvuv")
    SL, SR = O.load_gray("scene_flow_image_left.png"), O.load_gray("scene_flow_image_right.png")
    assert len(a.stereo_adaptor(SL, SR)["uvuv"]) == 83  # :26
    a.set("correspondence_finder", epi)
    assert len(a.stereo_adaptor(SL, SR)["uvuv"]) == 115  # :51


def test_icl_conf_mono_depth_adaptor(P):  # tests/test_measurement_adaptors.cpp:75-88 with icl.conf's scale 0.001
    m = P.Manager(GOLDEN / "configurations" / "icl_hotpath.conf")
    a = [x for x in m.modules() if x.class_name == "RawDataPreprocessorMonocularDepth"][0]
    for i in (0, 1, 50):
        img, depth = O.load_gray(f"icl_image_rgb_{i}.png"), O.load_depth(f"icl_image_depth_{i}.png")
        g = a.mono_depth_adaptor(img, depth)
        o = O.mono_depth_adaptor(img, depth, O.extract_cfg(5, 1, 500), depth_scale=0.001)
        assert g["status"] == 2
        same_cloud(g, {"uvz": o["uvd"], "intensity": o["intensity"], "desc": o["desc"]}, "uvz")
        if i == 0:
            assert len(g["uvz"]) == 321  # :75
    # float depth images take the TYPE_32FC1 branch (monocular_depth.cpp:126-129)
    g = a.mono_depth_adaptor(img, depth.astype(np.float32))
    o = O.mono_depth_adaptor(img, depth.astype(np.float32), O.extract_cfg(5, 1, 500), depth_scale=0.001)
    same_cloud(g, {"uvz": o["uvd"], "intensity": o["intensity"], "desc": o["desc"]}, "uvz")


def test_finders_known_answers(P):  # tests/test_correspondence_finders.cpp:37-41,72,126,176-180,214,274,290
    m = P.Manager()
    ex = m.create("IntensityFeatureExtractorBinned2D").set("detector_threshold", 5).set("target_number_of_keypoints", 500)
    icl = [ex.extract(O.load_gray(f"icl_image_rgb_{i}.png")) for i in (0, 1, 50)]
    assert [len(f["xy"]) for f in icl] == [321, 338, 261]
    bf = m.create("CorrespondenceFinderDescriptorBasedBruteforce2D2D")
    for j, n in ((0, 319), (1, 226), (2, 117)):
        bf.set_fixed(icl[0]["xy"], icl[0]["desc"])
        bf.set_moving(icl[j]["xy"], icl[j]["desc"])
        fi, mi, d = bf.compute()
        assert len(fi) == n
        o = O.match_bruteforce(icl[0]["desc"], icl[j]["desc"], 50, 0.9)
        assert np.array_equal(fi, o[0]) and np.array_equal(mi, o[1]) and np.array_equal(d, o[2])
    # change flags: a second compute without new clouds keeps the result (bruteforce_impl.cpp:13-15)
    assert len(bf.compute()[0]) == 117
    L0, R0 = kitti_pair(0)
    l0, r0 = ex.extract(L0), ex.extract(R0)
    assert (len(l0["xy"]), len(r0["xy"])) == (446, 444)
    epi = m.create("CorrespondenceFinderDescriptorBasedEpipolar2D2D")
    epi.set_fixed(l0["xy"], l0["desc"])
    epi.set_moving(l0["xy"], l0["desc"])
    fi, mi, d = epi.compute()
    assert len(fi) == 446 and np.array_equal(fi, mi) and not d.any()  # :176-180
    epi.set_moving(r0["xy"], r0["desc"])
    assert len(epi.compute()[0]) == 150  # :274
    epi.set("epipolar_line_thickness_pixels", 1)
    epi.set_moving(r0["xy"], r0["desc"])
    assert len(epi.compute()[0]) == 241  # :290
    bf.set_fixed(l0["xy"], l0["desc"])
    bf.set_moving(r0["xy"], r0["desc"])
    assert len(bf.compute()[0]) == 237  # :214


def kitti_chain():
    c = O.extract_cfg(threshold=15, target=500)
    meas = [O.stereo_adaptor(*kitti_pair(i), c, "epipolar", 50, 0.8) for i in (0, 1)]
    xyz, _ = O.triangulate(meas[0]["uvuv"], K_KITTI, np.float32(718.856) * np.float32(BASELINE_M), 0.0)
    return meas, xyz, O.pose_mul(O.pose_inverse(CAM00), CAM01)


def test_projective_finder_state_machine(P):
    """kitti.conf "cf_projective_circle" (Circle4D3D: radius 50 -> 10, distance 25 -> 75, re-projection every 5th call)
    driven like the aligner drives it: 60 compute() calls with a moving estimate, then a second frame.  Every call
    must leave the same correspondences AND the same adaptive state as the oracle's restatement of
    correspondence_finder_projective_base_impl.cpp:104-293."""
    m = P.Manager(GOLDEN / "configurations" / "kitti_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    pf = sl.link("finder")
    pr = pf.link("projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    meas, xyz, cam01_in_00 = kitti_chain()
    of = O.ProjectiveFinder(K_KITTI, 376, 1241, "circle", max_desc_dist=75, ratio=0.8, min_matching_ratio=0.1,
                            min_desc_dist=25, desc_step=5, max_radius=50, min_radius=10, radius_step=10,
                            min_iterations=5, max_change_norm=0.01, iters_per_projection=5)
    target = O.pose_inverse(cam01_in_00)
    searches = 0
    for frame in range(2):
        pf.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
        pf.set_moving(xyz, meas[0]["desc"])
        of.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
        of.set_moving(xyz, meas[0]["desc"])
        for k in range(60):
            t = min(1.0, k / 20.0)  # estimate moves from identity to the ground truth, then rests
            pose = (np.eye(3, 4).reshape(12) * (1 - t) + target * t).astype(np.float32)
            pf.set_local_map_in_sensor(pose)
            of.set_estimate(pose)
            g, o = pf.compute(), of.compute()
            assert all(np.array_equal(a, b) for a, b in zip(g, o)), (frame, k)
            gs, os_ = pf.projective_state(), of.state()
            for key in ("radius", "descriptor_distance", "iteration", "converged"):
                assert gs[key] == os_[key], (frame, k, key, gs, os_)
        searches = pf.projective_state()["searches"]
        assert pf.projective_state()["converged"]
    assert 0 < searches < 40  # most calls keep the previous correspondences (projective_base_impl.cpp:171-178)
    assert al.class_name == "MultiAligner3DQR"


def test_projective_kdtree_finder_state_machine(P):
    """CorrespondenceFinderProjectiveKDTree4D3D with kitti.conf's "cf_projective_kd" parameters (configurations/kitti.conf:
    679-695) through the plugin mirror: the same adaptive state machine over the exact radius query (shape 3); every call
    equals the CPU restatement.  Parity with the reference's approximate KDTree is unpinned."""
    m = P.Manager()
    pf = m.create("CorrespondenceFinderProjectiveKDTree4D3D", "cf_projective_kd")
    assert pf.get("minimum_number_of_points_per_cluster") == 10  # correspondence_finder_projective_kdtree.h:24-28
    pf.set("maximum_descriptor_distance", 75).set("maximum_distance_ratio_to_second_best", 0.8)
    pf.set("minimum_matching_ratio", 0.1).set("minimum_descriptor_distance", 25).set("descriptor_distance_step_size_pixels", 5)
    pf.set("maximum_search_radius_pixels", 50).set("minimum_search_radius_pixels", 10).set("search_radius_step_size_pixels", 10)
    pf.set("minimum_number_of_iterations", 5).set("maximum_estimate_change_norm_for_convergence", 0.01)
    pf.set("number_of_solver_iterations_per_projection", 5)
    pr = m.create("PointIntensityDescriptor3fProjectorPinhole", "kd_projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241).set("range_min", 0.1).set("range_max", 1000.0)
    pf.set("projector", pr)
    meas, xyz, cam01_in_00 = kitti_chain()
    of = O.ProjectiveFinder(K_KITTI, 376, 1241, "kdtree", max_desc_dist=75, ratio=0.8, min_matching_ratio=0.1,
                            min_desc_dist=25, desc_step=5, max_radius=50, min_radius=10, radius_step=10,
                            min_iterations=5, max_change_norm=0.01, iters_per_projection=5)
    target = O.pose_inverse(cam01_in_00)
    pf.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    pf.set_moving(xyz, meas[0]["desc"])
    of.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    of.set_moving(xyz, meas[0]["desc"])
    n_last = 0
    for k in range(40):
        t = min(1.0, k / 20.0)
        pose = (np.eye(3, 4).reshape(12) * (1 - t) + target * t).astype(np.float32)
        pf.set_local_map_in_sensor(pose)
        of.set_estimate(pose)
        g, o = pf.compute(), of.compute()
        assert all(np.array_equal(a, b) for a, b in zip(g, o)), k
        gs, os_ = pf.projective_state(), of.state()
        for key in ("radius", "descriptor_distance", "iteration", "converged"):
            assert gs[key] == os_[key], (k, key, gs, os_)
        n_last = len(g[0])
    assert pf.projective_state()["converged"] and n_last > 30


def test_projective_finders_share_the_device_cache(P):
    """Two projective finder instances and other modules share the process-wide device context.  Each finder's cached
    clouds / lattice must survive (or be restored after) the other's uploads, a brute-force match, a triangulation and a
    stereo adaptor call in between -- the finders only set their changed-flags once, like the reference's callers."""
    m = P.Manager()
    meas, xyz, cam01_in_00 = kitti_chain()
    pose = O.pose_inverse(cam01_in_00).astype(np.float32)
    finders, oracles = [], []
    for shape, radius in (("Circle", 30), ("Square", 12)):
        pr = m.create("PointIntensityDescriptor3fProjectorPinhole")
        pr.set_camera_matrix(K_KITTI)
        pr.set("canvas_rows", 376).set("canvas_cols", 1241).set("range_min", 0.1).set("range_max", 1000.0)
        f = m.create(f"CorrespondenceFinderProjective{shape}4D3D")
        f.set("projector", pr).set("maximum_search_radius_pixels", radius).set("minimum_search_radius_pixels", radius)
        f.set("minimum_descriptor_distance", 60).set("maximum_descriptor_distance", 60).set("minimum_matching_ratio", 0.0)
        f.set("number_of_solver_iterations_per_projection", 1)
        o = O.ProjectiveFinder(K_KITTI, 376, 1241, shape.lower(), max_desc_dist=60, ratio=f.get("maximum_distance_ratio_to_second_best"),
                               min_matching_ratio=0.0, min_desc_dist=60, max_radius=radius, min_radius=radius, iters_per_projection=1)
        finders.append(f)
        oracles.append(o)
    # finder 0 matches frame 1 against the map, finder 1 matches frame 0 against itself: different clouds in the same cache
    clouds = [(meas[1]["uvuv"], meas[1]["desc"], pose), (meas[0]["uvuv"], meas[0]["desc"], np.eye(3, 4, dtype=np.float32).reshape(12))]
    for f, o, (fx, fd, _) in zip(finders, oracles, clouds):
        for x in (f, o):
            x.set_fixed(fx, fd)
            x.set_moving(xyz, meas[0]["desc"])
    bf = m.create("CorrespondenceFinderDescriptorBasedBruteforce4D3D")
    ad = m.create("RawDataPreprocessorStereoProjective", "adaptor")
    for rnd in range(3):
        for f, o, (_, _, T) in zip(finders, oracles, clouds):
            f.set_local_map_in_sensor(T)
            o.set_estimate(T)
            g, w = f.compute(), o.compute()
            assert len(g[0]) > 20 and all(np.array_equal(a, b) for a, b in zip(g, w)), rnd
            # other users of the context between two finder calls
            bf.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
            bf.set_moving(xyz, meas[0]["desc"])
            assert len(bf.compute()[0]) > 10
            if rnd == 1:
                assert len(ad.stereo_adaptor(*kitti_pair(2))["uvuv"]) > 50


def manifold_error(estimate, cam01_in_00):
    return O.t2tnq(O.pose_mul(estimate, cam01_in_00))


def test_kitti_conf_aligner(P):
    """configurations/kitti.conf "aligner": MultiAligner3DQR -> AlignerSliceProcessorProjectiveStereo (info [1,2,1],
    inverse-depth weighting, saturated chi 25) -> CorrespondenceFinderProjectiveCircle4D3D, GN damping 1,
    100 iterations.  KITTI 00 -> 01 from the identity guess; poses within 1e-6 m / 1e-6 rad of the oracle loop."""
    m = P.Manager(GOLDEN / "configurations" / "kitti_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    pr = sl.link("projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    meas, xyz, cam01_in_00 = kitti_chain()
    al.aligner_set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    al.aligner_set_moving(xyz, meas[0]["desc"])
    al.aligner_set_moving_in_fixed(np.eye(3, 4, dtype=np.float32))
    al.aligner_set_left_camera_in_right([-BASELINE_M, 0, 0])  # right_in_left = +0.537166 (tests/fixtures.hpp:813-816)
    g = al.aligner_compute()
    of = O.ProjectiveFinder(K_KITTI, 376, 1241, "circle", max_desc_dist=75, ratio=0.8, min_matching_ratio=0.1,
                            min_desc_dist=25, desc_step=5, max_radius=50, min_radius=10, radius_step=10,
                            min_iterations=5, max_change_norm=0.01, iters_per_projection=5)
    of.set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    of.set_moving(xyz, meas[0]["desc"])
    base = (K_KITTI.reshape(3, 3) @ np.array([-BASELINE_M, 0, 0], np.float32)).astype(np.float32)
    o = O.align(of, "stereo", K_KITTI, 376, 1241, meas[1]["uvuv"], xyz, [1, 2, 1], baseline=base,
                inverse_depth_weighting=True, chi_threshold=25.0, max_iterations=100, damping=1.0,
                min_num_inliers=6, min_num_correspondences=10,
                prior=(np.eye(3, 4).reshape(12), np.eye(6)))  # slice 2 of the conf: motion model, empty trajectory chunk
    assert g["status"] == o["status"] == O.ALIGNER_STATUS["Success"]
    assert g["iterations"] == len(o["stats"]) == 100
    assert np.array_equal(g["stats"][:, :3], o["stats"][:, :3])  # correspondences / inliers / outliers per iteration
    assert np.allclose(g["stats"][:, 3], o["stats"][:, 3], rtol=1e-7, atol=1e-9)
    assert all(np.array_equal(a, b) for a, b in zip(g["corr"], o["corr"]))
    d = O.t2tnq(O.pose_mul(O.pose_inverse(o["pose"]), g["pose"]))
    assert np.abs(d[:3]).max() < 1e-6 and np.abs(d[3:]).max() < 1e-6, d  # north_star: 1e-6 m / 1e-6 rad per pose
    e = manifold_error(g["pose"], cam01_in_00)
    assert np.all(np.abs(e[:3]) < 0.25) and np.all(np.abs(e[3:]) < 0.005), e  # see tests/test_oracle_solver.py


def unproject(uvz, K):
    K = K.reshape(3, 3).astype(np.float32)
    x = (uvz[:, 0] - K[0, 2]) / K[0, 0] * uvz[:, 2]
    y = (uvz[:, 1] - K[1, 2]) / K[1, 1] * uvz[:, 2]
    return np.stack([x, y, uvz[:, 2]], 1).astype(np.float32)


def test_icl_conf_aligner(P):
    """configurations/icl.conf "aligner": projective depth slice (info [1,1,10], saturated chi 10), Circle3D3D finder,
    GN damping from the file; frame 0 -> frame 1 of test_data/icl."""
    m = P.Manager(GOLDEN / "configurations" / "icl_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveDepth"][0]
    pf = sl.link("finder")
    pr = sl.link("projector")
    assert pf.link("projector").get("range_max") == pr.get("range_max")
    for p in (pr, pf.link("projector")):
        p.set_camera_matrix(K_ICL)
        p.set("canvas_rows", 480).set("canvas_cols", 640)
    ad = [x for x in m.modules() if x.class_name == "RawDataPreprocessorMonocularDepth"][0]
    meas = [ad.mono_depth_adaptor(O.load_gray(f"icl_image_rgb_{i}.png"), O.load_depth(f"icl_image_depth_{i}.png")) for i in (0, 1)]
    xyz = unproject(meas[0]["uvz"], K_ICL)
    al.aligner_set_fixed(meas[1]["uvz"], meas[1]["desc"])
    al.aligner_set_moving(xyz, meas[0]["desc"])
    al.aligner_set_moving_in_fixed(np.eye(3, 4, dtype=np.float32))
    g = al.aligner_compute()
    kw = {k: pf.get(k) for k in ("maximum_descriptor_distance", "minimum_descriptor_distance", "descriptor_distance_step_size_pixels",
                                 "maximum_search_radius_pixels", "minimum_search_radius_pixels", "search_radius_step_size_pixels",
                                 "minimum_number_of_iterations", "maximum_estimate_change_norm_for_convergence",
                                 "number_of_solver_iterations_per_projection", "minimum_matching_ratio",
                                 "maximum_distance_ratio_to_second_best")}
    of = O.ProjectiveFinder(K_ICL, 480, 640, "circle", max_desc_dist=kw["maximum_descriptor_distance"],
                            ratio=kw["maximum_distance_ratio_to_second_best"], min_matching_ratio=kw["minimum_matching_ratio"],
                            min_desc_dist=kw["minimum_descriptor_distance"], desc_step=kw["descriptor_distance_step_size_pixels"],
                            max_radius=int(kw["maximum_search_radius_pixels"]), min_radius=int(kw["minimum_search_radius_pixels"]),
                            radius_step=int(kw["search_radius_step_size_pixels"]), min_iterations=int(kw["minimum_number_of_iterations"]),
                            max_change_norm=kw["maximum_estimate_change_norm_for_convergence"],
                            iters_per_projection=int(kw["number_of_solver_iterations_per_projection"]),
                            range_min=pr.get("range_min"), range_max=pr.get("range_max"))
    of.set_fixed(meas[1]["uvz"], meas[1]["desc"])
    of.set_moving(xyz, meas[0]["desc"])
    o = O.align(of, "depth", K_ICL, 480, 640, meas[1]["uvz"], xyz, sl.get_numbers("diagonal_info_matrix"),
                chi_threshold=sl.link("robustifier").get("chi_threshold"), max_iterations=int(al.get("max_iterations")),
                damping=al.link("solver").link("algorithm").get("damping"), min_num_inliers=int(al.get("min_num_inliers")),
                min_num_correspondences=int(sl.get("min_num_correspondences")),
                prior=(np.eye(3, 4).reshape(12), np.eye(6)),  # slice 2 of the conf: motion model, empty trajectory chunk
                enable_inlier_only_runs=bool(al.get("enable_inlier_only_runs")),
                keep_only_inlier_correspondences=bool(al.get("keep_only_inlier_correspondences")))
    assert al.get("enable_inlier_only_runs") == 1 and al.get("keep_only_inlier_correspondences") == 1  # icl.conf:55-58
    assert g["status"] == o["status"] == O.ALIGNER_STATUS["Success"]
    assert np.array_equal(g["stats"][:, :3], o["stats"][:, :3])
    assert np.array_equal(g["inlier_run_stats"][:, :3], o["inlier_run_stats"][:, :3]) and len(g["inlier_run_stats"]) > 0
    assert all(np.array_equal(a, b) for a, b in zip(g["corr"], o["corr"]))
    d = O.t2tnq(O.pose_mul(O.pose_inverse(o["pose"]), g["pose"]))
    assert np.abs(d).max() < 1e-6, d
    assert np.abs(O.t2tnq(g["pose"])).max() < 0.02  # frames 0 and 1 are ~1 cm apart (tests/fixtures.hpp:597-603)


def test_aligner_not_enough_correspondences(P):
    m = P.Manager(GOLDEN / "configurations" / "kitti_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    pr = sl.link("projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    meas, xyz, _ = kitti_chain()
    rng = np.random.default_rng(0)
    al.aligner_set_fixed(meas[1]["uvuv"], rng.integers(0, 256, meas[1]["desc"].shape, dtype=np.uint8))  # unrelated descriptors
    al.aligner_set_moving(xyz, meas[0]["desc"])
    al.aligner_set_left_camera_in_right([-BASELINE_M, 0, 0])
    g = al.aligner_compute()
    assert g["status"] == O.ALIGNER_STATUS["NotEnoughCorrespondences"]


def paint_tracking_mask(rows, cols, projections, radius, to_left=False, to_right=False):
    """test-side restatement of the rectangles of intensity_feature_extractor_selective.cpp:80-144"""
    mask = np.zeros((rows, cols), np.uint8)
    rp = radius + 10
    for x, y in projections:
        row, col = int(np.round(y)), int(np.round(x))
        tl_row = max(row - rp, 0)
        h = min(2 * rp, rows - tl_row)
        if to_left and to_right:
            x0, w = 0, cols
        elif to_left:
            x0, w = 0, col
        elif to_right:
            x0, w = col, cols - col
        else:
            x0 = max(col - rp, 0)
            w = min(2 * rp, cols - x0)
        mask[tl_row:tl_row + h, x0:x0 + w] = 1
    return mask


@pytest.mark.parametrize("to_left,to_right,seeding", [(False, False, True), (True, False, True), (False, True, False), (True, True, True)])
def test_selective_extractor(P, to_left, to_right, seeding):
    """IntensityFeatureExtractorSelective3D (intensity_feature_extractor_selective.cpp:49-205): tracking keypoints in the
    painted rectangles first, seeded keypoints of the complement after, every FAST detection kept; without projections
    plain detection.  Oracle: cv::FAST-with-mask semantics = the masked mode of the oracle extractor."""
    m = P.Manager()
    ex = m.create("IntensityFeatureExtractorSelective3D", "selective")
    ex.set("detector_threshold", 25).set("enable_full_distance_to_left", int(to_left)).set("enable_full_distance_to_right", int(to_right))
    ex.set("enable_seeding_when_tracking", int(seeding))
    img = O.load_gray("kitti_city_image_left_1.png")
    rows, cols = img.shape
    cfg = O.extract_cfg(25, 1, 10 ** 6, 1, 1)
    plain = ex.extract(img)  # seeding phase: no projections
    o_all = O.extract_binned(img, cfg, mask=np.ones_like(img))
    assert len(plain["xy"]) == len(o_all["xy"]) > 300
    for k in ("xy", "intensity", "desc"):
        assert np.array_equal(plain[k], o_all[k]), k
    rng = np.random.default_rng(5)
    proj = np.stack([rng.uniform(0, cols - 1, 40), rng.uniform(0, rows - 1, 40), np.zeros(40)], 1).astype(np.float32)
    ex.set_projections(proj, 12)
    g = ex.extract(img)
    mask = paint_tracking_mask(rows, cols, proj[:, :2], 12, to_left, to_right)
    o_t = O.extract_binned(img, cfg, mask=mask)
    parts = [o_t] + ([O.extract_binned(img, cfg, mask=1 - mask)] if seeding else [])
    assert ex.number_of_tracking_keypoints() == len(o_t["xy"]) > 0
    for k in ("xy", "intensity", "desc"):
        assert np.array_equal(g[k], np.concatenate([p[k] for p in parts])), k
    # projections are consumed by one compute (:177): the next call seeds again
    again = ex.extract(img)
    assert np.array_equal(again["xy"], plain["xy"])

"""Host-side plugin mirror (libpslam_plugin.so), CPU part: the BOSS .conf reader / writer, the class registry and the
PARAM surface -- no compute calls.  The unchanged reference configurations are parsed when /root/reference is
present (build container); the derived hot-path subsets under tests/golden/configurations travel to the GPU box."""
import pathlib
import re
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "configurations"
REF_CONF = pathlib.Path("/root/reference/configurations")


@pytest.fixture(scope="module")
def P():
    subprocess.run(["make", "-C", str(ROOT / "srrg2_proslam_b200" / "csrc"), "-j8", "-s"], check=True)
    subprocess.run(["make", "-C", str(ROOT / "srrg2_proslam_b200" / "host"), "-s"], check=True)
    from srrg2_proslam_b200 import plugin
    return plugin


def test_exports_every_declared_symbol(P):
    header = (ROOT / "include" / "pslam_plugin.h").read_text()
    names = sorted(set(re.findall(r"\b(psp_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 40
    lib = P.lib()
    for n in names:
        assert hasattr(lib, n), n


# class names a .conf may use to select the CUDA-backed modules
HOT_CLASSES = ["IntensityFeatureExtractorBinned2D", "IntensityFeatureExtractorBinned3D", "IntensityFeatureExtractorSelective2D",
               "IntensityFeatureExtractorSelective3D", "RawDataPreprocessorStereoProjective",
               "RawDataPreprocessorMonocularDepth", "CorrespondenceFinderDescriptorBasedEpipolar2D2D",
               "CorrespondenceFinderDescriptorBasedEpipolar3D3D", "AlignerSliceProcessorProjective",
               "AlignerSliceProcessorProjectiveDepth", "AlignerSliceProcessorProjectiveStereo", "MultiAligner3DQR",
               "SceneClipperProjective3D", "LandmarkEstimatorProjectiveEKF3D", "LandmarkEstimatorProjectiveDepthEKF3D",
               "LandmarkEstimatorStereoProjectiveEKF3D", "ProjectivePointEKF3D", "ProjectiveDepthPointEKF3D",
               "StereoProjectivePointEKF3D", "LandmarkEstimatorWeightedMean2D3D", "LandmarkEstimatorWeightedMean3D3D",
               "LandmarkEstimatorWeightedMean4D3D", "LandmarkEstimatorPoseBasedSmoother2D3D",
               "LandmarkEstimatorPoseBasedSmoother3D3D", "LandmarkEstimatorPoseBasedSmoother4D3D",
               "MergerRigidStereoTriangulation", "MergerRigidStereoProjectiveEKF", "MergerProjectiveDepthEKF"]
HOT_CLASSES += [f"CorrespondenceFinderDescriptorBasedBruteforce{d}" for d in ("2D2D", "2D3D", "3D3D", "4D3D")]
HOT_CLASSES += [f"CorrespondenceFinderProjective{s}{d}" for s in ("Square", "Circle", "Rhombus", "KDTree") for d in ("2D3D", "3D3D", "4D3D")]


@pytest.mark.parametrize("name", HOT_CLASSES)
def test_registry(P, name):
    assert P.is_registered(name) and P.is_registered(name + "CUDA")
    m = P.Manager()
    a, b = m.create(name, "x"), m.create(name + "CUDA", "y")
    assert not a.is_generic and not b.is_generic and a.class_name == name and b.class_name == name + "CUDA"


def test_unknown_class_is_generic_and_round_trips(P):
    text = '"SomethingElse3D" { "#id" : 3, "name" : "s", "alpha" : 0.25, "topics" : [ "a", "b" ], "link" : { "#pointer" : 4 } }\n' \
           '"RobustifierSaturated" { "#id" : 4, // comment\n "chi_threshold" : 25 }'
    m = P.Manager(text=text)
    s = m.get("s")
    assert s.is_generic and s.get("alpha") == 0.25 and s.has("topics")
    out = ROOT / "tests" / "golden" / ".." / ".pytest_roundtrip.conf"
    m.write(out)
    m2 = P.Manager(out)
    assert len(m2) == 2 and m2.get("s").get("alpha") == 0.25
    out.unlink()


def test_defaults_match_reference_headers(P):
    m = P.Manager()
    e = m.create("IntensityFeatureExtractorBinned3D")  # base.h:24-58, binned.h:17-27
    assert (e.get_string("descriptor_type"), e.get_string("detector_type")) == ("ORB-256", "FAST")
    assert [e.get(k) for k in ("detector_threshold", "target_bin_width_pixels", "enable_non_maximum_suppression",
                               "target_number_of_keypoints", "number_of_detectors_horizontal",
                               "number_of_detectors_vertical")] == [10, 10, 1, 500, 3, 3]
    f = m.create("CorrespondenceFinderDescriptorBasedEpipolar3D3D")  # bruteforce.h:22-36, epipolar.h:22-32
    assert [f.get(k) for k in ("maximum_descriptor_distance", "minimum_matching_ratio", "maximum_disparity_pixels",
                               "epipolar_line_thickness_pixels")] == [50, 0.25, 100, 0]
    assert abs(f.get("maximum_distance_ratio_to_second_best") - 0.9) < 1e-7
    p = m.create("CorrespondenceFinderProjectiveCircle4D3D")  # projective_base.h:30-74
    assert [p.get(k) for k in ("minimum_descriptor_distance", "descriptor_distance_step_size_pixels",
                               "maximum_search_radius_pixels", "minimum_search_radius_pixels",
                               "search_radius_step_size_pixels", "minimum_number_of_iterations",
                               "number_of_solver_iterations_per_projection")] == [25, 5, 100, 10, 5, 10, 25]
    assert abs(p.get("maximum_estimate_change_norm_for_convergence") - 1e-5) < 1e-12
    a = m.create("RawDataPreprocessorStereoProjective")  # stereo_projective.h:17-39
    assert a.get_string("topic_camera_left") == "/camera_left/image_raw"
    assert a.link("correspondence_finder").class_name == "" or True  # default instance is unnamed / unregistered
    s = m.create("AlignerSliceProcessorProjectiveStereo")  # aligner_slice_processor_projective.cpp:7-20, .h:129-151
    assert s.link("robustifier").get("chi_threshold") == 100 * 100
    assert s.get("enable_inverse_depth_weighting") == 0 and s.get_string("frame_camera_right") == "camera_right"
    assert list(s.get_numbers("diagonal_info_matrix")) == [0, 0, 0]


def check_kitti_like(m, thr, epi):
    a = m.get("adaptor_stereo_projective")
    assert a.class_name == "RawDataPreprocessorStereoProjective" and not a.is_generic
    ex = a.link("feature_extractor")
    assert ex.class_name == "IntensityFeatureExtractorBinned3D" and ex.get("detector_threshold") == thr
    assert ex.get("target_number_of_keypoints") == 1000
    f = a.link("correspondence_finder")
    assert f.class_name == "CorrespondenceFinderDescriptorBasedEpipolar3D3D"
    assert [f.get("maximum_descriptor_distance"), f.get("maximum_disparity_pixels"), f.get("epipolar_line_thickness_pixels")] == epi
    assert abs(f.get("maximum_distance_ratio_to_second_best") - 0.5) < 1e-7
    al = m.get("aligner")
    assert al.class_name == "MultiAligner3DQR" and al.get("max_iterations") == 100 and al.get("min_num_inliers") == 6
    return al


def check_kitti(m):  # configurations/kitti.conf:229-255,262-315,484-501,593-615,834-875,980-1010
    al = check_kitti_like(m, 15, [100, 100, 0])
    assert al.link("solver").link("algorithm").get("damping") == 1
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    assert list(sl.get_numbers("diagonal_info_matrix")) == [1, 2, 1] and sl.get("enable_inverse_depth_weighting") == 1
    assert sl.link("robustifier").get("chi_threshold") == 25 and sl.get("min_num_correspondences") == 10
    pf = sl.link("finder")
    assert pf.class_name == "CorrespondenceFinderProjectiveCircle4D3D"
    assert [pf.get(k) for k in ("minimum_descriptor_distance", "maximum_descriptor_distance", "descriptor_distance_step_size_pixels",
                                "maximum_search_radius_pixels", "minimum_search_radius_pixels", "search_radius_step_size_pixels",
                                "number_of_solver_iterations_per_projection", "minimum_number_of_iterations")] == \
        [25, 75, 5, 50, 10, 10, 5, 5]
    assert abs(pf.get("maximum_distance_ratio_to_second_best") - 0.8) < 1e-7
    assert abs(pf.get("maximum_estimate_change_norm_for_convergence") - 0.01) < 1e-8
    assert abs(pf.get("minimum_matching_ratio") - 0.1) < 1e-7
    pr = sl.link("projector")
    assert pr.class_name == "PointIntensityDescriptor3fProjectorPinhole" and pr.get("range_max") == 1000
    assert abs(pr.get("range_min") - 0.1) < 1e-7


def check_icl(m):  # configurations/icl.conf
    ad = [x for x in m.modules() if x.class_name == "RawDataPreprocessorMonocularDepth"][0]
    assert abs(ad.get("depth_scaling_factor_to_meters") - 0.001) < 1e-9
    ex = ad.link("feature_extractor")
    assert ex.get("detector_threshold") == 5 and ex.get("target_number_of_keypoints") == 500
    al = m.get("aligner")
    assert al.get("enable_inlier_only_runs") == 1 and al.get("keep_only_inlier_correspondences") == 1
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveDepth"][0]
    assert list(sl.get_numbers("diagonal_info_matrix")) == [1, 1, 10]
    assert sl.link("finder").class_name == "CorrespondenceFinderProjectiveCircle3D3D"


@pytest.mark.skipif(not REF_CONF.exists(), reason="reference configurations only exist in the build container")
def test_unchanged_reference_configurations_load(P):
    for f in sorted(REF_CONF.glob("*.conf")):
        m = P.Manager(f)
        assert len(m) > 40
        hot = [x.class_name for x in m.modules() if not x.is_generic]
        assert any(c.startswith("IntensityFeatureExtractorBinned") for c in hot), f
        assert "MultiAligner3DQR" in hot and any(c.startswith("CorrespondenceFinderProjectiveCircle") for c in hot)
    check_kitti(P.Manager(REF_CONF / "kitti.conf"))
    check_icl(P.Manager(REF_CONF / "icl.conf"))
    check_kitti_like(P.Manager(REF_CONF / "euroc.conf"), 10, [75, 200, 0])


def test_hotpath_fixtures(P):
    check_kitti(P.Manager(GOLDEN / "kitti_hotpath.conf"))
    check_icl(P.Manager(GOLDEN / "icl_hotpath.conf"))
    check_kitti_like(P.Manager(GOLDEN / "euroc_hotpath.conf"), 10, [75, 200, 0])


@pytest.mark.skipif(not REF_CONF.exists(), reason="reference configurations only exist in the build container")
def test_hotpath_fixtures_are_current(P, tmp_path):
    """the committed fixtures are what tools/make_golden.py derives from the unchanged files"""
    m = P.Manager(REF_CONF / "kitti.conf")
    m.write(tmp_path / "k.conf", ["adaptor_stereo_projective", "aligner", "cf_bruteforce", "clipper_stereo_projective", "landmark_estimator_ekf",
                                 "landmark_estimator_weighted_mean", "landmark_estimator_smoother", "merger_triangulation", "merger_ekf"])
    assert (tmp_path / "k.conf").read_text() == (GOLDEN / "kitti_hotpath.conf").read_text()


def test_merger_modules_of_the_shipped_configurations(P):
    """kitti.conf:188-230 / :552-600, icl.conf MergerProjectiveDepthEKF: parameters and links of the merger modules"""
    m = P.Manager(GOLDEN / "kitti_hotpath.conf")
    for name, cls, est in (("merger_triangulation", "MergerRigidStereoTriangulation", "LandmarkEstimatorPoseBasedSmoother4D3D"),
                           ("merger_ekf", "MergerRigidStereoProjectiveEKF", "LandmarkEstimatorStereoProjectiveEKF3D")):
        g = m.get(name)
        assert not g.is_generic and g.class_name == cls
        assert [g.get(k) for k in ("enable_binning", "enable_conservative_addition", "number_of_row_bins", "number_of_col_bins",
                                   "target_merge_ratio")] == [1, 0, 20, 60, 0.5]
        assert g.link("projector").class_name == "PointIntensityDescriptor3fProjectorPinhole"
        assert g.link("landmark_estimator").class_name == est
        assert g.link("triangulator").is_generic  # TriangulatorRigidStereo: control plane, kept as a generic module
    d = [x for x in P.Manager(GOLDEN / "icl_hotpath.conf").modules() if x.class_name == "MergerProjectiveDepthEKF"][0]
    assert not d.is_generic and d.link("landmark_estimator").class_name == "LandmarkEstimatorProjectiveDepthEKF3D"
    fresh = P.Manager().create("MergerRigidStereoTriangulation")  # merger_projective.h:41-66 defaults
    assert [fresh.get(k) for k in ("maximum_distance_appearance", "number_of_row_bins", "number_of_col_bins", "target_merge_ratio",
                                   "enable_conservative_addition")] == [50, 10, 30, 0.5, 0]


def test_errors_keep_the_reference_texts(P):
    with pytest.raises(P.PluginError, match="dangling #pointer"):
        P.Manager(text='"RawDataPreprocessorStereoProjective" { "#id" : 1, "feature_extractor" : { "#pointer" : 9 } }')
    with pytest.raises(P.PluginError, match="wrong class"):
        P.Manager(text='"RawDataPreprocessorStereoProjective" { "#id" : 1, "feature_extractor" : { "#pointer" : 2 } }\n'
                       '"RobustifierSaturated" { "#id" : 2 }')
    m = P.Manager()
    img = np.zeros((64, 64), np.uint8)
    e = m.create("IntensityFeatureExtractorBinned3D")
    e.set("number_of_detectors_vertical", 0)
    with pytest.raises(P.PluginError, match=r"IntensityFeatureExtractor::init\|ERROR: invalid number of vertical detectors"):
        e.extract(img)  # binned.cpp:13-16
    e.set("number_of_detectors_vertical", 3).set("detector_type", "SIFT")
    with pytest.raises(P.PluginError, match=r"_setDetector\|ERROR: unknown detector type chosen: SIFT"):
        e.extract(img)  # base.cpp:139-143
    f = m.create("CorrespondenceFinderDescriptorBasedBruteforce3D3D")
    with pytest.raises(P.PluginError, match=r"CorrespondenceFinderDescriptorBased::compute\|ERROR: fixed not set"):
        f.compute()  # bruteforce_impl.cpp:203-205
    a = m.create("RawDataPreprocessorMonocularDepth")
    with pytest.raises(P.PluginError, match="unknown depth image type"):
        a.mono_depth_adaptor(img, np.zeros((64, 64), np.float64))  # monocular_depth.cpp:131-134


def test_no_cpu_fallback(P):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = P.Manager(GOLDEN / "kitti_hotpath.conf")
    with pytest.raises(P.PluginError, match="no usable CUDA device"):
        m.get("adaptor_stereo_projective").stereo_adaptor(np.zeros((376, 1241), np.uint8), np.zeros((376, 1241), np.uint8))


def test_merger_configuration_errors_keep_the_reference_texts(P):
    """merger_projective_impl.cpp:21-47: the checks of compute() fire before any device work, with the reference's texts"""
    m = P.Manager()
    g = m.create("MergerRigidStereoTriangulation", "g")
    pr = g.link("projector")
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    meas = np.array([[10, 10, 5, 10]], np.float32)
    g.set("number_of_row_bins", 1000)
    with pytest.raises(P.PluginError, match="row bin width must be at least 1 pixel"):
        g.merger_select_updates(meas, [0], [10.0])
    g.set("number_of_row_bins", 10).set("number_of_col_bins", 5000)
    with pytest.raises(P.PluginError, match="col bin width must be at least 1 pixel"):
        g.merger_select_additions(meas)
    g.set("number_of_col_bins", 30).set("enable_conservative_addition", 1)
    with pytest.raises(P.PluginError, match="conservative addition is currently disabled"):
        g.merger_select_additions(meas)
    g.set("enable_conservative_addition", 0).set("projector", None)
    with pytest.raises(P.PluginError, match="projector not set"):
        g.merger_plan(meas, [0], [10.0])
    assert g.merger_wants_additions(5, 100, 0) and g.merger_wants_additions(5, 100, 50)      # :56-58, :158-165
    assert not g.merger_wants_additions(100, 300, 150) and not g.merger_wants_additions(40, 40, 40)


def test_selective_extractor_known_answers(P):
    """Row a5, tests/test_feature_extractors.cpp:168-213 (KITTI, IntensityFeatureExtractorSelective_GFFT_ORB256).  That test
    runs the selective extractor with detector_type "GFTT" in BOTH phases (the "TRACKING with FAST" comments are stale:
    _keypoint_detector is the GFTT detector re-created with target 1000), so its constants pin the extractor's own logic --
    the tracking rectangles of intensity_feature_extractor_selective.cpp:63-144 -- and ORB's 31-pixel border drop, not a
    corner detector of the hot path.  cv2's GFTTDetector stands in for cv::GFTTDetector (test side only, parameters of
    intensity_feature_extractor_base.cpp:107-116: quality 0.01, minimum distance = target_bin_width_pixels = 10); the mask
    is painted by the plugin class IntensityFeatureExtractorSelective2D itself (host-side code of the product)."""
    import cv2
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    img = O.load_gray("kitti_city_image_left_0.png")
    rows, cols = img.shape

    def gftt(target, mask=None):
        kps = cv2.GFTTDetector_create(target, 0.01, 10).detect(img, mask)
        pts = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
        keep = (pts[:, 0] >= 31) & (pts[:, 0] < cols - 31) & (pts[:, 1] >= 31) & (pts[:, 1] < rows - 31)  # ORB::compute
        return pts[keep]

    m = P.Manager()
    ex = m.create("IntensityFeatureExtractorSelective2D", "selective")
    ex.set("enable_seeding_when_tracking", 0)
    seeds = gftt(100)
    assert len(seeds) == 94  # :182
    for radius, expected in ((100, 719), (50, 581), (10, 294), (5, 237)):  # :191,198,205,212
        mask = ex.paint_tracking_mask(rows, cols, seeds, radius)
        assert len(gftt(1000, mask)) == expected, radius
    # the three "full bar" variants (selective.cpp:80-128) against a direct restatement
    rng = np.random.default_rng(1)
    proj = np.stack([rng.uniform(0, cols - 1, 30), rng.uniform(0, rows - 1, 30)], 1).astype(np.float32)
    for to_left, to_right in ((1, 0), (0, 1), (1, 1)):
        ex.set("enable_full_distance_to_left", to_left).set("enable_full_distance_to_right", to_right)
        want = np.zeros((rows, cols), np.uint8)
        for x, y in proj:
            r, c = int(np.round(y)), int(np.round(x))
            t = max(r - 22, 0)
            h = min(44, rows - t)
            x0, w = (0, cols) if (to_left and to_right) else ((0, c) if to_left else (c, cols - c))
            want[t:t + h, x0:x0 + w] = 1
        assert np.array_equal(ex.paint_tracking_mask(rows, cols, proj, 12), want)

"""The reference's six conf-driven aligner tests (tests/test_aligners.cpp:883-1337) on the GPU, written the way the
reference writes them: load the shipped conf, take "aligner", swap the slice / finder exactly as the test does, set the
clouds, compute, compare t2tnq(aligner->movingInFixed() * camera_b_in_a) with the reference's own bounds.  The CUDA
path (projective search, brute-force match, fused linearise + H,b + prior + solve) must also agree with the CPU oracle:
identical correspondences and per-iteration inlier counts, poses within 1e-6 m / 1e-6 rad (north_star)."""
import pathlib

import numpy as np
import pytest

import aligner_fixtures as A
import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def P(oracle):
    from srrg2_proslam_b200 import plugin
    return plugin


def configure(P, name):
    sc, d, fixed, fdesc, xyz, mdesc, gt, init = A.scenario_inputs(name)
    m = P.Manager(GOLDEN / "configurations" / f"{sc['data']}_hotpath.conf")
    al = m.get("aligner")
    assert al.aligner_num_slice_processors() == 2  # ASSERT_EQ(aligner->param_slice_processors.size(), 2)
    slice_class = {"mono": "AlignerSliceProcessorProjective", "depth": "AlignerSliceProcessorProjectiveDepth",
                   "stereo": "AlignerSliceProcessorProjectiveStereo"}[sc["kind"]]
    if sc["kind"] == "mono":  # :896-901: a fresh projective-only slice replaces slice 0
        sl = m.create("AlignerSliceProcessorProjective", "slice_projective")
        assert sl.link("robustifier") is not None
        al.aligner_set_slice_processor(0, sl)
    else:
        sl = [x for x in m.modules() if x.class_name == slice_class][0]
    projector = m.create("PointIntensityDescriptor3fProjectorPinhole", "fixture_projector")  # tests/fixtures.hpp:626-630,823-830
    projector.set_camera_matrix(d["K"])
    projector.set("canvas_rows", d["rows"]).set("canvas_cols", d["cols"])
    if sc["data"] == "kitti":
        projector.set("range_min", 0.1).set("range_max", 1000.0)
        sl.link("robustifier").set("chi_threshold", 1000)  # :1118
    shape, kw = sc["finder"]
    if shape == "bruteforce":
        if sc["data"] == "icl":
            finder = m.get("cf_bruteforce_2d" if sc["kind"] == "mono" else "cf_bruteforce_3d")  # :903-905, :982-984
            assert finder.get("maximum_descriptor_distance") == kw["max_dist"]
        else:
            finder = m.create("CorrespondenceFinderDescriptorBasedBruteforce4D3D")  # :1121-1123
            finder.set("maximum_descriptor_distance", kw["max_dist"])
            finder.set("maximum_distance_ratio_to_second_best", kw["ratio"])
    else:
        finder = m.get("cf_projective_circle")
        finder.set("projector", projector)
        if sc["data"] == "kitti":  # :1223-1228
            finder.set("minimum_descriptor_distance", 50).set("maximum_descriptor_distance", 100)
            finder.set("maximum_distance_ratio_to_second_best", 0.8)
            finder.set("minimum_search_radius_pixels", 10).set("maximum_search_radius_pixels", 50)
            finder.set("number_of_solver_iterations_per_projection", 5)
    sl.set("finder", finder).set("projector", projector)
    al.fixture_finder = finder
    al.fixture_slice = sl
    al.fixture_manager = m
    sl.set("diagonal_info_matrix", sc["diag"])
    al.aligner_set_fixed(fixed, fdesc)
    al.aligner_set_moving(xyz, mdesc)
    al.aligner_set_trajectory_chunk(np.zeros((0, 12), np.float32))  # empty "trajectory_chunk" property
    al.aligner_set_moving_in_fixed(init.astype(np.float32))
    if sc["kind"] == "stereo":
        al.aligner_set_left_camera_in_right([-A.BASELINE_M, 0, 0])  # slice->setPlatform(_platform)
    return al, gt


@pytest.mark.parametrize("name", sorted(A.SCENARIOS))
def test_reference_aligner_scenarios_gpu(P, name):
    al, gt = configure(P, name)
    g = al.aligner_compute()
    assert g["status"] == 1  # AlignerBase::Success
    e = O.t2tnq(O.pose_mul(g["pose"], gt))
    assert np.all(np.abs(e) < A.SCENARIOS[name]["bounds"]), e  # the reference's ASSERT_LT_ABS bounds
    o, eo = A.oracle_align(name)
    assert g["iterations"] == len(o["stats"])
    assert np.array_equal(g["stats"][:, :3], o["stats"][:, :3])
    assert np.allclose(g["stats"][:, 3], o["stats"][:, 3], rtol=1e-7, atol=1e-9)
    assert all(np.array_equal(a, b) for a, b in zip(g["corr"], o["corr"]))
    if "inlier_run_stats" in o:
        assert np.array_equal(g["inlier_run_stats"][:, :3], o["inlier_run_stats"][:, :3])
    d = O.t2tnq(O.pose_mul(O.pose_inverse(o["pose"]), g["pose"]))
    assert np.abs(d).max() < 1e-6, d


def test_constant_velocity_prior_seeds_and_pulls(P):
    """trajectory chunk with two poses: the motion model predicts the next pose; a strong information matrix keeps the
    estimate at the prediction, the default (identity) leaves the result to the measurements"""
    al, gt = configure(P, "kitti_00to01_projective_circle")
    step = O.pose_inverse(gt)  # local map in sensor after one frame of the true motion
    chunk = [O.pose_mul(gt, O.pose_inverse(gt)), gt]  # robot at identity, then at camera_01_in_00
    al.aligner_set_trajectory_chunk(np.asarray(chunk, np.float32))
    pred = O.constant_velocity_prediction(chunk)
    al.aligner_set_prior_information(1e12 * np.eye(6))
    g = al.aligner_compute()
    assert np.abs(O.t2tnq(O.pose_mul(O.pose_inverse(pred), g["pose"]))).max() < 1e-4
    assert np.abs(pred - O.pose_mul(step, step)).max() < 1e-5  # two steps of the same motion (the chunk travels as fp32)


@pytest.mark.parametrize("name", sorted(n for n in A.SCENARIOS if A.SCENARIOS[n]["finder"][0] != "bruteforce"))
def test_device_resident_alignment_equals_call_by_call(P, name, monkeypatch):
    """pslam_projective_align (the finder's state machine on the device, one download per batch of search phases) against the
    call-by-call loop (one pslam_projective_match_gn per search): bit-identical poses, per-iteration stats, correspondences
    and finder state -- on a frame and on the frame after it (the finder carries radius / descriptor distance over)."""
    def run(env):
        monkeypatch.setenv("PSLAM_ALIGN_DEVICE", env)
        al, gt = configure(P, name)
        out = []
        for _ in range(2):
            sc, d, fixed, fdesc, xyz, mdesc, gt_, init = A.scenario_inputs(name)
            al.aligner_set_fixed(fixed, fdesc)
            al.aligner_set_moving(xyz, mdesc)
            al.aligner_set_moving_in_fixed(init.astype(np.float32))
            g = al.aligner_compute()
            g["finder"] = al.fixture_finder.projective_state()
            out.append(g)
        return out
    a, b = run("1"), run("0")
    for ga, gb in zip(a, b):
        assert ga["status"] == gb["status"] and ga["iterations"] == gb["iterations"]
        assert np.array_equal(ga["pose"], gb["pose"])
        assert np.array_equal(ga["stats"], gb["stats"])
        assert all(np.array_equal(x, y) for x, y in zip(ga["corr"], gb["corr"]))
        assert np.array_equal(ga["inlier_run_stats"], gb["inlier_run_stats"])
        assert ga["finder"] == gb["finder"]


@pytest.mark.parametrize("case", ["repeat_with_wider_search", "too_few_correspondences", "single_iteration_budget",
                                  "more_phases_than_the_log"])
def test_device_resident_alignment_hands_back(P, case, monkeypatch, capfd):
    """the decisions pslam_projective_align leaves to the caller's loop: a low matching ratio while the search window can still
    grow (the finder repeats the call, base_impl.cpp:228-262), fewer correspondences than the slice wants, and a budget that ends
    inside the first phase -- same results as the call-by-call path in every case"""
    name = "kitti_00to01_projective_circle"

    def run(env):
        monkeypatch.setenv("PSLAM_ALIGN_DEVICE", env)
        al, gt = configure(P, name)
        finder = al.fixture_finder
        if case == "too_few_correspondences":
            al.fixture_slice.set("min_num_correspondences", 1000)
        if case == "single_iteration_budget":
            al.set("max_iterations", 1)
        if case == "more_phases_than_the_log":  # a search before every solver iteration, never converged: 100 phases > 64
            finder.set("number_of_solver_iterations_per_projection", 1).set("minimum_number_of_iterations", 1000)
        out = []
        for frame in range(3):
            sc, d, fixed, fdesc, xyz, mdesc, gt_, init = A.scenario_inputs(name)
            if case == "repeat_with_wider_search" and frame == 1:
                finder.set("minimum_matching_ratio", 0.95)  # the shrunken window of frame 0 cannot satisfy this: internal repeat
            al.aligner_set_fixed(fixed, fdesc)
            al.aligner_set_moving(xyz, mdesc)
            al.aligner_set_moving_in_fixed(init.astype(np.float32))
            g = al.aligner_compute()
            g["finder"] = finder.projective_state()
            out.append(g)
        return out

    a, b = run("1"), run("0")
    for ga, gb in zip(a, b):
        assert ga["status"] == gb["status"] and ga["iterations"] == gb["iterations"]
        assert np.array_equal(ga["pose"], gb["pose"])
        assert np.array_equal(ga["stats"], gb["stats"])
        assert all(np.array_equal(x, y) for x, y in zip(ga["corr"], gb["corr"]))
        assert ga["finder"] == gb["finder"]
    err = capfd.readouterr().err
    if case == "repeat_with_wider_search":
        assert err.count("triggering internal repeat with increased search radius") == 2  # once per path: it really happened
    if case == "too_few_correspondences":
        assert all(g["status"] != 1 for g in a)  # AlignerBase::NotEnoughCorrespondences
    if case == "more_phases_than_the_log":
        assert all(g["iterations"] == 100 and g["finder"]["searches"] % 100 == 0 for g in a)


@pytest.mark.parametrize("cls", ["CorrespondenceFinderProjectiveSquare4D3D", "CorrespondenceFinderProjectiveRhombus4D3D",
                                 "CorrespondenceFinderProjectiveKDTree4D3D"])
def test_device_resident_alignment_other_windows(P, cls, monkeypatch):
    """the device-resident registration with the other window shapes of the projective finder family (square, rhombus, exact
    radius query of the KD-tree variant) in place of the circle: identical to the call-by-call path"""
    name = "kitti_00to01_projective_circle"

    def run(env):
        monkeypatch.setenv("PSLAM_ALIGN_DEVICE", env)
        al, gt = configure(P, name)
        m = al.fixture_manager
        finder = m.create(cls)
        circle = al.fixture_finder
        finder.set("projector", circle.link("projector"))
        for k in ("minimum_descriptor_distance", "maximum_descriptor_distance", "maximum_distance_ratio_to_second_best",
                  "minimum_search_radius_pixels", "maximum_search_radius_pixels", "number_of_solver_iterations_per_projection"):
            finder.set(k, circle.get(k))
        al.fixture_slice.set("finder", finder)
        out = []
        for frame in range(2):
            sc, d, fixed, fdesc, xyz, mdesc, gt_, init = A.scenario_inputs(name)
            al.aligner_set_fixed(fixed, fdesc)
            al.aligner_set_moving(xyz, mdesc)
            al.aligner_set_moving_in_fixed(init.astype(np.float32))
            g = al.aligner_compute()
            g["finder"] = finder.projective_state()
            out.append(g)
        return out

    a, b = run("1"), run("0")
    for ga, gb in zip(a, b):
        assert ga["status"] == gb["status"] == 1 and ga["iterations"] == gb["iterations"]
        assert np.array_equal(ga["pose"], gb["pose"]) and np.array_equal(ga["stats"], gb["stats"])
        assert all(np.array_equal(x, y) for x, y in zip(ga["corr"], gb["corr"])) and len(ga["corr"][0]) > 20
        assert ga["finder"] == gb["finder"]

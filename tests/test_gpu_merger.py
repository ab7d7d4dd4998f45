"""N3: the binning of MergerProjective_::compute on the GPU -- pslam_merger_select_updates /
pslam_merger_select_additions against the CPU oracle's sequential walk (bit exact: every decision, the blocked-bin bitmap,
the order of the addition candidates), on the inputs of the reference's merger tests (tests/test_mergers.cpp) and on random
crowded clouds with ties."""
import numpy as np
import pytest

import oracle_lib as O
from merger_fixtures import icl_00_01, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(device=0, max_images=2, max_rows=480, max_cols=1241, max_features=4096, max_raw_per_bin=8192)
    yield c
    c.close()


def both(ctx, meas, moving, resp, rows, cols, kind, row_bins=10, col_bins=30, gate=50.0, binning=True):
    from srrg2_proslam_b200 import capi
    cfg = capi.merger_cfg(rows, cols, row_bins, col_bins, gate, binning, kind)
    o_sel, o_occ = O.merger_select_updates(meas, moving, resp, rows, cols, row_bins, col_bins, gate, binning, kind)
    g_sel, g_occ = ctx.merger_select_updates(cfg, meas, moving, resp)
    assert np.array_equal(g_sel, o_sel) and np.array_equal(g_occ, o_occ)
    o_win = O.merger_select_additions(meas, o_occ, rows, cols, row_bins, col_bins, binning, kind)
    g_win = ctx.merger_select_additions(cfg, meas, g_occ)
    assert np.array_equal(g_win, o_win)
    p_sel, p_occ, p_win = ctx.merger_plan(cfg, meas, moving, resp)  # both passes in one round trip
    assert np.array_equal(p_sel, o_sel) and np.array_equal(p_occ, o_occ) and np.array_equal(p_win, o_win)
    assert np.array_equal(ctx.merger_select_additions(cfg, meas, None),
                          O.merger_select_additions(meas, None, rows, cols, row_bins, col_bins, binning, kind))
    return g_sel, g_win


def test_icl_merger_test_inputs(ctx):
    """tests/test_mergers.cpp:248-357: 00 -> 00 adds nothing, 00 -> 01 grows the scene from 321 to 337 points"""
    m0, m1, corr = icl_00_01()
    n = len(m0["uvd"])
    sel, win = both(ctx, m0["uvd"], np.arange(n), np.zeros(n), 480, 640, "depth")
    assert len(win) == 0
    sel, win = both(ctx, m1["uvd"], corr[:, 1], corr[:, 2], 480, 640, "depth")
    assert 321 + len(win) == 337


def test_kitti_stereo_measurements(ctx):
    """stereo adaptor output of KITTI frame 1 merged over the correspondences of an exhaustive match against frame 0"""
    c = O.extract_cfg(threshold=15, target=1000)
    m = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"), c,
                          "epipolar", 100, 0.5, 100, 0) for i in (0, 1)]
    fi, mi, d = O.match_bruteforce(m[0]["desc"], m[1]["desc"], 75, 0.8)
    sel, win = both(ctx, m[1]["uvuv"], mi, d, 376, 1241, "stereo")
    assert 0 < sel.sum() <= len(mi) and len(win) > 0
    both(ctx, m[1]["uvuv"], mi, d, 376, 1241, "base")
    both(ctx, m[1]["uvuv"], mi, d, 376, 1241, "stereo", binning=False)


@pytest.mark.parametrize("seed", range(6))
def test_random_clouds(ctx, seed):
    dim = 4 if seed % 3 else 3
    meas, moving, resp = random_case(seed, 4000, 2500, dim=dim, crowded=seed % 2 == 1)
    both(ctx, meas, moving, resp, 376, 1241, "stereo" if dim == 4 else "depth")
    both(ctx, meas, moving, resp, 376, 1241, "base", row_bins=47, col_bins=155)  # 7 488 bins of 8 px


def test_edges(ctx):
    from srrg2_proslam_b200 import capi
    cfg = capi.merger_cfg(376, 1241)
    meas, moving, resp = random_case(9, 50, 20)
    sel, occ = ctx.merger_select_updates(cfg, meas, moving[:0], resp[:0])  # no correspondences: nothing blocked
    assert len(sel) == 0 and not occ.any()
    assert len(ctx.merger_select_additions(cfg, meas[:0], None)) == 0
    outside = meas.copy()
    outside[7, 0] = 5000.0  # a measurement outside the canvas: the reference asserts, the library refuses
    with pytest.raises(capi.PslamError):
        ctx.merger_select_additions(cfg, outside, None)
    with pytest.raises(capi.PslamError):
        ctx.merger_select_updates(capi.merger_cfg(376, 1241, row_bins=400), meas, moving, resp)  # bin width < 1 px


def test_conf_merger_pass_kitti_00_to_01(oracle):
    """kitti.conf "merger_ekf" (MergerRigidStereoProjectiveEKF -> "landmark_estimator_ekf" -> "point_filter") instantiated by
    class name: one merger pass of frame 01 into the landmarks of frame 00, chained like MergerProjective_::compute
    (merger_projective_impl.cpp:61-171): binned update selection -> batched EKF update of the selected landmarks -> merge
    count -> binned addition candidates -> their triangulation; every stage against the CPU restatement"""
    import pathlib
    from srrg2_proslam_b200 import plugin as P
    from test_oracle_known_answers import CAM00, CAM01, K_KITTI
    m = P.Manager(pathlib.Path(__file__).resolve().parent / "golden" / "configurations" / "kitti_hotpath.conf")
    mg = m.get("merger_ekf")
    assert mg.class_name == "MergerRigidStereoProjectiveEKF" and not mg.is_generic
    pr = mg.link("projector")
    pr.set_camera_matrix(K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    est = mg.link("landmark_estimator")
    b_x = np.float32(718.856) * np.float32(0.537166)
    est.link("filter").filter_set_camera(K_KITTI, (float(b_x), 0.0))
    rows_bins, col_bins, gate = int(mg.get("number_of_row_bins")), int(mg.get("number_of_col_bins")), mg.get("maximum_distance_appearance")
    assert (rows_bins, col_bins, gate) == (20, 60, 100)

    c = O.extract_cfg(threshold=15, target=1000)
    f = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"), c,
                          "epipolar", 100, 0.5, 100, 0) for i in (0, 1)]
    scene, _ = O.triangulate(f[0]["uvuv"], K_KITTI, b_x, 0.0)           # landmarks: frame 00, world = camera 00
    fixed, moving, resp = O.match_bruteforce(f[0]["desc"], f[1]["desc"], 75, 0.8)
    meas = f[1]["uvuv"]
    cam01_in_00 = O.pose_mul(O.pose_inverse(CAM00), CAM01).astype(np.float32)

    # 1. update pass
    g_sel = mg.merger_select_updates(meas, moving, resp)
    o_sel, o_occ = O.merger_select_updates(meas, moving, resp, 376, 1241, rows_bins, col_bins, gate, True, "stereo")
    assert np.array_equal(g_sel, o_sel) and 0 < g_sel.sum() < len(moving)
    # 2. _updatePoint of the selected correspondences = the linked estimator, batched
    state, cov = scene[fixed[g_sel]], np.tile(np.eye(3, dtype=np.float32), (int(g_sel.sum()), 1, 1))
    est.estimator_set_transforms(cam01_in_00, cam01_in_00)
    g = est.estimator_compute_batch(state, cov, meas[moving[g_sel]])
    o = O.landmarks_ekf_update("stereo", K_KITTI, (float(b_x), 0.0), cam01_in_00, cam01_in_00, state, cov, meas[moving[o_sel]],
                               min_cov=est.get("minimum_state_element_covariance"),
                               max_cov_norm2=est.get("maximum_covariance_norm_squared"),
                               max_dist2=est.get("maximum_distance_geometry_meters_squared"))
    assert np.array_equal(g[3], o[3])
    merged = int(g[3].sum())
    assert 0 < merged <= int(g_sel.sum())
    # 3. additions (:158-165): merge target of the file not reached -> binned candidates, then triangulated
    assert mg.merger_wants_additions(merged, len(meas), len(moving)) == (merged < mg.get("target_number_of_merges") and merged < len(meas))
    g_win = mg.merger_select_additions(meas)
    o_win = O.merger_select_additions(meas, o_occ, 376, 1241, rows_bins, col_bins, True, "stereo")
    assert np.array_equal(g_win, o_win) and len(g_win) > 0
    p_sel, p_win = mg.merger_plan(meas, moving, resp)
    assert np.array_equal(p_sel, o_sel) and np.array_equal(p_win, o_win)
    from srrg2_proslam_b200 import capi
    ctx = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
    try:
        g_xyz, g_valid, _ = ctx.triangulate(meas[g_win], K_KITTI, b_x, 0.0)
    finally:
        ctx.close()
    o_xyz, _ = O.triangulate(meas[o_win], K_KITTI, b_x, 0.0)
    assert np.array_equal(g_xyz, o_xyz) and g_valid.all()

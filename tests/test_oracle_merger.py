"""CPU restatement of the merger's binning (oracle/pslam_oracle_mapping.hpp, merger_projective_impl.cpp) against the
reference's own merger test constants (tests/test_mergers.cpp) and hand-made cases of its rules."""
import numpy as np

import oracle_lib as O
from merger_fixtures import icl_00_01, random_case

ICL = dict(canvas_rows=480, canvas_cols=640)  # image_rows / image_cols of the ICL fixture


def test_icl_00_to_00_adds_nothing(oracle):
    """tests/test_mergers.cpp:248-296 (00To00_MergerCorrespondenceProjectiveDepthEKF_Sparse): mirror correspondences with
    response 0 block every bin that holds a measurement, the addition pass finds no free bin: 321 -> 321"""
    m0, _, _ = icl_00_01()
    n = len(m0["uvd"])
    assert n == 321
    sel, occ = O.merger_select_updates(m0["uvd"], np.arange(n), np.zeros(n), max_distance_appearance=50, kind="depth", **ICL)
    assert 0 < sel.sum() < n  # binning: one update per bin
    assert len(O.merger_select_additions(m0["uvd"], occ, kind="depth", **ICL)) == 0


def test_icl_00_to_01_known_scene_size(oracle):
    """tests/test_mergers.cpp:298-357 (00To01_MergerCorrespondenceProjectiveDepthEKF_Sparse): scene 321 -> 337, i.e. 16
    measurements of frame 01 fall into bins that no gated correspondence blocks (all depths are valid)"""
    m0, m1, corr = icl_00_01()
    assert len(m0["uvd"]) == 321 and len(m1["uvd"]) == 338  # :329 and fixtures
    sel, occ = O.merger_select_updates(m1["uvd"], corr[:, 1], corr[:, 2], max_distance_appearance=50, kind="depth", **ICL)
    win = O.merger_select_additions(m1["uvd"], occ, kind="depth", **ICL)
    assert (m1["uvd"][win, 2] > 0).all()
    assert 321 + len(win) == 337  # ASSERT_EQ(points_in_camera_00.size(), 337)


def test_rules_by_hand(oracle):
    # canvas 100 x 100, 10 x 10 bins of 10 px: bin = round(coordinate / 10)
    meas = np.array([[12, 12, 10, 12],    # 0: bin (1,1), disparity 2
                     [14, 13, 4, 13],     # 1: bin (1,1), disparity 10
                     [9, 8, 3, 8],        # 2: bin (1,1) too (round(.9), round(.8)), disparity 6
                     [55, 55, 50, 55],    # 3: bin (6,6)  (round(5.5) = 6, half away from zero), disparity 5
                     [56, 57, 46, 57],    # 4: bin (6,6), disparity 10
                     [58, 58, 48, 58],    # 5: bin (6,6), disparity 10 (tie with 4: the earlier one stays)
                     [90, 20, 80, 20]],   # 6: bin (2,9)
                    np.float32)
    kw = dict(canvas_rows=100, canvas_cols=100, row_bins=10, col_bins=10)
    # correspondence 0 fails the gate (does not block), 1 blocks bin (1,1), 2 is skipped, 3 blocks (2,9)
    sel, occ = O.merger_select_updates(meas, [1, 0, 2, 6], [80, 50, 10, 0], max_distance_appearance=50, **kw)
    assert sel.tolist() == [False, True, False, True]
    bits = {b for b in range(121) if (occ[b >> 5] >> (b & 31)) & 1}
    assert bits == {1 * 11 + 1, 2 * 11 + 9}
    # additions: only bin (6,6) is free; occupant 3 is replaced by 4 (larger disparity), 5 ties and does not replace
    assert O.merger_select_additions(meas, occ, kind="stereo", **kw).tolist() == [4]
    assert O.merger_select_additions(meas, occ, kind="base", **kw).tolist() == [3]
    # nothing blocked: slots in order of each bin's first arrival
    assert O.merger_select_additions(meas, None, kind="stereo", **kw).tolist() == [1, 4, 6]
    # binning disabled: every gated correspondence is processed, every measurement is a candidate
    sel, occ = O.merger_select_updates(meas, [1, 0, 2, 6], [80, 50, 10, 0], max_distance_appearance=50, enable_binning=False, **kw)
    assert sel.tolist() == [False, True, True, True] and not occ.any()
    assert O.merger_select_additions(meas, occ, enable_binning=False, **kw).tolist() == list(range(7))
    # depth measurements: the smaller depth takes the bin
    uvd = np.array([[12, 12, 3.0], [14, 13, 2.0], [13, 13, 2.0], [70, 70, 1.0]], np.float32)
    assert O.merger_select_additions(uvd, None, kind="depth", **kw).tolist() == [1, 3]


def test_properties_random(oracle):
    for seed in range(4):
        meas, moving, resp = random_case(seed, 3000, 1500, crowded=seed % 2 == 1)
        sel, occ = O.merger_select_updates(meas, moving, resp, 376, 1241, max_distance_appearance=50)
        assert not sel[resp > 50].any()
        bins = np.round(meas[:, 1] / np.float32(37.6)).astype(int) * 31 + np.round(meas[:, 0] / (np.float32(1241) / np.float32(30))).astype(int)
        assert len(set(bins[moving[sel]])) == sel.sum()                      # one update per bin
        assert set(bins[moving[sel]]) == set(bins[moving[resp <= 50]])       # every gated bin got its update
        win = O.merger_select_additions(meas, occ, 376, 1241)
        assert len(set(bins[win])) == len(win) and not (set(bins[win]) & set(bins[moving[sel]]))
        free = set(bins) - set(bins[moving[sel]])
        assert set(bins[win]) == free
        disp = meas[:, 0] - meas[:, 2]
        for w in win:
            assert disp[w] == disp[bins == bins[w]].max()

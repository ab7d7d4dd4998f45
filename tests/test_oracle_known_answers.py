"""Pins the CPU oracle against the reference's own known-answer test constants.

Every expected value below is a literal from the reference's gtest files (cited per case);
inputs are the reference's test_data images (tests/golden/, made by tools/make_golden.py).
"""
import numpy as np
import pytest

import oracle_lib as O

K_KITTI = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)


@pytest.fixture(scope="module")
def imgs(oracle):
    g = O.load_gray
    return dict(
        L=[g(f"kitti_city_image_left_{i}.png") for i in range(5)],
        R=[g(f"kitti_city_image_right_{i}.png") for i in range(5)],
        I={i: g(f"icl_image_rgb_{i}.png") for i in (0, 1, 50)},
        D={i: g(f"icl_image_depth_{i}.png") for i in (0, 1, 50)},
        SL=g("scene_flow_image_left.png"), SR=g("scene_flow_image_right.png"))


def n_feat(img, **kw):
    return len(O.extract_binned(img, O.extract_cfg(**kw))["xy"])


def test_fast_matches_opencv(imgs):
    """FAST-9/16+NMS restatement is bit-exact against cv2 (coordinates, order, response)."""
    import cv2
    for img, thr in [(imgs["L"][0], 15), (imgs["L"][0], 5), (imgs["I"][0], 5), (imgs["SL"], 5)]:
        xy, r = O.fast_detect(img, thr)
        kp = cv2.FastFeatureDetector_create(thr, True).detect(img)
        assert np.array_equal(xy, np.array([[k.pt[0], k.pt[1]] for k in kp], np.float32))
        assert np.array_equal(r, np.array([k.response for k in kp], np.float32))
    xy, r = O.fast_detect(imgs["I"][0], 5, nms=False)
    kp = cv2.FastFeatureDetector_create(5, False).detect(imgs["I"][0])
    assert np.array_equal(xy, np.array([[k.pt[0], k.pt[1]] for k in kp], np.float32))


def test_feature_counts_kitti_1x1(imgs):  # tests/test_feature_extractors.cpp:22
    assert n_feat(imgs["L"][0], threshold=5, target=1000, nh=1, nv=1) == 887


def test_feature_counts_icl_1x1(imgs):  # tests/test_feature_extractors.cpp:111,115
    assert n_feat(imgs["I"][0], threshold=5, target=300, nh=1, nv=1) == 259
    assert n_feat(imgs["I"][1], threshold=5, target=300, nh=1, nv=1) == 254


def test_feature_counts_icl_3x3(imgs):  # tests/test_feature_extractors.cpp:131,135
    assert n_feat(imgs["I"][0], threshold=5, target=300) == 220
    assert n_feat(imgs["I"][1], threshold=5, target=300) == 228


def test_feature_counts_kitti_3x3(imgs):  # tests/test_feature_extractors.cpp:151-165
    got = [n_feat(x, threshold=5, target=300)
           for x in (imgs["L"][0], imgs["L"][1], imgs["R"][0], imgs["R"][1])]
    assert got == [272, 280, 270, 271]


def test_feature_counts_finder_fixtures(imgs):  # tests/test_correspondence_finders.cpp:25,62,116,203-204
    assert [n_feat(imgs["I"][i], threshold=5, target=500) for i in (0, 1, 50)] == [321, 338, 261]
    assert n_feat(imgs["L"][0], threshold=5, target=500) == 446
    assert n_feat(imgs["R"][0], threshold=5, target=500) == 444


def test_bruteforce_icl(imgs):  # tests/test_correspondence_finders.cpp:37-41,72,77-95,126
    f = {i: O.extract_binned(imgs["I"][i], O.extract_cfg(threshold=5, target=500)) for i in (0, 1, 50)}
    fi, mi, d = O.match_bruteforce(f[0]["desc"], f[0]["desc"], 50, 0.9)
    assert len(fi) == 319 and np.array_equal(fi, mi) and (d == 0).all()
    fi, mi, d = O.match_bruteforce(f[0]["desc"], f[1]["desc"], 50, 0.9)
    assert len(fi) == 226
    fi2, mi2, d2 = O.match_bruteforce(f[1]["desc"], f[0]["desc"], 50, 0.9)
    assert len(fi2) == 226
    assert set(zip(fi.tolist(), mi.tolist())) == set(zip(mi2.tolist(), fi2.tolist()))
    assert len(set(fi.tolist())) == len(fi) and len(set(mi.tolist())) == len(mi)  # bijective
    assert len(O.match_bruteforce(f[0]["desc"], f[50]["desc"], 50, 0.9)[0]) == 117


def test_kitti_matchers(imgs):  # tests/test_correspondence_finders.cpp:176-180,214,274,290
    c = O.extract_cfg(threshold=5, target=500)
    fl, fr = O.extract_binned(imgs["L"][0], c), O.extract_binned(imgs["R"][0], c)
    fi, mi, d = O.match_epipolar(fl["xy"], fl["desc"], fl["xy"], fl["desc"], 50, 0.9, 100, 0)
    assert len(fi) == len(fl["xy"]) == 446 and np.array_equal(fi, mi)
    assert len(O.match_bruteforce(fl["desc"], fr["desc"], 50, 0.9)[0]) == 237
    assert len(O.match_epipolar(fl["xy"], fl["desc"], fr["xy"], fr["desc"], 50, 0.9, 100, 0)[0]) == 150
    assert len(O.match_epipolar(fl["xy"], fl["desc"], fr["xy"], fr["desc"], 50, 0.9, 100, 1)[0]) == 241


def _scene_flow_inliers(uvuv):
    gt = {}
    for line in open(O.GOLDEN / "scene_flow_gt_stereo_matching_threshold-100.txt"):
        r, c, _, _, disp = line.split()
        gt[(int(r), int(c))] = float(disp)
    n = 0
    for u_l, v_l, u_r, _ in uvuv:  # tests/fixtures.hpp:513-535
        key = (int(v_l), int(u_l))
        if key in gt and abs(gt[key] - abs(int(u_l) - int(u_r))) < 1.0:
            n += 1
    return n


def test_adaptor_scene_flow(imgs):  # tests/test_measurement_adaptors.cpp:26,31,51,56
    c = O.extract_cfg(threshold=5, target=500)
    a = O.stereo_adaptor(imgs["SL"], imgs["SR"], c, "bruteforce", 100, 0.8)
    assert len(a["uvuv"]) == 83 and _scene_flow_inliers(a["uvuv"]) == 43
    b = O.stereo_adaptor(imgs["SL"], imgs["SR"], c, "epipolar", 100, 0.8)
    assert len(b["uvuv"]) == 115 and _scene_flow_inliers(b["uvuv"]) == 59


def test_adaptor_kitti(imgs):  # tests/test_measurement_adaptors.cpp:110,130
    c = O.extract_cfg(threshold=5, target=500)
    assert len(O.stereo_adaptor(imgs["L"][0], imgs["R"][0], c, "bruteforce", 100, 0.8)["uvuv"]) == 213
    assert len(O.stereo_adaptor(imgs["L"][0], imgs["R"][0], c, "epipolar", 100, 0.8)["uvuv"]) == 177


def test_adaptor_icl_mono_depth(imgs):  # tests/test_measurement_adaptors.cpp:75-88
    m = O.mono_depth_adaptor(imgs["I"][0], imgs["D"][0], O.extract_cfg(threshold=5, target=500), 1.0)
    assert len(m["uvd"]) == 321
    for (u, v, d), inten in zip(m["uvd"], m["intensity"]):
        r, c = int(np.rint(v)), int(np.rint(u))
        assert inten == imgs["I"][0][r, c] and d == imgs["D"][0][r, c] and d > 0


def kitti_pose(R, t):
    return np.concatenate([np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3, 1)],
                          axis=1).reshape(12)


CAM00 = kitti_pose([1.0, 9.043680e-12, 2.326809e-11, 9.043683e-12, 1.0, 2.392370e-10, 2.326810e-11,
                    2.392370e-10, 9.999999e-01], [5.551115e-17, 3.330669e-16, -4.440892e-16])
CAM01 = kitti_pose([9.999978e-01, 5.272628e-04, -2.066935e-03, -5.296506e-04, 9.999992e-01,
                    -1.154865e-03, 2.066324e-03, 1.155958e-03, 9.999971e-01],
                   [-4.690294e-02, -2.839928e-02, 8.586941e-01])  # tests/fixtures.hpp:884-892


def kitti_fixture_chain(imgs):
    """tests/fixtures.hpp:832-846,926-952: adaptor (thr 15, target 500, epipolar 50/0.8) on frame 0,
    triangulate (min disparity 0), moving cloud for the projective finders."""
    c = O.extract_cfg(threshold=15, target=500)
    m0 = O.stereo_adaptor(imgs["L"][0], imgs["R"][0], c, "epipolar", 50, 0.8)
    xyz, ninv = O.triangulate(m0["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
    assert ninv == 0
    return m0, xyz


def test_projective_circle_kitti(imgs):  # tests/test_correspondence_finders.cpp:472-509
    m0, xyz = kitti_fixture_chain(imgs)
    assert len(m0["uvuv"]) == 145
    f1 = O.extract_binned(imgs["L"][1], O.extract_cfg(threshold=5, target=500))
    assert len(f1["xy"]) == 458
    cam01_in_00 = O.pose_mul(O.pose_inverse(CAM00), CAM01)
    pf = O.ProjectiveFinder(K_KITTI, 376, 1241, shape="circle", max_desc_dist=50, ratio=0.9,
                            min_desc_dist=50, max_radius=10, min_radius=10)
    pf.set_fixed(f1["xy"], f1["desc"])
    pf.set_moving(xyz, m0["desc"])
    # finder is given local_map_in_sensor = camera_01_in_00^-1 (points of 00 seen from 01)
    pf.set_estimate(O.pose_inverse(cam01_in_00))
    fi, mi, d = pf.compute()
    assert len(fi) == 90

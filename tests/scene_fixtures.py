"""Clouds and camera poses of the reference's scene-clipper tests (tests/test_scene_clippers.cpp, fixtures built at
tests/fixtures.hpp:621-656 (ICL) and :926-952 (KITTI)).  Shared by the CPU known-answer test and the GPU parity test."""
import numpy as np

import oracle_lib as O

K_ICL = np.array([481.2, 0, 319.5, 0, -481, 239.5, 0, 0, 1], np.float32)  # tests/fixtures.hpp:577
K_KITTI = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)


def unproject(uvd, K):
    """srrg2_core PointUnprojectorPinhole_ (un-vendored): p = K^-1 * (u d, v d, d), fp32"""
    Ki = np.linalg.inv(K.reshape(3, 3).astype(np.float64)).astype(np.float32)
    d = uvd[:, 2]
    ud, vd = uvd[:, 0] * d, uvd[:, 1] * d
    x = (Ki[0, 0] * ud + Ki[0, 1] * vd) + Ki[0, 2] * d
    y = (Ki[1, 0] * ud + Ki[1, 1] * vd) + Ki[1, 2] * d
    return np.stack([x, y, d], 1).astype(np.float32)


def icl_depth_meters():
    """tests/fixtures.hpp:737-740: 16-bit depth image converted to float with scale 1e-3"""
    return (O.load_depth("icl_image_depth_0.png").astype(np.float32) * np.float32(1e-3)).astype(np.float32)


def icl_sparse():
    """321 adaptor measurements of frame 0 (thr 5, target 500, depth scale 1: tests/fixtures.hpp:567-571), unprojected"""
    m = O.mono_depth_adaptor(O.load_gray("icl_image_rgb_0.png"), icl_depth_meters(), O.extract_cfg(threshold=5, target=500), 1.0)
    return unproject(m["uvd"], K_ICL), m["desc"]


def icl_dense():
    """all 640x480 pixels of the depth image, unprojected (tests/fixtures.hpp:651-656)"""
    d = icl_depth_meters()
    v, u = np.mgrid[0:480, 0:640]
    return unproject(np.stack([u.ravel(), v.ravel(), d.ravel()], 1).astype(np.float32), K_ICL)


def kitti_sparse():
    """145 triangulated stereo points of frame 0 (tests/fixtures.hpp:832-846,926-944)"""
    c = O.extract_cfg(threshold=15, target=500)
    m0 = O.stereo_adaptor(O.load_gray("kitti_city_image_left_0.png"), O.load_gray("kitti_city_image_right_0.png"), c, "epipolar", 50, 0.8)
    xyz, _ = O.triangulate(m0["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
    return xyz, m0["desc"]


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def pose(R=None, t=(0, 0, 0)):
    R = np.eye(3) if R is None else R
    return np.concatenate([R, np.asarray(t, np.float64).reshape(3, 1)], 1).astype(np.float32).reshape(12)


# robot_in_local_map of every case (sensor_in_robot = identity) with the reference's expected survivor counts:
# (name, pose, sparse ICL count, dense ICL count) -- tests/test_scene_clippers.cpp:7-391
ICL_CASES = [
    ("no_motion", pose(), 321, None),  # dense: ASSERT_LE(.., 306671) depends on the external dense unprojector's rounding
    ("full_pitch", pose(rot_z(np.pi)), 321, None),
    ("full_roll", pose(rot_x(np.pi)), 0, 0),
    ("quarter_roll", pose(rot_x(np.pi / 4)), 51, 49872),
    ("translate_backward", pose(t=(0, 0, -1)), 321, 307200),
    ("translate_forward", pose(t=(0, 0, 1)), 242, 136022),
]
# tests/test_scene_clippers.cpp:393-460
KITTI_CASES = [("no_motion", pose(), 145), ("translate_forward", pose(t=(0, 0, 10)), 52)]

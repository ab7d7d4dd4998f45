"""Pins the oracle's scene clipper (= the pinhole projector over a whole cloud) against the literal
survivor counts of the reference's tests/test_scene_clippers.cpp."""
import numpy as np
import pytest

import oracle_lib as O
import scene_fixtures as F


@pytest.fixture(scope="module")
def clouds(oracle):
    return {"icl_sparse": F.icl_sparse()[0], "icl_dense": F.icl_dense(), "kitti": F.kitti_sparse()[0]}


@pytest.mark.parametrize("name,T,n_sparse,n_dense", F.ICL_CASES, ids=[c[0] for c in F.ICL_CASES])
def test_icl_known_answers(clouds, name, T, n_sparse, n_dense):
    assert len(clouds["icl_sparse"]) == 321 and len(clouds["icl_dense"]) == 307200
    xyz, uvz, idx = O.scene_clip(clouds["icl_sparse"], T, F.K_ICL, 480, 640, 0.1, 10.0)
    assert len(idx) == n_sparse
    assert np.all(xyz[:, 2] > 0)  # "points should still lie ahead of the camera"
    if n_dense is not None:
        xyz, uvz, idx = O.scene_clip(clouds["icl_dense"], T, F.K_ICL, 480, 640, 0.1, 10.0)
        assert len(idx) == n_dense
        assert np.all(np.diff(idx) > 0) and np.all(xyz[:, 2] > 0)


@pytest.mark.parametrize("name,T,n", F.KITTI_CASES, ids=[c[0] for c in F.KITTI_CASES])
def test_kitti_known_answers(clouds, name, T, n):
    assert len(clouds["kitti"]) == 145
    xyz, uvz, idx = O.scene_clip(clouds["kitti"], T, F.K_KITTI, 376, 1241, 0.1, 1000.0)
    assert len(idx) == n and np.all(xyz[:, 2] > 0)


def test_sensor_in_robot_and_order(clouds):
    """survivors keep the map order, indices map local -> global, sensor_in_robot moves them into the robot frame"""
    S = F.pose(F.rot_z(0.3), (0.1, -0.2, 0.05))
    a = O.scene_clip(clouds["icl_sparse"], F.pose(t=(0, 0, 1)), F.K_ICL, 480, 640, 0.1, 10.0)
    b = O.scene_clip(clouds["icl_sparse"], F.pose(t=(0, 0, 1)), F.K_ICL, 480, 640, 0.1, 10.0, sensor_in_robot=S)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
    R, t = S.reshape(3, 4)[:, :3], S.reshape(3, 4)[:, 3]
    assert np.allclose(b[0], a[0] @ R.T + t, atol=1e-5)
    assert np.allclose(a[0][:, 2], a[1][:, 2])  # depth of the projection = z in the camera frame

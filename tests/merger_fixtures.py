"""Inputs of the reference's merger tests (tests/test_mergers.cpp) rebuilt from the committed ICL / KITTI images:
measurement clouds of the adaptors and the fixture's "ideal" correspondences (tests/fixtures.hpp:658-705).  Shared by
the CPU known-answer test and the GPU parity test."""
import numpy as np

import oracle_lib as O
from scene_fixtures import K_ICL, icl_depth_meters, unproject


def quat_to_R(w, x, y, z):
    """Eigen::Quaternionf(w, x, y, z).toRotationMatrix()"""
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)


def icl_measurements(i):
    """RawDataPreprocessorMonocularDepth on ICL frame i (thr 5, target 500, depth scale 1: tests/fixtures.hpp:567-574)"""
    depth = (O.load_depth(f"icl_image_depth_{i}.png").astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    return O.mono_depth_adaptor(O.load_gray(f"icl_image_rgb_{i}.png"), depth, O.extract_cfg(threshold=5, target=500), 1.0)


def icl_00_01():
    """(measurements 00, measurements 01, correspondences_camera_01_from_00 as (fixed 00, moving 01, response))"""
    m0, m1 = icl_measurements(0), icl_measurements(1)
    p0, p1 = unproject(m0["uvd"], K_ICL).astype(np.float64), unproject(m1["uvd"], K_ICL).astype(np.float64)
    # ground truth (tests/fixtures.hpp:596-613)
    R0, t0 = quat_to_R(1, 0, 0, 0), np.array([0, 0, -2.25])
    R1, t1 = quat_to_R(0.999999, -0.00101358, 0.00052453, -0.000231475), np.array([0.000466347, 0.00895357, -2.24935])
    R01, t01 = R0.T @ R1, R0.T @ (t1 - t0)  # camera_01_in_00
    p0_in_01 = (p0 - t01) @ R01             # camera_01_in_00^-1 * p  (:669-671)
    d0 = np.unpackbits(m0["desc"], axis=1).astype(np.int32)
    d1 = np.unpackbits(m1["desc"], axis=1).astype(np.int32)
    used, corr = set(), []
    for i in range(len(p0)):  # :673-705: greedy, closest in appearance AND geometry, bijective
        best_a, best_g, best = 100.0, 0.1, -1
        ham = (d0[i][None, :] != d1).sum(1)
        geo = ((p1 - p0_in_01[i]) ** 2).sum(1)
        for j in range(len(p1)):
            if j in used:
                continue
            if ham[j] < best_a and geo[j] < best_g:
                best_a, best_g, best = float(ham[j]), float(geo[j]), j
        if best != -1:
            corr.append((i, best, best_a))
            used.add(best)
    return m0, m1, np.array(corr, np.float64)


def random_case(seed, n_meas, n_corr, rows=376, cols=1241, dim=4, crowded=False):
    """measurements inside the canvas, correspondences with bijective moving indices and responses around the gate"""
    r = np.random.default_rng(seed)
    span_u, span_v = (cols * (0.2 if crowded else 1.0)), (rows * (0.2 if crowded else 1.0))
    u = r.uniform(0, span_u - 1, n_meas).astype(np.float32)
    v = r.uniform(0, span_v - 1, n_meas).astype(np.float32)
    if dim == 4:
        disp = np.round(r.uniform(0, 60, n_meas)).astype(np.float32)  # integer disparities: ties between candidates
        meas = np.stack([u, v, np.maximum(u - disp, 0).astype(np.float32), v], 1).astype(np.float32)
    else:
        depth = (np.round(r.uniform(0.5, 10, n_meas) * 4) / 4).astype(np.float32)
        meas = np.stack([u, v, depth], 1).astype(np.float32)
    moving = r.permutation(n_meas)[:n_corr].astype(np.int32)
    response = np.round(r.uniform(0, 100, n_corr)).astype(np.float32)
    return meas, moving, response

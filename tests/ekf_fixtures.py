"""The synthetic scenario of the reference's filter tests (tests/fixtures.hpp:92-300 `Synthetic::generateContinousTransitions`,
camera fx = fy = 450, 600 x 400, stereo baseline 250 px): one point seen through a chain of small camera motions."""
import numpy as np

FX = FY = 450.0
COLS, ROWS = 600.0, 400.0
K = np.array([[FX, 0, COLS * 0.5], [0, FY, ROWS * 0.5], [0, 0, 1]])
BASELINE = np.array([250.0, 0.0, 0.0])  # pixels (tests/fixtures.hpp:328)
CAM6 = np.array([FX, FY, COLS * 0.5, ROWS * 0.5, BASELINE[0], BASELINE[1]])


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    R = np.eye(3)
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def project(p, baseline=np.zeros(3)):
    h = K @ p - baseline
    return (h / h[2])[:2]


def transitions(kind, n=100, sd_motion=0.0, sd_meas=0.0, seed=0):
    """kind: "translation" | "rotation" | "transform".  Returns ground-truth point positions [n+1][3], noisy camera
    transitions (world_in_sensor, [n][3][4]) and noisy measurements: mono [n+1][2], depth [n+1][3], stereo [n+1][4]."""
    rng = np.random.default_rng(seed)
    p = np.array([0.0, 0.0, np.sqrt(n)])
    gt, T_noisy = [p.copy()], []
    mono, depth, stereo = [], [], []

    def measure(p):
        uv, uvr = project(p), project(p, BASELINE)
        uvn = uv + rng.normal(0, sd_meas, 2) if sd_meas else uv
        uvrn = uvr + rng.normal(0, sd_meas, 2) if sd_meas else uvr
        mono.append(uvn)
        depth.append(np.array([uv[0], uv[1], p[2]]) + (rng.normal(0, sd_meas, 3) if sd_meas else 0))
        stereo.append(np.concatenate([uvn, uvrn]))

    measure(p)
    acc = np.zeros(3)
    for _ in range(n):
        R, t = np.eye(3), np.zeros(3)
        if kind in ("rotation", "transform"):
            a = rng.uniform(-1, 1, 3) * np.pi / 36.0
            a = np.where(np.abs(acc + a) > np.pi / 4, -a, a)
            R = rot(0, a[0]) @ rot(1, a[1]) @ rot(2, a[2])
            acc += a
        if kind in ("translation", "transform"):
            t = rng.uniform(-1, 1, 3) / 10.0
        q = R @ p + t
        if q[2] <= 0:
            t[2], R = -t[2], R.T
            q = R @ p + t
        Rn, tn = R.copy(), t.copy()
        if sd_motion:
            if kind in ("rotation", "transform"):
                a = rng.normal(0, sd_motion, 3)
                Rn = Rn @ rot(0, a[0]) @ rot(1, a[1]) @ rot(2, a[2])
            if kind in ("translation", "transform"):
                tn = tn + rng.normal(0, sd_motion, 3)
        T_noisy.append(np.concatenate([Rn, tn.reshape(3, 1)], 1))
        p = q
        gt.append(p.copy())
        measure(p)
    return np.array(gt), np.array(T_noisy), np.array(mono), np.array(depth), np.array(stereo)


# ---- the "LandmarkWorldNoNoise" scenario of tests/test_landmark_estimators.cpp:262-330 -----------------------------------
K_WORLD = np.array([200, 0, 100, 0, 200, 100, 0, 0, 1], np.float32)  # projection_matrix (:268), canvas 200 x 200 (:269)
BASELINE_WORLD_M = 0.5                                                # offset_second_sensor (:274)


def landmark_world(n_points=1000, n_poses=10, seed=0):
    """poses: sensor_in_world [n_poses][3][4] from the identity to a translation of (10, 10, 10) with a small deviation
    (:277-281); points around (10, 10, 10) +- 10 (:283-285); per pose the visible points with their exact stereo
    measurement (uL, vL, uR, vR) and their coordinates in the sensor frame."""
    rng = np.random.default_rng(seed)
    poses = []
    for k in range(n_poses):
        t = np.full(3, 10.0 * k / (n_poses - 1)) + (rng.normal(0, 0.01, 3) if k else 0)
        poses.append(np.concatenate([np.eye(3), t.reshape(3, 1)], 1))
    poses = np.array(poses, np.float32)
    pts = (np.array([10, 10, 10]) + rng.uniform(-10, 10, (n_points, 3)) + np.array([0, 0, 12.0])).astype(np.float32)
    K = K_WORLD.reshape(3, 3).astype(np.float64)
    frames = []
    for T in poses.astype(np.float64):
        pc = (pts.astype(np.float64) - T[:, 3]) @ T[:, :3]
        h = pc @ K.T
        uv = h[:, :2] / h[:, 2:3]
        vis = (pc[:, 2] > 0.5) & (uv[:, 0] >= 0) & (uv[:, 0] <= 200) & (uv[:, 1] >= 0) & (uv[:, 1] <= 200)
        uvr = uv.copy()
        uvr[:, 0] -= K[0, 0] * BASELINE_WORLD_M / pc[:, 2]
        frames.append(dict(visible=np.flatnonzero(vis), stereo=np.concatenate([uv, uvr], 1).astype(np.float32),
                           in_sensor=pc.astype(np.float32)))
    return poses, pts, frames


def unproject_world(uv, depth):
    """LandmarkWorldNoNoise::getPointUnprojected (tests/test_landmark_estimators.cpp:349-358)"""
    fx, fy, cx, cy = K_WORLD[0], K_WORLD[4], K_WORLD[2], K_WORLD[5]
    d = depth.astype(np.float32)
    return np.stack([d / fx * (uv[:, 0] - cx), d / fy * (uv[:, 1] - cy), d], 1).astype(np.float32)


def run_smoother_scenario(update, n_points=1000, n_poses=10, seed=0):
    """drives `update(K, frames_sensor_in_world, sensor_in_world, sensor_in_local_map, offsets, hist_frame, hist_uv,
    hist_point_in_camera, state_world, n_opt)` like the reference test drives the estimator (:210-258): landmarks seeded
    from the first frame with 1e-4 noise, every later frame re-observes its visible landmarks.  Returns (state, truth)."""
    rng = np.random.default_rng(seed + 100)
    poses, pts, frames = landmark_world(n_points, n_poses, seed)
    ids = frames[0]["visible"]
    lm_of = {int(p): i for i, p in enumerate(ids)}
    in0 = frames[0]["in_sensor"][ids] + (1e-4 * rng.uniform(-1, 1, (len(ids), 3))).astype(np.float32)
    state = (in0 @ poses[0][:, :3].T + poses[0][:, 3]).astype(np.float32)
    n_opt = np.zeros(len(ids), np.int32)
    hist = [[(0, frames[0]["stereo"][p, :2], frames[0]["in_sensor"][p])] for p in ids]
    for f in range(1, n_poses):
        sel = [lm_of[int(p)] for p in frames[f]["visible"] if int(p) in lm_of]
        if not sel:
            continue
        pidx = ids[sel]
        uv = frames[f]["stereo"][pidx, :2]
        pic = unproject_world(uv, frames[f]["in_sensor"][pidx, 2])
        for j, i in enumerate(sel):
            hist[i].append((f, uv[j], pic[j]))
        off = np.zeros(len(sel) + 1, np.int32)
        hf, huv, hpc = [], [], []
        for j, i in enumerate(sel):
            off[j + 1] = off[j] + len(hist[i])
            for (ff, a, b) in hist[i]:
                hf.append(ff)
                huv.append(a)
                hpc.append(b)
        st, no, loc, inl = update(K_WORLD, poses[:f + 1], poses[f], poses[f], off, np.array(hf, np.int32), np.array(huv, np.float32),
                                  np.array(hpc, np.float32), state[sel], n_opt[sel])
        state[sel], n_opt[sel] = st, no
    return state, pts[ids], n_opt

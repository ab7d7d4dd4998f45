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

"""per-iteration cost of the fused solver loop (gn_iterate_kernel) on a frame-sized problem: slope of the call time over the
iteration count, so uploads / launch / download drop out"""
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from srrg2_proslam_b200 import capi

K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float64)


def problem(n, seed=3):
    rng = np.random.default_rng(seed)
    xyz = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(3, 40, n)], 1)
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    pc = xyz @ R.T + np.array([0.05, -0.02, 0.4])
    h = pc @ K.reshape(3, 3).T
    u, v = h[:, 0] / h[:, 2], h[:, 1] / h[:, 2]
    noise = rng.normal(0, 0.5, (n, 3))
    noise[rng.random(n) < 0.1] *= 40
    meas = np.stack([u + noise[:, 0], v + noise[:, 1], (h[:, 0] - 386.1448) / h[:, 2] + noise[:, 2], v], 1)
    cf = np.arange(n, dtype=np.int32)
    info = np.tile([1, 2, 1], (n, 1)) * rng.uniform(0.5, 3.0, (n, 1))
    return xyz.astype(np.float32), meas.astype(np.float32), cf, cf.copy(), info.astype(np.float32), np.eye(3, 4).reshape(12)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 72
    c = capi.Context(max_images=2, max_rows=64, max_cols=128, max_features=256, max_raw_per_bin=1024)
    xyz, meas, cf, cm, info, pose = problem(n)
    cfg = c.linearize_cfg("stereo", K, 1241, 376, (-386.1448, 0, 0), 30.0, "saturated", 1000.0)
    prior = (np.eye(3, 4).reshape(12), np.eye(6))
    res = {}
    for iters in (50, 450):
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            out = c.gn_iterate_f32(cfg, iters, 1.0, pose, xyz, meas, cf, cm, info, prior=prior)
            ts.append(time.perf_counter() - t0)
        assert out[3] == iters, out[3]
        res[iters] = float(np.median(ts[5:]))
    print(f"dbg={os.environ.get('PSLAM_GN_DBG', '0')} n={n}: {1e6 * (res[450] - res[50]) / 400:.3f} us / iteration (calls: {1e6 * res[50]:.0f} us @50, {1e6 * res[450]:.0f} us @450)")


main()

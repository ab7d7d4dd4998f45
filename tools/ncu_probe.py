"""Single-kernel workloads for `ncu --set full` captures (never a bench number: ncu serialises and replays).
  python tools/ncu_probe.py hamming    -> bf_sweep_kernel<0> on BASELINE config 5 (64k x 64k 256-bit descriptors)
  python tools/ncu_probe.py linearize  -> linearize_kernel on 2^22 synthetic stereo correspondences
  python tools/ncu_probe.py match      -> pslam_match_bruteforce on config 5 (max_dist 50, ratio 0.9): sweep + resolve
"""
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from srrg2_proslam_b200 import capi, synth
    what = sys.argv[1] if len(sys.argv) > 1 else "hamming"
    dev = torch.device("cuda", 0)
    ctx = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=4096, max_raw_per_bin=8192)
    if what in ("hamming", "match"):
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
        q, t = synth.hamming_sets(n, n, seed=0)
        if what == "hamming":
            dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
            ob = torch.empty((3, n), dtype=torch.int32, device=dev)
            for _ in range(2):
                ctx.bf_best2_dev(n, dq.data_ptr(), n, dt_.data_ptr(), ob[0].data_ptr(), ob[1].data_ptr(), ob[2].data_ptr())
            ctx.synchronize()
            print("best mean", float(ob[0].float().mean()))
        else:
            t0 = time.perf_counter()
            fi, mi, d = ctx.match_bruteforce(q, t, capi.match_cfg(50, 0.9))
            print("matches", len(fi), "s", time.perf_counter() - t0)
    elif what == "natural":
        # stage 1 + stereo match on the five real KITTI pairs of tests/golden tiled to 384 pairs (one 768-image chunk)
        import cv2
        G = ROOT / "tests" / "golden"
        pairs = [np.stack([cv2.imread(str(G / f"kitti_city_image_{s_}_{i}.png"), cv2.IMREAD_UNCHANGED) for s_ in ("left", "right")])
                 for i in range(5)]
        P = 385
        batch = torch.from_numpy(np.stack(pairs)).to(dev).repeat(77, 1, 1, 1).contiguous()
        big = capi.Context(device=0, max_images=2 * P, max_rows=376, max_cols=1241, max_features=4096, max_raw_per_bin=8192,
                           max_bins=9, work_images=768)
        for _ in range(3):
            big.stereo_frontend_batch_dev(batch.data_ptr(), P, 376, 1241, 1241, 376 * 1241, capi.extract_cfg(15, 1, 4000),
                                          capi.match_cfg(100.0, 0.5, 100, 0))
        big.synchronize()
        print("stereo points per frame", float(big.stereo_counts(P).mean()))
        big.close()
    elif what == "linearize":
        n = 1 << 22
        K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)
        rng = np.random.default_rng(0)
        xyz = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(3, 40, n)], 1)
        h = xyz @ K.reshape(3, 3).astype(np.float64).T
        meas = np.stack([h[:, 0] / h[:, 2], h[:, 1] / h[:, 2], (h[:, 0] - 386.1448) / h[:, 2], h[:, 1] / h[:, 2]], 1)
        idx = np.arange(n, dtype=np.int32)
        info = np.tile([1.0, 2.0, 1.0], (n, 1))
        cfg = ctx.linearize_cfg("stereo", K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
        print("ms", ctx.linearize_timed(cfg, np.eye(3, 4).reshape(12), xyz, meas, idx, idx, info, reps=3))
    ctx.close()


if __name__ == "__main__":
    main()

"""Per-stage single-thread timing of the CPU oracle next to cv2 on the host this runs on (context for cpu_baseline)."""
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cv2  # noqa: E402
import oracle_lib as O  # noqa: E402
from srrg2_proslam_b200 import synth  # noqa: E402

imgs = synth.stereo_pairs(4, 376, 1241, seed=0).numpy()
L, R = imgs[0, 0], imgs[0, 1]
cfg = O.extract_cfg(15, 1, 4000)


def tm(f, n=5):
    f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = f()
        t.append(time.perf_counter() - t0)
    return 1e3 * min(t), r


print("oracle fast_detect ms", tm(lambda: O.fast_detect(L, 15))[0])
print("oracle blur7 ms", tm(lambda: O.blur7(L))[0])
print("oracle detect_binned ms", tm(lambda: O.detect_binned(L, cfg))[0])
t, f = tm(lambda: O.extract_binned(L, cfg))
print("oracle extract_binned ms", t, len(f["xy"]))
f2 = O.extract_binned(R, cfg)
print("oracle epipolar ms", tm(lambda: O.match_epipolar(f["xy"], f["desc"], f2["xy"], f2["desc"], 100, 0.5, 100, 0))[0])
print("oracle pair ms", tm(lambda: O.stereo_frontend_batch(imgs[:2], cfg, threads=1, max_dist=100, ratio=0.5, max_disp=100, thickness=0))[0] / 2)
cv2.setNumThreads(1)
fast = cv2.FastFeatureDetector_create(15, True)
t, k = tm(lambda: fast.detect(L))
print("cv2 FAST ms", t, len(k))
kk = sorted(k, key=lambda x: -x.response)[:3244]
print("cv2 ORB.compute ms", tm(lambda: cv2.ORB_create().compute(L, kk))[0])
print("cv2 GaussianBlur 7x7 ms", tm(lambda: cv2.GaussianBlur(L, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))[0])

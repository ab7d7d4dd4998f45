"""per-kernel device times of ONE stereo adaptor call (per-frame latency path)"""
import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, cv2
from srrg2_proslam_b200 import capi
G = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"
L = cv2.imread(str(G / "kitti_city_image_left_1.png"), cv2.IMREAD_UNCHANGED)
R = cv2.imread(str(G / "kitti_city_image_right_1.png"), cv2.IMREAD_UNCHANGED)
ctx = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
e, m = capi.extract_cfg(15, 1, 1000), capi.match_cfg(100.0, 0.5, 100, 0)
for _ in range(5):
    ctx.stereo_adaptor(L, R, e, m)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); ctx.stereo_adaptor(L, R, e, m); ts.append(time.perf_counter() - t0)
print("wall us per call (median):", 1e6 * float(np.median(ts)))
ctx.profile_enable(True)
for _ in range(10):
    ctx.stereo_adaptor(L, R, e, m)
prof = ctx.profile_read()
tot = 0
for k, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {1e3 * t / c:8.1f} us x{c // 10}/call")
    tot += 1e3 * t / 10
print("kernel (event-to-event) us per call:", tot)
ctx.close()

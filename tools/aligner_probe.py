"""per-frame latency of the conf-driven aligner (kitti.conf, KITTI 00 -> 01 of tests/golden): wall clock per compute() and the
device time of every kernel it launches (pslam_profile_*), to see what the latency is made of"""
import ctypes as C
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cv2
from srrg2_proslam_b200 import capi, plugin as P

G = ROOT / "tests" / "golden"
K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)
ld = lambda n: cv2.imread(str(G / n), cv2.IMREAD_UNCHANGED)
small = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
e, mcfg = capi.extract_cfg(15, 1, 500), capi.match_cfg(50, 0.8, 100, 0)
meas = [small.stereo_adaptor(ld(f"kitti_city_image_left_{i}.png"), ld(f"kitti_city_image_right_{i}.png"), e, mcfg) for i in (0, 1)]
xyz, _, _ = small.triangulate(meas[0]["uvuv"], K, float(np.float32(718.856) * np.float32(0.537166)), 0.0)
small.close()
m = P.Manager(G / "configurations" / "kitti_hotpath.conf")
al = m.get("aligner")
sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
pr = sl.link("projector")
pr.set_camera_matrix(K)
pr.set("canvas_rows", 376).set("canvas_cols", 1241)
al.aligner_set_left_camera_in_right([-0.537166, 0, 0])


def run():
    al.aligner_set_fixed(meas[1]["uvuv"], meas[1]["desc"])
    al.aligner_set_moving(xyz, meas[0]["desc"])
    al.aligner_set_moving_in_fixed(np.eye(3, 4, dtype=np.float32))
    t0 = time.perf_counter()
    r = al.aligner_run()
    return time.perf_counter() - t0, r


ts = [run()[0] for _ in range(12)]
print(f"aligner wall: median {1e6 * float(np.median(ts[2:])):.0f} us, min {1e6 * min(ts):.0f} us")
P.lib().psp_profile_enable(1)
reps = 5
for _ in range(reps):
    _, r = run()
cap, ln = 48, 64
names = C.create_string_buffer(cap * ln)
ms, cnt = (C.c_double * cap)(), (C.c_longlong * cap)()
n = P.lib().psp_profile_read(cap, names, ln, ms, cnt)
tot = 0.0
for i in range(n):
    nm = names.raw[i * ln:(i + 1) * ln].split(b"\0", 1)[0].decode()
    print(f"  {nm:32s} {1e3 * ms[i] / cnt[i]:7.1f} us x {cnt[i] / reps:5.1f} / frame = {1e3 * ms[i] / reps:7.1f} us")
    tot += 1e3 * ms[i] / reps
print(f"kernel (event-to-event) time per frame: {tot:.0f} us; iterations {r['iterations']}, correspondences {r['num_correspondences']}")
P.lib().psp_profile_enable(0)

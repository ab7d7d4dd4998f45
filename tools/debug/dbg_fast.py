import sys, pathlib, collections
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O
from srrg2_proslam_b200 import capi
img = O.load_gray("kitti_city_image_left_0.png")
ctx = capi.Context(max_images=2, max_rows=376, max_cols=1241, max_features=4096, max_raw_per_bin=200000, max_bins=1)
for nms in (1, 0):
    xy, r = ctx.fast_detect(img, 15, nms)
    oxy, orr = O.fast_detect(img, 15, nms)
    g = {(int(a), int(b)): float(c) for (a, b), c in zip(xy, r)}
    o = {(int(a), int(b)): float(c) for (a, b), c in zip(oxy, orr)}
    extra = sorted(set(g) - set(o), key=lambda p: (p[1], p[0]))
    missing = sorted(set(o) - set(g), key=lambda p: (p[1], p[0]))
    wrong = [p for p in g if p in o and g[p] != o[p]]
    print("nms", nms, "gpu", len(g), "oracle", len(o), "extra", len(extra), "missing", len(missing), "wrong response", len(wrong), "dups", len(xy) - len(g))
    print(" extra x%8", collections.Counter(p[0] % 8 for p in extra).most_common())
    print(" extra y%8", collections.Counter(p[1] % 8 for p in extra).most_common())
    print(" extra (x-2)%252", collections.Counter((p[0] - 2) % 252 for p in extra).most_common(6))
    print(" first extra", extra[:12])
    print(" first missing", missing[:12])
ctx.close()

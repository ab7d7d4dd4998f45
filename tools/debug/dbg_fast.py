import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_lib as O
from srrg2_proslam_b200 import capi
O.build()
ctx = capi.Context(max_images=2, max_rows=600, max_cols=1300, max_features=8192, max_raw_per_bin=40000)
img = O.load_gray("icl_image_rgb_0.png")
for thr, nms in ((15, 0), (15, 1)):
    xy, r = ctx.fast_detect(img, thr, nms)
    oxy, orr = O.fast_detect(img, thr, nms)
    g = {(int(a), int(b)): float(c) for (a, b), c in zip(xy, r)}
    o = {(int(a), int(b)): float(c) for (a, b), c in zip(oxy, orr)}
    print("thr", thr, "nms", nms, "gpu", len(g), "oracle", len(o), "common", len(set(g) & set(o)))
    go = sorted(set(g) - set(o), key=lambda p: (p[1], p[0]))[:15]
    og = sorted(set(o) - set(g), key=lambda p: (p[1], p[0]))[:15]
    print(" gpu-only", [(p, g[p]) for p in go])
    print(" oracle-only", [(p, o[p]) for p in og])
    diff = [(p, g[p], o[p]) for p in set(g) & set(o) if g[p] != o[p]][:10]
    print(" resp diff", diff)

"""Runs pslam_scene_clip_dev on a synthetic device-resident map (profiling target: ncu -k regex:scene_)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from srrg2_proslam_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
xyz = (torch.rand((n, 3), generator=g, device=dev) * 60 - 30).contiguous()
desc = torch.randint(0, 2 ** 31 - 1, (n, 8), generator=g, device=dev, dtype=torch.int32)
oxyz, ouvz = torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev)
oidx = torch.empty(n, dtype=torch.int32, device=dev)
odesc = torch.empty((n, 8), dtype=torch.int32, device=dev)
K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)
T = np.array([1, 0, 0, 0.3, 0, 1, 0, -0.1, 0, 0, 1, 1.5], np.float32)
ctx = capi.Context(device=0, max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
cfg = capi.clip_cfg(K, 376, 1241, T, 0.1, 1000.0)
torch.cuda.synchronize()
args = (n, xyz.data_ptr(), desc.data_ptr(), cfg, oxyz.data_ptr(), ouvz.data_ptr(), oidx.data_ptr(), odesc.data_ptr())
ctx.scene_clip_dev(*args, reps=3)
ctx.profile_enable(True)
kept, ms = ctx.scene_clip_dev(*args, reps=reps)
prof = ctx.profile_read()
print("points", n, "kept", kept, "ms/pass", ms, "GB/s(xyz only)", n * 12 / ms / 1e6)
for k, (t, c) in prof.items():
    print(f"  {k:24s} {1e3 * t / c:9.2f} us/launch  x{c}")
ctx.close()

"""Times pslam_bf_best2_dev on BASELINE config 5 (64k x 64k) with CUDA events; used for tuning runs
(PSLAM_BF_QPT=2|4 python tools/hamming_tune.py)."""
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from srrg2_proslam_b200 import capi, synth  # noqa: E402

n = 65536
dev = torch.device("cuda", 0)
ctx = capi.Context(device=0, max_images=2, max_rows=64, max_cols=128, max_features=256, max_raw_per_bin=1024)
q, t = synth.hamming_sets(n, n, seed=0)
dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
ob = torch.empty((3, n), dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for _ in range(3):
    ctx.bf_best2_dev(n, dq.data_ptr(), n, dt_.data_ptr(), ob[0].data_ptr(), ob[1].data_ptr(), ob[2].data_ptr())
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
reps = 20
for _ in range(reps):
    ctx.bf_best2_dev(n, dq.data_ptr(), n, dt_.data_ptr(), ob[0].data_ptr(), ob[1].data_ptr(), ob[2].data_ptr())
e1.record(stream)
ctx.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"PSLAM_BF_QPT={os.environ.get('PSLAM_BF_QPT', 'default')}: {ms:.3f} ms per sweep, {n * n / ms / 1e6:.1f} GPair/s")
ctx.close()

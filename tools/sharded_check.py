"""N-GPU check of the C-ABI sharded Hamming sweep (torchrun, one rank per GPU): every rank creates the NCCL communicator
through pslam_nccl_unique_id / pslam_nccl_comm_create (the id travels over torch.distributed's store), runs
pslam_bf_best2_sharded_dev on the SAME 64k x 64k sets (BASELINE config 5) and compares the gathered table with a
single-GPU sweep of its own; rank 0 prints one JSON line with the timing (CUDA events, max over ranks).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
"""
import json
import os
import pathlib
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from srrg2_proslam_b200 import capi, synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world, device_id=dev if world > 1 else None)
    ctx = capi.Context(device=local, max_images=2, max_rows=376, max_cols=1241, max_features=4096, max_raw_per_bin=8192)
    ids = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = ctx.nccl_comm_create(ids[0], rank, world)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    q, t = synth.hamming_sets(n, n, seed=0)
    dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    ref = torch.empty((3, n), dtype=torch.int32, device=dev)
    out = torch.full((3, n), -1, dtype=torch.int32, device=dev)
    ctx.bf_best2_dev(n, dq.data_ptr(), n, dt_.data_ptr(), ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr())
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    mode = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    if mode == "p2p":  # exchange fused into the merge kernel: peer tables mapped through CUDA IPC, no collective
        handles = [None] * world
        mine = ctx.p2p_table_export(n)
        if world > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles = [mine]
        ctx.p2p_table_import(rank, world, handles)
        run = lambda: ctx.bf_best2_sharded_p2p_dev(n, dq.data_ptr(), n, dt_.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                                   out[2].data_ptr())
    else:
        run = lambda: ctx.bf_best2_sharded_dev(comm, rank, world, n, dq.data_ptr(), n, dt_.data_ptr(), out[0].data_ptr(),
                                               out[1].data_ptr(), out[2].data_ptr())
    for _ in range(3):
        run()
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    ctx.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    ok = torch.tensor([int(torch.equal(out, ref))], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"metric": "hamming_best2_gpairs_per_s", "api": "pslam_bf_best2_sharded_p2p_dev (C ABI, merge kernel stores into every peer's table over NVLink, flags, no collective)"
                          if mode == "p2p" else "pslam_bf_best2_sharded_dev (C ABI, ncclAllGather on the context stream)",
                          "n_gpus": world, "n": n, "ms_per_sweep": float(ms.item()), "value": n * n / (float(ms.item()) * 1e-3) / 1e9,
                          "unit": "GPair/s", "parity_with_single_gpu_sweep": bool(ok.item())}))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if mode == "p2p":
        ctx.p2p_table_release()
    ctx.nccl_comm_destroy(comm)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Pinned host -> device copy rate of this box with ALL ranks copying at the same time (the bound of bench.py's end-to-end
number at N GPUs), for the host-buffer variants an application could choose:
  default        cudaHostAlloc(cudaHostAllocDefault)            (what torch's pin_memory gives)
  write_combined cudaHostAlloc(cudaHostAllocWriteCombined)      (no CPU cache snooping on the device's reads)
  two_streams    default pinned memory, each chunk split over two copy streams
Single GPU:  python tools/h2d_bandwidth.py
N GPUs:      python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_bandwidth.py
Rank 0 prints one JSON line: aggregate GB/s (bytes of all ranks / max time over ranks) and the per-rank rates."""
import ctypes
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rt = ctypes.CDLL("libcudart.so.12")
    chunk, n_chunks = 341 << 20, 6  # the pipeline's upload size (768 images of 1241 x 376)
    total = chunk * n_chunks
    dst = torch.empty(chunk, dtype=torch.uint8, device=dev)
    s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def host_alloc(flags):
        p = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(total), ctypes.c_uint(flags))
        assert rc == 0, rc
        ctypes.memset(p, 1, total)  # touch every page
        return p

    def memcpy_async(dst_ptr, src_ptr, nbytes, stream):
        rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), ctypes.c_size_t(nbytes), ctypes.c_int(1),
                                ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0, rc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(host, two):
        best = None
        for rep in range(4):  # first repetition = warm-up
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(s0)
            if two:
                s1.wait_event(e0)
            for i in range(n_chunks):
                src = host.value + i * chunk
                if two:
                    memcpy_async(dst.data_ptr(), src, chunk // 2, s0)
                    memcpy_async(dst.data_ptr() + chunk // 2, src + chunk // 2, chunk - chunk // 2, s1)
                else:
                    memcpy_async(dst.data_ptr(), src, chunk, s0)
            if two:
                e2.record(s1)
                s0.wait_event(e2)
            e1.record(s0)
            s0.synchronize()
            mine = e0.elapsed_time(e1)
            t = torch.tensor([mine], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0 and (best is None or float(t.item()) < best[0]):
                best = (float(t.item()), mine)
        return world * total / (best[0] * 1e-3) / 1e9, total / (best[1] * 1e-3) / 1e9

    out = {"n_gpus": world, "chunk_mib": chunk >> 20, "chunks": n_chunks}
    for name, flags, two in (("default", 0, False), ("write_combined", 4, False), ("two_streams", 0, True)):
        host = host_alloc(flags)
        agg, mine = run(host, two)
        rates = [None] * world
        if world > 1:
            dist.all_gather_object(rates, (rank, round(mine, 1), torch.cuda.get_device_properties(local).pci_bus_id))
        else:
            rates = [(0, round(mine, 1), torch.cuda.get_device_properties(local).pci_bus_id)]
        out[name] = {"aggregate_gbs": agg, "per_rank_gbs": rates}
        rt.cudaFreeHost(host)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

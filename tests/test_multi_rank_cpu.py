"""World-size-2 (gloo, CPU) tests of the N > 1 host logic: frame sharding with a host-side gather and
query-row sharding of the Hamming sweep with one all-gather.  The per-rank compute is done by the CPU oracle here
(no GPU in this container); on the GPU box bench.py runs the same sharding with the CUDA path and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from srrg2_proslam_b200 import sharding, synth


def test_ranges_partition_exactly():
    for n in (0, 1, 7, 8, 10000, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            fr = [sharding.frame_range(n, r, world) for r in range(world)]
            assert fr[0][0] == 0 and fr[-1][1] == n and all(a[1] == b[0] for a, b in zip(fr, fr[1:]))
            assert max(e - b for b, e in fr) - min(e - b for b, e in fr) <= 1
            qr = [sharding.query_rows(n, r, world) for r in range(world)]
            assert qr[0][0] == 0 and qr[-1][1] == n and all(a[1] == b[0] for a, b in zip(qr, qr[1:]))
            assert all(b % 256 == 0 for b, e in qr if e > b)


def test_c_abi_shard_ranges_equal_the_python_ones():
    """pslam_shard_rows / pslam_shard_frames (pure functions of the C ABI, include/pslam_cuda.h) give the same ranges as
    the torch.distributed harness"""
    from srrg2_proslam_b200 import capi
    for n in (0, 1, 255, 256, 257, 700, 10000, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                b, e, per = capi.shard_rows(n, r, world)
                assert (b, e) == sharding.query_rows(n, r, world) and per % 256 == 0 and per * world >= n
                assert capi.shard_frames(n, r, world) == sharding.frame_range(n, r, world)
    with pytest.raises(capi.PslamError):
        capi.shard_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nq, nt, n_pairs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O.build()
        # ---- Hamming sweep: query rows sharded, train set replicated, one all-gather --------------------------
        q, t = synth.hamming_sets(nq, nt, seed=3)
        b, e = sharding.query_rows(nq, rank, world)
        if e > b:
            best, second, idx = O.bf_best2(q[b:e].view(np.uint8).reshape(-1, 32), t.view(np.uint8).reshape(-1, 32))
            local = torch.from_numpy(np.stack([best, second, idx]).astype(np.int32))
        else:
            local = torch.zeros((3, 0), dtype=torch.int32)
        full = sharding.allgather_best2(local, nq)
        # ---- frame sharding: contiguous ranges, host-side gather of variable-length results -------------------
        imgs = synth.stereo_pairs(n_pairs, 120, 320, seed=11, device="cpu").numpy()
        fb, fe = sharding.frame_range(n_pairs, rank, world)
        cfg = O.extract_cfg(15, 1, 400)
        counts, pts = [], []
        for p in range(fb, fe):
            r = O.stereo_adaptor(imgs[p, 0], imgs[p, 1], cfg, "epipolar", 100, 0.5, 100, 0)
            counts.append(len(r["uvuv"]))
            pts.append(r["uvuv"])
        lc = torch.tensor(counts, dtype=torch.int64)
        lp = torch.from_numpy(np.concatenate(pts) if pts else np.zeros((0, 4), np.float32))
        gc, gp = sharding.gather_frame_results(lc, lp)
        if rank == 0:
            out["best2"] = full.numpy()
            out["counts"] = gc.numpy()
            out["points"] = gp.numpy()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world_size_2_gloo(oracle):
    nq, nt, n_pairs = 700, 900, 5  # ragged on purpose: shards of 512 + 188 rows, 3 + 2 frames
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), nq, nt, n_pairs, out), nprocs=2, join=True)
    q, t = synth.hamming_sets(nq, nt, seed=3)
    best, second, idx = O.bf_best2(q.view(np.uint8).reshape(-1, 32), t.view(np.uint8).reshape(-1, 32))
    assert np.array_equal(out["best2"], np.stack([best, second, idx]).astype(np.int32))
    imgs = synth.stereo_pairs(n_pairs, 120, 320, seed=11, device="cpu").numpy()
    cfg = O.extract_cfg(15, 1, 400)
    ref = [O.stereo_adaptor(imgs[p, 0], imgs[p, 1], cfg, "epipolar", 100, 0.5, 100, 0)["uvuv"] for p in range(n_pairs)]
    assert np.array_equal(out["counts"], [len(r) for r in ref]) and out["counts"].sum() > 0
    assert np.array_equal(out["points"], np.concatenate(ref))

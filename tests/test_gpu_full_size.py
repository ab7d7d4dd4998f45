"""BASELINE.json's full-size configurations checked through size-independent properties (the oracle is too slow to run
them whole): config 5 -- 64k x 64k 256-bit Hamming sweep -- and a config-4-shaped batch (KITTI-shaped synthetic pairs,
4k target features) whose result must not depend on how the batch is chunked or sharded."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
POPC = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_rows(q_rows, t):
    return POPC[q_rows[:, None, :] ^ t[None, :, :]].sum(axis=2)


def test_hamming_64k_x_64k_properties(oracle):
    """config 5: planted near-duplicates are found with their exact distance, a random sample of query rows agrees with
    a numpy popcount sweep over ALL 65536 train rows (best, second best, first argmin), and the row-sharded sweep
    (BASELINE: query rows sharded at 1/2/4/8 GPUs) returns the same table as the unsharded one"""
    from srrg2_proslam_b200 import capi, sharding, synth
    nq = nt = 65536
    q, t = synth.hamming_sets(nq, nt, seed=0)
    ctx = capi.Context(device=0, max_images=2, max_rows=64, max_cols=64, max_features=256, max_raw_per_bin=1024)
    try:
        best, second, idx = ctx.bf_best2(q, t)
        assert best.shape == (nq,) and np.all(best >= 0) and np.all(best <= second) and np.all((idx >= 0) & (idx < nt))
        # the reported distance is the distance to the reported train row
        sample = np.random.default_rng(1).choice(nq, 4096, replace=False)
        assert np.array_equal(POPC[q[sample] ^ t[idx[sample]]].sum(axis=1), best[sample])
        # exhaustive check of 48 rows (half of them planted: best <= 20)
        planted = np.flatnonzero(best <= 20)
        assert 0.09 * nq < len(planted) < 0.11 * nq
        rows = np.concatenate([planted[:24], np.flatnonzero(best > 20)[:24]])
        d = hamming_rows(q[rows], t)
        srt = np.sort(d, axis=1)
        assert np.array_equal(srt[:, 0], best[rows]) and np.array_equal(srt[:, 1], second[rows])
        assert np.array_equal(np.argmin(d, axis=1), idx[rows])  # strict <: the first minimum wins
        # row sharding (world sizes 2, 4, 8): concatenating the shards' tables reproduces the table
        for world in (2, 4, 8):
            parts = []
            for rank in range(world):
                b, e = sharding.query_rows(nq, rank, world)
                parts.append(ctx.bf_best2(q[b:e], t) if rank in (0, world - 1) else (best[b:e], second[b:e], idx[b:e]))
            for k, full in enumerate((best, second, idx)):
                assert np.array_equal(np.concatenate([p[k] for p in parts]), full)
    finally:
        ctx.close()


def test_frame_batch_is_chunking_and_sharding_invariant(oracle):
    """config 4 shape (1241 x 376, 4k target features, kitti.conf matcher): 96 synthetic pairs processed (a) in one batch
    with 512-image chunks, (b) with 14-image chunks, (c) as two rank shards -- identical packed clouds; a 16-pair sample
    is checked against the oracle bit for bit"""
    import torch
    from srrg2_proslam_b200 import capi, sharding, synth
    P, rows, cols = 96, 376, 1241
    imgs = synth.stereo_pairs(P, rows, cols, seed=5, device="cuda")
    e, m = capi.extract_cfg(15, 1, 4000), capi.match_cfg(100.0, 0.5, 100, 0)
    want = ("uvuv", "intensity", "desc")

    def run(images, n, work_images):
        ctx = capi.Context(device=0, max_images=2 * n, max_rows=rows, max_cols=cols, max_features=4096, max_raw_per_bin=8192,
                           max_bins=9, work_images=work_images)
        try:
            ctx.stereo_frontend_batch_dev(images.data_ptr(), n, rows, cols, cols, rows * cols, e, m)
            r = ctx.download_stereo_batch(n, 4096 * n, want=want)
            return {k: np.array(r[k][:r["n"]] if k != "offsets" else r[k]) for k in want + ("offsets",)}
        finally:
            ctx.close()

    a = run(imgs, P, 512)
    b = run(imgs, P, 14)
    for k in want + ("offsets",):
        assert np.array_equal(a[k], b[k]), k
    assert a["offsets"][-1] > 200 * P  # hundreds of accepted stereo points per frame (ratio test 0.5) out of ~3.2 k features
    parts = []
    for rank in range(2):
        lo, hi = sharding.frame_range(P, rank, 2)
        parts.append(run(imgs[lo:hi].contiguous(), hi - lo, 64))
    for k in want:
        assert np.array_equal(np.concatenate([p[k] for p in parts]), a[k]), k
    assert np.array_equal(np.concatenate([parts[0]["offsets"], parts[1]["offsets"][1:] + parts[0]["offsets"][-1]]), a["offsets"])
    # oracle sample
    sample = imgs[:16].cpu().numpy()
    counts, chk = O.stereo_frontend_batch(sample, O.extract_cfg(15, 1, 4000), threads=8, max_dist=100.0, ratio=0.5, max_disp=100, thickness=0)
    off = a["offsets"]
    assert np.array_equal(np.diff(off)[:16], counts)
    for p in range(16):
        s, t_ = off[p], off[p + 1]
        assert O.fnv1a_points(a["uvuv"][s:t_], a["intensity"][s:t_], a["desc"][s:t_]) == chk[p], p

"""GPU parity of the batched (chunked, double-buffered) stereo frontend and its packed download against the
CPU oracle, per pair, bit-exact; device-resident and host-pointer entry points; ragged chunk sizes."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

MATCH = dict(max_dist=100.0, ratio=0.5, max_disp=100, thickness=0)


def golden_pairs():
    names = [(f"kitti_city_image_left_{i}.png", f"kitti_city_image_right_{i}.png") for i in range(5)]
    return np.stack([np.stack([O.load_gray(l), O.load_gray(r)]) for l, r in names])


def check_batch(ctx, imgs, res, cfg):
    n = len(imgs)
    counts, chk = O.stereo_frontend_batch(imgs, cfg, threads=4, **MATCH)
    off = res["offsets"]
    assert off[0] == 0 and off[n] == res["n"]
    assert np.array_equal(np.diff(off), counts)
    assert np.array_equal(ctx.stereo_counts(n), counts)
    for p in range(n):
        a, b = off[p], off[p + 1]
        assert O.fnv1a_points(res["uvuv"][a:b], res["intensity"][a:b], res["desc"][a:b]) == chk[p], p
    # per-pair accessors agree with the packed result
    for p in (0, n - 1):
        one = ctx.download_stereo_points(p)
        a, b = off[p], off[p + 1]
        assert np.array_equal(one["uvuv"], res["uvuv"][a:b])
        assert np.array_equal(one["left_idx"], res["left_idx"][a:b])
        assert np.array_equal(one["right_idx"], res["right_idx"][a:b])
        assert np.array_equal(one["distance"], res["distance"][a:b])
        f = ctx.download_features(2 * p)
        assert np.array_equal(f["desc"][one["left_idx"]], res["desc"][a:b])


@pytest.mark.parametrize("work_images", [2, 4, 6, 64])
def test_batch_golden_kitti(oracle, work_images):
    import torch
    from srrg2_proslam_b200 import capi
    imgs = golden_pairs()
    n, _, rows, cols = imgs.shape
    ctx = capi.Context(max_images=2 * n, max_rows=rows, max_cols=cols, max_features=2048, max_raw_per_bin=8192,
                       work_images=work_images)
    try:
        e, m = capi.extract_cfg(15, 1, 1000), capi.match_cfg(**MATCH)
        ocfg = O.extract_cfg(15, 1, 1000)
        # host-pointer entry point (pageable memory)
        ctx.stereo_frontend_batch(imgs, n, rows, cols, cols, rows * cols, e, m)
        check_batch(ctx, imgs, ctx.download_stereo_batch(n, 2048 * n), ocfg)
        # device-resident entry point
        d = torch.from_numpy(imgs).cuda()
        ctx.stereo_frontend_batch_dev(d.data_ptr(), n, rows, cols, cols, rows * cols, e, m)
        check_batch(ctx, imgs, ctx.download_stereo_batch(n, 2048 * n), ocfg)
        # a sub-batch re-uses the same context: only the first 3 pairs
        ctx.stereo_frontend_batch(imgs, 3, rows, cols, cols, rows * cols, e, m)
        check_batch(ctx, imgs[:3], ctx.download_stereo_batch(3, 2048 * 3), ocfg)
    finally:
        ctx.close()


def test_batch_synthetic_4k(oracle):
    """BASELINE config 4 shape: KITTI-sized synthetic pairs, 4k target features / frame, pinned host input"""
    import torch
    from srrg2_proslam_b200 import capi, synth
    n = 7
    t = synth.stereo_pairs(n, seed=123, device="cuda")
    imgs = t.cpu().numpy()
    rows, cols = imgs.shape[2:]
    ctx = capi.Context(max_images=2 * n, max_rows=rows, max_cols=cols, max_features=4096, max_raw_per_bin=8192,
                       work_images=6)
    try:
        e, m = capi.extract_cfg(15, 1, 4000), capi.match_cfg(**MATCH)
        ocfg = O.extract_cfg(15, 1, 4000)
        ctx.stereo_frontend_batch_dev(t.data_ptr(), n, rows, cols, cols, rows * cols, e, m)
        res = ctx.download_stereo_batch(n, 4096 * n)
        check_batch(ctx, imgs, res, ocfg)
        assert res["n"] > 200 * n
        pinned = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(t)
        torch.cuda.synchronize()
        for _ in range(2):  # twice: staging-buffer reuse across calls
            ctx.stereo_frontend_batch(pinned.data_ptr(), n, rows, cols, cols, rows * cols, e, m)
            check_batch(ctx, imgs, ctx.download_stereo_batch(n, 4096 * n), ocfg)
    finally:
        ctx.close()


def test_batch_strided_host_images(oracle):
    """host images with row padding (stride > cols) and a gap between images: the 2-D copy path"""
    from srrg2_proslam_b200 import capi
    imgs = golden_pairs()[:2]
    n, _, rows, cols = imgs.shape
    stride, pitch = cols + 39, (rows + 3) * (cols + 39) + 17
    buf = np.zeros(2 * n * pitch, np.uint8)
    for i in range(2 * n):
        v = np.lib.stride_tricks.as_strided(buf[i * pitch:], (rows, cols), (stride, 1))
        v[:] = imgs.reshape(2 * n, rows, cols)[i]
    ctx = capi.Context(max_images=2 * n, max_rows=rows + 8, max_cols=cols + 64, max_features=2048, max_raw_per_bin=8192,
                       work_images=2)
    try:
        e, m = capi.extract_cfg(15, 1, 1000), capi.match_cfg(**MATCH)
        ctx.stereo_frontend_batch(buf, n, rows, cols, stride, pitch, e, m)
        check_batch(ctx, imgs, ctx.download_stereo_batch(n, 2048 * n), O.extract_cfg(15, 1, 1000))
    finally:
        ctx.close()


def test_batch_empty_and_limits(oracle):
    from srrg2_proslam_b200 import capi
    imgs = golden_pairs()[:1]
    _, _, rows, cols = imgs.shape
    ctx = capi.Context(max_images=2, max_rows=rows, max_cols=cols, max_features=2048, max_raw_per_bin=8192)
    try:
        e, m = capi.extract_cfg(15, 1, 1000), capi.match_cfg(**MATCH)
        ctx.stereo_frontend_batch(imgs, 0, rows, cols, cols, rows * cols, e, m)
        assert ctx.download_stereo_batch(0, 16)["n"] == 0
        with pytest.raises(capi.PslamError):
            ctx.stereo_frontend_batch(np.zeros((2, 2, rows, cols), np.uint8), 2, rows, cols, cols, rows * cols, e, m)
        flat = np.full((1, 2, rows, cols), 90, np.uint8)   # no corners at all
        ctx.stereo_frontend_batch(flat, 1, rows, cols, cols, rows * cols, e, m)
        assert ctx.download_stereo_batch(1, 16)["n"] == 0
    finally:
        ctx.close()


@pytest.mark.parametrize("max_rows", [376, 510, 511, 600, 4200])
def test_batch_matcher_shared_memory_layouts(oracle, max_rows):
    """the pipeline matcher picks its shared-memory layout from the context's row limit: the lean one-pass layout up to
    510 rows, the full counting-sort layout up to 4096, the bitonic one above -- all three give the oracle's result"""
    from srrg2_proslam_b200 import capi, synth
    n = 3
    imgs = synth.stereo_pairs(n, seed=77, device="cuda").cpu().numpy()
    rows, cols = imgs.shape[2:]
    ctx = capi.Context(max_images=2 * n, max_rows=max(rows, max_rows), max_cols=cols, max_features=4096,
                       max_raw_per_bin=8192, work_images=4)
    try:
        e, m = capi.extract_cfg(15, 1, 4000), capi.match_cfg(**MATCH)
        ctx.stereo_frontend_batch(imgs, n, rows, cols, cols, rows * cols, e, m)
        check_batch(ctx, imgs, ctx.download_stereo_batch(n, 4096 * n), O.extract_cfg(15, 1, 4000))
    finally:
        ctx.close()


def test_batch_euroc_shaped_synthetic(oracle):
    """EuRoC-shaped synthetic (752x480 pairs), euroc.conf parameters: FAST threshold 10, target
    1000 per image, epipolar matcher max distance 75, ratio 0.5, disparities up to 200 px"""
    from srrg2_proslam_b200 import capi, synth
    n = 5
    imgs = synth.stereo_pairs(n, rows=480, cols=752, seed=0, device="cuda", max_disp=200).cpu().numpy()
    match = dict(max_dist=75.0, ratio=0.5, max_disp=200, thickness=0)
    ctx = capi.Context(max_images=2 * n, max_rows=480, max_cols=752, max_features=2048, max_raw_per_bin=8192, work_images=4)
    try:
        ctx.stereo_frontend_batch(imgs, n, 480, 752, 752, 480 * 752, capi.extract_cfg(10, 1, 1000), capi.match_cfg(**match))
        res = ctx.download_stereo_batch(n, 2048 * n)
        counts, chk = O.stereo_frontend_batch(imgs, O.extract_cfg(10, 1, 1000), threads=4, **match)
        off = res["offsets"]
        assert np.array_equal(np.diff(off), counts) and res["n"] > 20 * n
        for p in range(n):
            a, b = off[p], off[p + 1]
            assert O.fnv1a_points(res["uvuv"][a:b], res["intensity"][a:b], res["desc"][a:b]) == chk[p], p
    finally:
        ctx.close()


def test_chunk_lanes_give_identical_results(oracle):
    """pslam_set_lanes: a batch that spans several chunks produces the same stereo clouds whether its chunks alternate over
    two streams with their own intermediates (default) or run one after the other on one stream"""
    import torch
    from srrg2_proslam_b200 import capi, synth
    n_pairs = 9  # 18 images, chunks of 4: five chunks, the last one short
    imgs = synth.stereo_pairs(n_pairs, 120, 320, seed=21, device="cuda")
    ctx = capi.Context(max_images=2 * n_pairs, max_rows=120, max_cols=320, max_features=1024, max_raw_per_bin=4096, work_images=4)
    try:
        ecfg, mcfg = capi.extract_cfg(15, 1, 400), capi.match_cfg(100, 0.5, 100, 0)
        outs = []
        for lanes in (2, 1, 2):
            ctx.set_lanes(lanes)
            ctx.stereo_frontend_batch_dev(imgs.data_ptr(), n_pairs, 120, 320, 320, 120 * 320, ecfg, mcfg)
            ctx.synchronize()
            counts = ctx.stereo_counts(n_pairs)
            outs.append((counts.copy(), [ctx.download_stereo_points(p)["uvuv"] for p in range(n_pairs)]))
        assert outs[0][0].sum() > 50
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0])
            assert all(np.array_equal(a, b) for a, b in zip(o[1], outs[0][1]))
        h = imgs.cpu().numpy()
        ref = O.stereo_adaptor(h[3, 0], h[3, 1], O.extract_cfg(15, 1, 400), "epipolar", 100, 0.5, 100, 0)
        assert np.array_equal(outs[0][1][3], ref["uvuv"])
    finally:
        ctx.close()


def test_native_c_caller(tmp_path):
    """tests/native/abi_smoke.c: a C99 program against include/pslam_cuda.h + libpslam_cuda.so -- extraction and the stereo
    adaptor on a synthetic pair (right = left shifted by 12 px: every stereo point has disparity 12), deterministic"""
    from test_cpu_abi import _build_native_caller
    import subprocess
    r = subprocess.run([str(_build_native_caller(tmp_path))], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr

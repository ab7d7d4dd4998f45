"""GPU parity, stage 1 (detect, select, describe) through the C ABI vs the CPU oracle. Bit-exact."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

IMAGES = ["kitti_city_image_left_0.png", "kitti_city_image_right_1.png", "icl_image_rgb_0.png",
          "icl_image_rgb_50.png", "scene_flow_image_left.png", "kitti_highway_image_left_274.png"]


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(max_images=2, max_rows=600, max_cols=1300, max_features=8192, max_raw_per_bin=60000)
    yield c
    c.close()


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("thr,nms", [(15, 1), (5, 1), (40, 1), (10, 0)])
def test_fast_nms(ctx, name, thr, nms):
    img = O.load_gray(name)
    xy, r = ctx.fast_detect(img, thr, nms)
    oxy, orr = O.fast_detect(img, thr, nms)
    assert len(xy) == len(oxy) and len(xy) > 20
    assert np.array_equal(xy, oxy) and np.array_equal(r, orr)


@pytest.mark.parametrize("name", IMAGES[:4])
def test_blur(ctx, name):
    img = O.load_gray(name)
    assert np.array_equal(ctx.blur7(img), O.blur7(img))


CFGS = [dict(threshold=15, target=1000), dict(threshold=5, target=500), dict(threshold=5, target=300),
        dict(threshold=5, target=1000, nh=1, nv=1), dict(threshold=10, target=1000, nh=4, nv=2),
        dict(threshold=5, target=300, nh=1, nv=1), dict(threshold=20, target=5, nh=3, nv=3),
        dict(threshold=15, target=4000)]


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("cfg", CFGS)
def test_extract_binned(ctx, name, cfg):
    from srrg2_proslam_b200 import capi
    img = O.load_gray(name)
    g = ctx.extract_binned(img, capi.extract_cfg(**cfg))
    o = O.extract_binned(img, O.extract_cfg(**cfg))
    assert len(g["xy"]) == len(o["xy"])
    for k in ("xy", "response", "intensity", "desc"):
        assert np.array_equal(g[k], o[k]), k


def test_known_answer_counts(ctx):
    """the reference's own constants, straight through the CUDA path (tests/test_feature_extractors.cpp)"""
    from srrg2_proslam_b200 import capi
    L0 = O.load_gray("kitti_city_image_left_0.png")
    assert len(ctx.extract_binned(L0, capi.extract_cfg(5, 1, 1000, 1, 1))["xy"]) == 887
    got = [len(ctx.extract_binned(O.load_gray(n), capi.extract_cfg(5, 1, 300, 3, 3))["xy"])
           for n in ("kitti_city_image_left_0.png", "kitti_city_image_left_1.png",
                     "kitti_city_image_right_0.png", "kitti_city_image_right_1.png")]
    assert got == [272, 280, 270, 271]
    assert [len(ctx.extract_binned(O.load_gray(f"icl_image_rgb_{i}.png"), capi.extract_cfg(5, 1, 500))["xy"])
            for i in (0, 1, 50)] == [321, 338, 261]


def test_mask(ctx):
    from srrg2_proslam_b200 import capi
    img = O.load_gray("kitti_city_image_left_0.png")
    mask = np.zeros_like(img)
    mask[50:300, 200:700] = 255
    mask[100:120, 300:340] = 0
    g = ctx.extract_binned(img, capi.extract_cfg(10, 1, 500), mask=mask)
    o = O.extract_binned(img, O.extract_cfg(10, 1, 500), mask=mask)
    assert len(g["xy"]) == len(o["xy"]) > 100
    for k in ("xy", "response", "intensity", "desc"):
        assert np.array_equal(g[k], o[k]), k


def test_edge_images(ctx):
    """flat image (no corners), tiny image, synthetic ties (few distinct responses)"""
    from srrg2_proslam_b200 import capi
    flat = np.full((100, 160), 77, np.uint8)
    assert len(ctx.extract_binned(flat, capi.extract_cfg(5, 1, 100))["xy"]) == 0
    rng = np.random.default_rng(3)
    # coarsely quantised real image: thousands of corners share 5 distinct responses -> exercises the
    # unstable-sort ties of the per-region selection (binned.cpp:188-192)
    img = (O.load_gray("kitti_city_image_left_0.png") // 48 * 48).astype(np.uint8)
    for cfg in (dict(threshold=20, target=200), dict(threshold=20, target=90, nh=2, nv=2),
                dict(threshold=20, target=1000, nh=1, nv=1)):
        g = ctx.extract_binned(img, capi.extract_cfg(**cfg))
        o = O.extract_binned(img, O.extract_cfg(**cfg))
        assert len(g["xy"]) == len(o["xy"]) > 0
        for k in ("xy", "response", "intensity", "desc"):
            assert np.array_equal(g[k], o[k]), (cfg, k)
    small = rng.integers(0, 255, (70, 90)).astype(np.uint8)
    g = ctx.extract_binned(small, capi.extract_cfg(5, 1, 100))
    o = O.extract_binned(small, O.extract_cfg(5, 1, 100))
    assert np.array_equal(g["xy"], o["xy"]) and np.array_equal(g["desc"], o["desc"])


def test_errors(ctx):
    from srrg2_proslam_b200 import capi
    img = O.load_gray("icl_image_rgb_0.png")
    with pytest.raises(capi.PslamError) as e:
        ctx.extract_binned(img, capi.extract_cfg(5, 1, 100, 0, 3))
    assert "invalid number of horizontal detectors" in str(e.value)
    big = np.zeros((700, 100), np.uint8)
    with pytest.raises(capi.PslamError):
        ctx.extract_binned(big, capi.extract_cfg(5, 1, 100))


def test_strided_single_image(ctx):
    """per-frame entry point with a row stride larger than the width: a view that still fits the staging slot travels as one
    linear copy and is processed with its host stride, a wider one takes the 2-D re-pitching copy; same features either way"""
    import ctypes as C
    from srrg2_proslam_b200 import capi
    img = O.load_gray("kitti_city_image_left_0.png")
    rows, cols = img.shape
    cfg, ocfg = capi.extract_cfg(15, 1, 1000), O.extract_cfg(15, 1, 1000)
    want = O.extract_binned(img, ocfg)
    for stride in (cols + 7, cols + 39, 2 * cols):  # 1248, 1280 (= the slot pitch), 2482 (does not fit the slot: 2-D copy)
        buf = np.full((rows, stride), 255, np.uint8)
        buf[:, :cols] = img
        cap = ctx.limits.max_features
        xy, resp = np.zeros((cap, 2), np.float32), np.zeros(cap, np.float32)
        inten, desc = np.zeros(cap, np.float32), np.zeros((cap, 32), np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n = capi.lib().pslam_extract_binned(ctx._h, p(buf), rows, cols, stride, C.byref(cfg), None, cap, p(xy), p(resp), p(inten), p(desc))
        assert n == len(want["xy"]), (stride, n)
        assert np.array_equal(xy[:n], want["xy"]) and np.array_equal(desc[:n], want["desc"]) and np.array_equal(inten[:n], want["intensity"])

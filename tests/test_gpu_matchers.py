"""GPU parity, stage 2 (epipolar / brute-force / projective matching, adaptors) vs the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

K_KITTI = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(max_images=2, max_rows=600, max_cols=1300, max_features=4096, max_raw_per_bin=40000)
    yield c
    c.close()


@pytest.fixture(scope="module")
def feats(oracle):
    c = O.extract_cfg(threshold=5, target=500)
    names = dict(L0="kitti_city_image_left_0.png", R0="kitti_city_image_right_0.png",
                 L1="kitti_city_image_left_1.png", I0="icl_image_rgb_0.png", I1="icl_image_rgb_1.png",
                 I50="icl_image_rgb_50.png", SL="scene_flow_image_left.png", SR="scene_flow_image_right.png")
    return {k: O.extract_binned(O.load_gray(v), c) for k, v in names.items()}


def same_corr(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("pair", [("L0", "R0"), ("L0", "L0"), ("SL", "SR"), ("I0", "I1")])
@pytest.mark.parametrize("params", [dict(max_dist=50, ratio=0.9, max_disp=100, thickness=0),
                                    dict(max_dist=50, ratio=0.9, max_disp=100, thickness=1),
                                    dict(max_dist=100, ratio=0.8, max_disp=100, thickness=0),
                                    dict(max_dist=75, ratio=0.5, max_disp=200, thickness=3),
                                    dict(max_dist=30, ratio=0.99, max_disp=20, thickness=2)])
def test_epipolar(ctx, feats, pair, params):
    from srrg2_proslam_b200 import capi
    f, m = feats[pair[0]], feats[pair[1]]
    g = ctx.match_epipolar(f["xy"], f["desc"], m["xy"], m["desc"], capi.match_cfg(**params))
    o = O.match_epipolar(f["xy"], f["desc"], m["xy"], m["desc"], **params)
    assert len(g[0]) == len(o[0])
    assert same_corr(g, o)


def test_epipolar_known_answers(ctx, feats):  # tests/test_correspondence_finders.cpp:176-180,274,290
    from srrg2_proslam_b200 import capi
    f, m = feats["L0"], feats["R0"]
    fi, mi, d = ctx.match_epipolar(f["xy"], f["desc"], f["xy"], f["desc"], capi.match_cfg(50, 0.9, 100, 0))
    assert len(fi) == 446 and np.array_equal(fi, mi)
    assert len(ctx.match_epipolar(f["xy"], f["desc"], m["xy"], m["desc"], capi.match_cfg(50, 0.9, 100, 0))[0]) == 150
    assert len(ctx.match_epipolar(f["xy"], f["desc"], m["xy"], m["desc"], capi.match_cfg(50, 0.9, 100, 1))[0]) == 241


def test_epipolar_empty_and_ragged(ctx, feats):
    from srrg2_proslam_b200 import capi
    f = feats["L0"]
    e_xy, e_d = np.zeros((0, 2), np.float32), np.zeros((0, 32), np.uint8)
    assert len(ctx.match_epipolar(e_xy, e_d, f["xy"], f["desc"], capi.match_cfg())[0]) == 0
    assert len(ctx.match_epipolar(f["xy"], f["desc"], e_xy, e_d, capi.match_cfg())[0]) == 0
    g = ctx.match_epipolar(f["xy"][:7], f["desc"][:7], f["xy"], f["desc"], capi.match_cfg(50, 0.9, 100, 1))
    o = O.match_epipolar(f["xy"][:7], f["desc"][:7], f["xy"], f["desc"], 50, 0.9, 100, 1)
    assert same_corr(g, o)


def test_epipolar_synthetic_rows(ctx):
    """dense rows with duplicate descriptors: ordering constraint, ties, ratio failures"""
    from srrg2_proslam_b200 import capi
    rng = np.random.default_rng(11)
    n = 1500
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)

    def cloud():
        xy = np.stack([rng.integers(0, 300, n), rng.integers(0, 12, n)], 1).astype(np.float32)
        _, first = np.unique(xy, axis=0, return_index=True)
        xy = xy[np.sort(first)]
        d = base[rng.integers(0, 40, len(xy))].copy()
        flip = rng.random(d.shape) < 0.02
        d ^= (flip * (1 << rng.integers(0, 8, d.shape))).astype(np.uint8)
        return xy, d

    fx, fd = cloud()
    mx, md = cloud()
    for params in (dict(max_dist=50, ratio=0.9, max_disp=100, thickness=0),
                   dict(max_dist=80, ratio=0.95, max_disp=30, thickness=2)):
        g = ctx.match_epipolar(fx, fd, mx, md, capi.match_cfg(**params))
        o = O.match_epipolar(fx, fd, mx, md, **params)
        assert len(o[0]) > 20 and same_corr(g, o)


@pytest.mark.parametrize("pair,n", [(("I0", "I0"), 319), (("I0", "I1"), 226), (("I0", "I50"), 117),
                                    (("L0", "R0"), 237)])
def test_bruteforce_known_answers(ctx, feats, pair, n):  # tests/test_correspondence_finders.cpp:37-41,72,126,214
    from srrg2_proslam_b200 import capi
    f, m = feats[pair[0]], feats[pair[1]]
    g = ctx.match_bruteforce(f["desc"], m["desc"], capi.match_cfg(50, 0.9))
    o = O.match_bruteforce(f["desc"], m["desc"], 50, 0.9)
    assert len(g[0]) == n == len(o[0])
    assert same_corr(g, o)


@pytest.mark.parametrize("max_dist,ratio", [(100, 0.8), (30, 0.99), (257, 0.5), (1, 0.9)])
def test_bruteforce_params(ctx, feats, max_dist, ratio):
    from srrg2_proslam_b200 import capi
    f, m = feats["SL"], feats["SR"]
    g = ctx.match_bruteforce(f["desc"], m["desc"], capi.match_cfg(max_dist, ratio))
    o = O.match_bruteforce(f["desc"], m["desc"], max_dist, ratio)
    assert same_corr(g, o)


def test_bruteforce_collisions(ctx):
    """duplicated descriptors on both sides: pools with crossed fixed/moving, Lowe list edge cases"""
    from srrg2_proslam_b200 import capi
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (60, 32), dtype=np.uint8)

    def cloud(n):
        d = base[rng.integers(0, 60, n)].copy()
        flip = rng.random(d.shape) < 0.01
        d ^= (flip * (1 << rng.integers(0, 8, d.shape))).astype(np.uint8)
        return d

    for nf, nm in ((300, 280), (1, 50), (50, 1), (1, 1), (513, 1025), (1500, 1700)):  # last: > 20 k candidates, sorted in global memory
        f, m = cloud(nf), cloud(nm)
        g = ctx.match_bruteforce(f, m, capi.match_cfg(50, 0.9))
        o = O.match_bruteforce(f, m, 50, 0.9)
        assert same_corr(g, o), (nf, nm)
    e = np.zeros((0, 32), np.uint8)
    assert len(ctx.match_bruteforce(e, cloud(5), capi.match_cfg())[0]) == 0


@pytest.mark.parametrize("nf,nm", [(1000, 3000), (257, 255), (2048, 2048), (5, 70000)])
def test_bf_best2(ctx, nf, nm):
    rng = np.random.default_rng(nf + nm)
    f = rng.integers(0, 256, (nf, 32), dtype=np.uint8)
    m = rng.integers(0, 256, (nm, 32), dtype=np.uint8)
    m[rng.integers(0, nm, nf // 4)] = f[rng.integers(0, nf, nf // 4)]  # exact duplicates -> ties
    gb, gs, gi = ctx.bf_best2(f, m)
    ob, os_, oi = O.bf_best2(f, m)
    assert np.array_equal(gb, ob) and np.array_equal(gs, os_) and np.array_equal(gi, oi)


def test_bf_best2_sharded_single_rank_through_nccl(ctx):
    """pslam_bf_best2_sharded_dev with a one-rank NCCL communicator created through the C ABI: shard -> sweep ->
    ncclAllGather on the context's stream -> unshard == the plain sweep (the 2- and 8-GPU runs: tools/sharded_check.py)"""
    import torch
    rng = np.random.default_rng(9)
    nq, nt = 1500, 2100
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    comm = ctx.nccl_comm_create(ctx.nccl_unique_id(), 0, 1)
    try:
        dq, dt_ = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
        out = torch.full((3, nq), -7, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.bf_best2_sharded_dev(comm, 0, 1, nq, dq.data_ptr(), nt, dt_.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                 out[2].data_ptr())
        ctx.synchronize()
    finally:
        ctx.nccl_comm_destroy(comm)
    ob, os_, oi = O.bf_best2(q, t)
    g = out.cpu().numpy()
    assert np.array_equal(g[0], ob) and np.array_equal(g[1], os_) and np.array_equal(g[2], oi)


def test_bf_best2_sharded_p2p_single_rank(ctx):
    """pslam_bf_best2_sharded_p2p_dev with a one-rank world: export / import of the result table, merge kernel storing into
    it, flag + wait + hand-out, two consecutive epochs (both table parities) == the plain sweep (N ranks: tools/sharded_check.py)"""
    import torch
    rng = np.random.default_rng(19)
    nq, nt = 1300, 900
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    ctx.p2p_table_import(0, 1, [ctx.p2p_table_export(nq)])
    try:
        for rep in range(3):
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            dq, dt_ = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
            out = torch.full((3, nq), -7, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            ctx.bf_best2_sharded_p2p_dev(nq, dq.data_ptr(), nt, dt_.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
            ctx.synchronize()
            ob, os_, oi = O.bf_best2(q, t)
            g = out.cpu().numpy()
            assert np.array_equal(g[0], ob) and np.array_equal(g[1], os_) and np.array_equal(g[2], oi), rep
    finally:
        ctx.p2p_table_release()


def test_adaptors_known_answers(ctx):  # tests/test_measurement_adaptors.cpp:51,130
    from srrg2_proslam_b200 import capi
    e = capi.extract_cfg(5, 1, 500)
    for l, r, n in (("scene_flow_image_left.png", "scene_flow_image_right.png", 115),
                    ("kitti_city_image_left_0.png", "kitti_city_image_right_0.png", 177)):
        L, R = O.load_gray(l), O.load_gray(r)
        g = ctx.stereo_adaptor(L, R, e, capi.match_cfg(100, 0.8, 100, 0))
        o = O.stereo_adaptor(L, R, O.extract_cfg(5, 1, 500), "epipolar", 100, 0.8)
        assert len(g["uvuv"]) == n
        for k in ("uvuv", "intensity", "desc"):
            assert np.array_equal(g[k], o[k]), k


def kitti_chain():
    m0 = O.stereo_adaptor(O.load_gray("kitti_city_image_left_0.png"), O.load_gray("kitti_city_image_right_0.png"),
                          O.extract_cfg(threshold=15, target=500), "epipolar", 50, 0.8)
    xyz, _ = O.triangulate(m0["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
    return m0, xyz


@pytest.mark.parametrize("shape", ["circle", "square", "rhombus"])
@pytest.mark.parametrize("radius,dd,ratio", [(10, 50, 0.9), (25, 75, 0.8), (100, 25, 0.8), (3, 100, 0.99)])
def test_projective(ctx, feats, shape, radius, dd, ratio):
    from test_oracle_known_answers import CAM00, CAM01
    m0, xyz = kitti_chain()
    f1 = feats["L1"]
    pose = O.pose_inverse(O.pose_mul(O.pose_inverse(CAM00), CAM01))
    pf = O.ProjectiveFinder(K_KITTI, 376, 1241, shape=shape, max_desc_dist=dd, ratio=ratio, min_desc_dist=dd,
                            max_radius=radius, min_radius=radius, min_matching_ratio=0.0)
    pf.set_fixed(f1["xy"], f1["desc"])
    pf.set_moving(xyz, m0["desc"])
    pf.set_estimate(pose)
    o = pf.compute()
    ctx.projective_set_fixed(f1["xy"], f1["desc"])
    ctx.projective_set_moving(xyz, m0["desc"])
    gf, gm, gd, nproj = ctx.projective_match(pose, K_KITTI, 376, 1241, shape=shape, radius=radius,
                                             descriptor_distance=dd, ratio=ratio)
    assert nproj == pf.state()["n_projected"]
    assert len(gf) == len(o[0])
    if shape == "circle" and (radius, dd, ratio) == (10, 50, 0.9):
        assert len(gf) == 90  # tests/test_correspondence_finders.cpp:509
    assert same_corr((gf, gm, gd), o)  # including the unordered_map output order


@pytest.mark.parametrize("radius,dd,maxd,ratio", [(10, 50, 50, 0.9), (25, 25, 75, 0.8), (100, 25, 40, 0.8), (3, 100, 100, 0.99)])
def test_projective_kdtree(ctx, feats, radius, dd, maxd, ratio):
    """CorrespondenceFinderProjectiveKDTree (..._kdtree_impl.cpp:28-79) as the exact radius query (shape 3): GPU vs the CPU
    restatement, bit for bit incl. the unordered_map output order.  Parity with the reference itself is UNPINNED for this
    variant (approximate external KDTree); the exact query is its superset."""
    from test_oracle_known_answers import CAM00, CAM01
    m0, xyz = kitti_chain()
    f1 = feats["L1"]
    pose = O.pose_inverse(O.pose_mul(O.pose_inverse(CAM00), CAM01))
    pf = O.ProjectiveFinder(K_KITTI, 376, 1241, shape="kdtree", max_desc_dist=maxd, ratio=ratio, min_desc_dist=dd,
                            max_radius=radius, min_radius=radius, min_matching_ratio=0.0)
    pf.set_fixed(f1["xy"], f1["desc"])
    pf.set_moving(xyz, m0["desc"])
    pf.set_estimate(pose)
    o = pf.compute()
    ctx.projective_set_fixed(f1["xy"], f1["desc"])
    ctx.projective_set_moving(xyz, m0["desc"])
    g = ctx.projective_match(pose, K_KITTI, 376, 1241, shape="kdtree", radius=radius, descriptor_distance=dd, ratio=ratio,
                             max_descriptor_distance=maxd)
    assert g[3] == pf.state()["n_projected"]
    assert len(g[0]) == len(o[0]) and (radius < 10 or len(g[0]) > 10)
    assert same_corr(g[:3], o)
    # geometry: every match lies strictly inside the search radius of its projection (fp32 arithmetic of the query)
    uvz, idx = O.project(xyz, pose, K_KITTI, 376, 1241, 0.1, 1000.0)
    uv = np.full((len(xyz), 2), np.nan, np.float32)
    uv[idx] = uvz[:, :2]
    d2 = ((f1["xy"][g[0]] - uv[g[1]]) ** 2).sum(1)
    assert (d2 < radius * radius).all() and (g[2] < min(dd, maxd)).all()


def test_triangulate(ctx):  # mapping/triangulator_rigid_stereo.cpp:7-85
    m0, _ = kitti_chain()
    b_x = float(np.float32(718.856) * np.float32(0.537166))
    uvuv = m0["uvuv"].copy()
    uvuv[3, 2] = uvuv[3, 0]          # zero disparity -> infinity depth when the minimum disparity allows it
    uvuv[5, 2] = uvuv[5, 0] - 0.5    # below the default minimum disparity of 1 px -> INVALID placeholder
    for min_disp in (0.0, 1.0):
        g, valid, n_valid = ctx.triangulate(uvuv, K_KITTI, b_x, min_disp)
        o, n_inv = O.triangulate(uvuv, K_KITTI, b_x, min_disp)
        assert np.array_equal(g, o)  # bit-exact fp32, same operation order
        assert n_valid == len(uvuv) - n_inv == int(valid.sum())
        assert (min_disp == 0.0) == bool(valid[3]) and (min_disp == 0.0) == bool(valid[5])
    assert ctx.triangulate(np.zeros((0, 4), np.float32), K_KITTI, b_x)[2] == 0


def test_odd_feature_capacity(oracle):
    """a context created with an odd max_features (the capacity is a stride of 16 / 32-bit shared-memory arrays in the
    stereo matcher): whole stereo adaptor on the lean path (thickness 0) and with thickness 1"""
    from srrg2_proslam_b200 import capi
    L, R = O.load_gray("kitti_city_image_left_0.png"), O.load_gray("kitti_city_image_right_0.png")
    for mf in (1001, 1003, 999):
        c = capi.Context(max_images=2, max_rows=376, max_cols=1241, max_features=mf, max_raw_per_bin=8192)
        try:
            for thickness in (0, 1):
                g = c.stereo_adaptor(L, R, capi.extract_cfg(15, 1, 1000), capi.match_cfg(100, 0.5, 100, thickness))
                o = O.stereo_adaptor(L, R, O.extract_cfg(15, 1, 1000), "epipolar", 100, 0.5, 100, thickness)
                assert len(g["uvuv"]) > 100 and np.array_equal(g["uvuv"], o["uvuv"]) and np.array_equal(g["desc"], o["desc"])
        finally:
            c.close()


def test_epipolar_argument_checks(ctx, feats):
    """negative line thickness = the offset-0 pass only (epipolar_impl.cpp:72-79); coordinates that do not fit the packed
    (row, col) sort key are refused instead of silently ordered differently from the reference"""
    from srrg2_proslam_b200 import capi
    a, b = feats["L0"], feats["R0"]
    g0 = ctx.match_epipolar(a["xy"], a["desc"], b["xy"], b["desc"], capi.match_cfg(50, 0.9, 100, 0))
    gn = ctx.match_epipolar(a["xy"], a["desc"], b["xy"], b["desc"], capi.match_cfg(50, 0.9, 100, -3))
    assert len(g0[0]) > 50 and same_corr(g0, gn)
    bad = a["xy"].copy()
    bad[3, 0] = 70000.0
    with pytest.raises(capi.PslamError) as e:
        ctx.match_epipolar(bad, a["desc"], b["xy"], b["desc"], capi.match_cfg(50, 0.9, 100, 0))
    assert e.value.code == capi.PSLAM_E_INVALID
    bad = a["xy"].copy()
    bad[0, 1] = -1.0
    with pytest.raises(capi.PslamError):
        ctx.match_epipolar(a["xy"], a["desc"], bad, b["desc"], capi.match_cfg(50, 0.9, 100, 0))

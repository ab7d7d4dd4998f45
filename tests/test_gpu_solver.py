"""GPU parity, stages 3+4 (SE3 factor linearisation, H/b reduction, GN step) vs the fp64 oracle.
Tolerance: relative 1e-9 on H and b (BASELINE.json north_star), 1e-9 on the pose update."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float64)
RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(max_images=2, max_rows=64, max_cols=128, max_features=256, max_raw_per_bin=1024)
    yield c
    c.close()


def synth(n, kind, seed, outlier_frac=0.1):
    rng = np.random.default_rng(seed)
    xyz = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(3, 40, n)], 1)
    xyz[: n // 20, 2] = -1.0  # behind the camera -> suppressed
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.05, -0.02, 0.4])
    pc = xyz @ R.T + t
    h = pc @ K.reshape(3, 3).T
    u, v = h[:, 0] / h[:, 2], h[:, 1] / h[:, 2]
    noise = rng.normal(0, 0.5, (n, 3))
    out = rng.random(n) < outlier_frac
    noise[out] *= 40
    if kind == "stereo":
        meas = np.stack([u + noise[:, 0], v + noise[:, 1], (h[:, 0] - 386.1448) / h[:, 2] + noise[:, 2], v], 1)
    elif kind == "depth":
        meas = np.stack([u + noise[:, 0], v + noise[:, 1], pc[:, 2] + 0.01 * noise[:, 2]], 1)
    else:
        meas = np.stack([u + noise[:, 0], v + noise[:, 1]], 1)
    perm = rng.permutation(n)
    cf, cm = perm.astype(np.int32), np.arange(n, dtype=np.int32)
    meas_f = np.zeros_like(meas)
    meas_f[cf] = meas[cm]
    info = np.tile({"stereo": [1, 2, 1], "depth": [1, 1, 10], "mono": [1, 1, 0]}[kind], (n, 1)).astype(np.float64)
    info *= rng.uniform(0.5, 3.0, (n, 1))
    pose0 = np.concatenate([np.eye(3), np.zeros((3, 1))], 1).reshape(12)
    return xyz, meas_f, cf, cm, info, pose0


@pytest.mark.parametrize("kind", ["stereo", "depth", "mono"])
@pytest.mark.parametrize("n", [0, 1, 7, 300, 5000, 100000])
@pytest.mark.parametrize("robust,chi", [("saturated", 25.0), ("clamp", 10.0), ("none", 1.0)])
def test_linearize(ctx, kind, n, robust, chi):
    xyz, meas, cf, cm, info, pose = synth(max(n, 1), kind, 17 + n)
    cf, cm = cf[:n], cm[:n]
    md = 35.0 if kind == "stereo" and n % 2 else 0.0
    ocfg = O.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), md, robust, chi)
    gcfg = ctx.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), md, robust, chi)
    Ho, bo, so = O.linearize(ocfg, pose, xyz, meas, cf, cm, info)
    Hg, bg, sg = ctx.linearize(gcfg, pose, xyz, meas, cf, cm, info)
    assert (sg["inliers"], sg["outliers"], sg["suppressed"]) == (so["inliers"], so["outliers"], so["suppressed"])
    scale_H, scale_b = max(np.abs(Ho).max(), 1e-300), max(np.abs(bo).max(), 1e-300)
    assert np.abs(Hg - Ho).max() <= RTOL * scale_H
    assert np.abs(bg - bo).max() <= RTOL * scale_b
    assert abs(sg["chi"] - so["chi"]) <= RTOL * max(abs(so["chi"]), 1e-300)
    assert np.array_equal(Hg, Hg.T)


@pytest.mark.parametrize("kind,damping", [("stereo", 1.0), ("depth", 0.1), ("mono", 0.0)])
def test_gauss_newton_converges_like_oracle(ctx, kind, damping):
    xyz, meas, cf, cm, info, pose = synth(400, kind, 99, outlier_frac=0.05)
    ocfg = O.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    gcfg = ctx.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    po, pg = pose.copy(), pose.copy()
    for it in range(15):
        Ho, bo, _ = O.linearize(ocfg, po, xyz, meas, cf, cm, info)
        rc, po, dxo = O.gn_step(Ho, bo, damping, po)
        assert rc == 0
        Hg, bg, _ = ctx.linearize(gcfg, pg, xyz, meas, cf, cm, info)
        pg, dxg = ctx.gn_step(Hg, bg, damping, pg)
        assert np.abs(pg - po).max() < 1e-9
    # recovered the synthetic motion (t = (0.05,-0.02,0.4), 0.02 rad about y)
    assert abs(pg[3] - 0.05) < 0.05 and abs(pg[11] - 0.4) < 0.1


def test_gn_not_spd(ctx):
    from srrg2_proslam_b200 import capi
    with pytest.raises(capi.PslamError) as e:
        ctx.gn_step(np.zeros((6, 6)), np.ones(6), 0.0, np.eye(3, 4).reshape(12))
    assert e.value.code == capi.PSLAM_E_NOT_SPD


# sizes on both sides of the kernel's regimes: <= 96 shared-memory reduction, <= 256 one correspondence per thread, above: grid-stride
@pytest.mark.parametrize("kind,damping,n", [("stereo", 1.0, 400), ("depth", 0.1, 37), ("mono", 0.0, 3000), ("depth", 0.1, 96),
                                            ("stereo", 1.0, 97), ("stereo", 1.0, 150), ("mono", 0.0, 256), ("depth", 1.0, 257)])
def test_gn_iterate_fused(ctx, kind, damping, n):
    """pslam_gn_iterate: K solver iterations in one launch == K x (linearise + GN step) of the oracle"""
    xyz, meas, cf, cm, info, pose = synth(n, kind, 5 + n, outlier_frac=0.05)
    ocfg = O.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    gcfg = ctx.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    iters = 12
    pg, poses, stats, done, ok = ctx.gn_iterate(gcfg, iters, damping, pose, xyz, meas, cf, cm, info)
    assert done == iters and ok
    po = pose.copy()
    for it in range(iters):
        Ho, bo, so = O.linearize(ocfg, po, xyz, meas, cf, cm, info)
        rc, po, _ = O.gn_step(Ho, bo, damping, po)
        assert rc == 0
        assert np.abs(poses[it] - po).max() < 1e-9
        assert (int(stats[it, 1]), int(stats[it, 2]), int(stats[it, 3])) == (so["inliers"], so["outliers"], so["suppressed"])
        assert abs(stats[it, 0] - so["chi"]) <= RTOL * max(abs(so["chi"]), 1e-300)
    assert np.array_equal(pg, poses[-1])
    # zero iterations: nothing happens
    p0, _, _, d0, ok0 = ctx.gn_iterate(gcfg, 0, damping, pose, xyz, meas, cf, cm, info)
    assert d0 == 0 and ok0 and np.array_equal(p0, pose)


def test_gn_iterate_not_spd(ctx):
    xyz, meas, cf, cm, info, pose = synth(50, "stereo", 3)
    gcfg = ctx.linearize_cfg("stereo", K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    # no correspondences: H = 0, not positive definite without damping; the first iteration is linearised, not solved
    pg, poses, stats, done, ok = ctx.gn_iterate(gcfg, 5, 0.0, pose, xyz, meas, cf[:0], cm[:0], info)
    assert done == 1 and not ok and np.array_equal(pg, pose) and stats[0, 1] == 0


def rand_prior(seed):
    rng = np.random.default_rng(seed)
    v = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.05, 3)])
    _, Z, _ = O.gn_step(np.eye(6), -v, 0.0, np.eye(3, 4).reshape(12))
    M = rng.normal(size=(6, 6))
    return Z, 50.0 * (M @ M.T + np.eye(6))


@pytest.mark.parametrize("kind", ["stereo", "depth", "mono"])
@pytest.mark.parametrize("n", [0, 5, 300, 20000])
@pytest.mark.parametrize("with_prior", [False, True])
def test_linearize_f32_prior_status(ctx, kind, n, with_prior):
    """pslam_linearize_se3_f32: fp32 clouds in HBM (widened in registers), pose-prior factor summed on the device,
    per-correspondence factor status -- against the oracle on the SAME fp32 values"""
    xyz, meas, cf, cm, info, pose = synth(max(n, 1), kind, 31 + n)
    xyz, meas, info = xyz.astype(np.float32), meas.astype(np.float32), info.astype(np.float32)
    cf, cm = cf[:n], cm[:n]
    prior = rand_prior(n) if with_prior else None
    ocfg = O.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    gcfg = ctx.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    Ho, bo, so = O.linearize(ocfg, pose, xyz, meas, cf, cm, info, prior=prior, want_status=True)
    Hg, bg, sg = ctx.linearize_f32(gcfg, pose, xyz, meas, cf, cm, info, prior=prior)
    assert (sg["inliers"], sg["outliers"], sg["suppressed"]) == (so["inliers"], so["outliers"], so["suppressed"])
    assert np.array_equal(sg["status"], so["status"])
    assert np.abs(Hg - Ho).max() <= RTOL * max(np.abs(Ho).max(), 1e-300)
    assert np.abs(bg - bo).max() <= RTOL * max(np.abs(bo).max(), 1e-300)
    assert abs(sg["chi"] - so["chi"]) <= RTOL * max(abs(so["chi"]), 1e-300)
    assert abs(sg["prior_chi"] - so["prior_chi"]) <= RTOL * max(abs(so["prior_chi"]), 1e-300)


@pytest.mark.parametrize("kind,damping,n", [("stereo", 1.0, 400), ("depth", 0.1, 37), ("mono", 0.0, 3000), ("stereo", 1.0, 96),
                                            ("depth", 0.1, 97), ("stereo", 1.0, 256)])
def test_gn_iterate_f32_with_prior(ctx, kind, damping, n):
    """the fused launch with the motion-model slice's factor inside == K x (linearise + prior + GN step) of the oracle"""
    xyz, meas, cf, cm, info, pose = synth(n, kind, 5 + n, outlier_frac=0.05)
    xyz, meas, info = xyz.astype(np.float32), meas.astype(np.float32), info.astype(np.float32)
    prior = rand_prior(n)
    ocfg = O.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    gcfg = ctx.linearize_cfg(kind, K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    iters = 12
    pg, poses, stats, done, ok, status = ctx.gn_iterate_f32(gcfg, iters, damping, pose, xyz, meas, cf, cm, info, prior=prior)
    assert done == iters and ok
    po = pose.copy()
    for it in range(iters):
        Ho, bo, so = O.linearize(ocfg, po, xyz, meas, cf, cm, info, prior=prior, want_status=True)
        rc, po, _ = O.gn_step(Ho, bo, damping, po)
        assert rc == 0
        assert np.abs(poses[it] - po).max() < 1e-9
        assert (int(stats[it, 1]), int(stats[it, 2]), int(stats[it, 3])) == (so["inliers"], so["outliers"], so["suppressed"])
    assert np.array_equal(status, so["status"])  # status of the LAST linearised iteration
    # a strong prior dominates: the estimate ends at the prediction
    strong = (prior[0], 1e12 * np.eye(6))
    pg2, *_ = ctx.gn_iterate_f32(gcfg, 20, 0.0, pose, xyz, meas, cf, cm, info, prior=strong)
    assert np.abs(O.t2tnq(O.pose_mul(O.pose_inverse(prior[0]), pg2))).max() < 1e-4


@pytest.mark.parametrize("with_prior,weights", [(False, False), (True, True)])
def test_projective_match_gn_equals_match_then_iterate(oracle, with_prior, weights):
    """pslam_projective_match_gn (search + filter + K fused solver iterations in one device round trip) == pslam_projective_match
    followed by pslam_gn_iterate_f32 on the same correspondences in ascending fixed index, with the information the aligner
    slice would set up (diagonal x per-point scale); factor status comes back in the order of the returned correspondences"""
    from srrg2_proslam_b200 import capi
    from test_oracle_known_answers import CAM00, CAM01, K_KITTI
    c = capi.Context(max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
    try:
        e = O.extract_cfg(threshold=15, target=500)
        m = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"), e,
                              "epipolar", 50, 0.8) for i in (0, 1)]
        xyz, _ = O.triangulate(m[0]["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
        pose = O.pose_inverse(O.pose_mul(O.pose_inverse(CAM00), CAM01))
        guess = np.eye(3, 4).reshape(12)
        base = (K_KITTI.reshape(3, 3) @ np.array([-0.537166, 0, 0], np.float32)).astype(np.float64)
        lcfg = c.linearize_cfg("stereo", K_KITTI.astype(np.float64), 1241, 376, base, 0.0, "saturated", 1000.0)
        diag = np.array([1, 2, 1], np.float32)
        rng = np.random.default_rng(4)
        scale = (1 + rng.integers(0, 3, len(xyz))).astype(np.float32) if weights else np.ones(len(xyz), np.float32)
        prior = rand_prior(8) if with_prior else None
        c.projective_set_fixed(m[1]["uvuv"], m[1]["desc"])
        c.projective_set_moving(xyz, m[0]["desc"])
        if weights:
            c.projective_set_moving_weights(scale)
        f0, m0, d0, np0 = c.projective_match(pose, K_KITTI, 376, 1241, "circle", 30, 75.0, 0.8)
        f1, m1, d1, np1, g = c.projective_match_gn(pose, K_KITTI, 376, 1241, lcfg, diag, 7, 1.0, guess, "circle", 30, 75.0, 0.8,
                                                   prior=prior)
        assert np1 == np0 and np.array_equal(f0, f1) and np.array_equal(m0, m1) and np.array_equal(d0, d1) and len(f0) > 20
        order = np.argsort(f0)
        info = np.zeros((len(m[1]["uvuv"]), 3), np.float32)
        info[f0] = diag[None, :] * scale[m0][:, None]
        pg, poses, stats, done, ok, status = c.gn_iterate_f32(lcfg, 7, 1.0, guess, xyz, m[1]["uvuv"], f0[order], m0[order], info,
                                                             prior=prior)
        assert g["done"] == done == 7 and g["spd"] and ok
        assert np.array_equal(g["poses"], poses) and np.array_equal(g["stats"], stats) and np.array_equal(g["pose"], pg)
        assert np.array_equal(g["status"][order], status)
    finally:
        c.close()


def test_projective_align_equals_host_driven_phases(oracle):
    """pslam_projective_align (finder state machine on the device) == the same phases driven from the host with
    pslam_projective_match_gn: re-project when iteration % N == 0 or == 1, converge when the fp32 norm of
    t2tnq(X^-1 X_previous) drops below the threshold after the minimum number of calls, then spend the rest of the budget"""
    from srrg2_proslam_b200 import capi
    from test_oracle_known_answers import K_KITTI
    c = capi.Context(max_images=2, max_rows=376, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
    try:
        e = O.extract_cfg(threshold=15, target=500)
        m = [O.stereo_adaptor(O.load_gray(f"kitti_city_image_left_{i}.png"), O.load_gray(f"kitti_city_image_right_{i}.png"), e,
                              "epipolar", 50, 0.8) for i in (0, 1)]
        xyz, _ = O.triangulate(m[0]["uvuv"], K_KITTI, np.float32(718.856) * np.float32(0.537166), 0.0)
        guess = np.eye(3, 4).reshape(12)
        base = (K_KITTI.reshape(3, 3) @ np.array([-0.537166, 0, 0], np.float32)).astype(np.float64)
        lcfg = c.linearize_cfg("stereo", K_KITTI.astype(np.float64), 1241, 376, base, 0.0, "saturated", 1000.0)
        diag = np.array([1, 2, 1], np.float32)
        c.projective_set_fixed(m[1]["uvuv"], m[1]["desc"])
        c.projective_set_moving(xyz, m[0]["desc"])
        N, MIN_IT, BUDGET, THR = 5, 10, 60, 1e-5
        f, mm, d, a = c.projective_align(K_KITTI, 376, 1241, lcfg, diag, BUDGET, 1.0, guess, N, MIN_IT, THR, 0.1, 0, 10,
                                         shape="circle", radius=50, descriptor_distance=75.0, ratio=0.8)
        assert a["stop_reason"] == 1 and a["done"] == BUDGET and a["spd"] and len(a["phases"]) >= 4
        # host-driven replay
        est, prev, ci, it, conv = guess.copy(), np.eye(3, 4, dtype=np.float32), 0, 0, False
        poses, stats, phases = [], [], []
        while it < BUDGET:
            X = est.astype(np.float32).reshape(3, 4)
            E = O.pose_mul(O.pose_inverse(X.reshape(12).astype(np.float64)), prev.reshape(12).astype(np.float64))
            norm = float(np.linalg.norm(O.t2tnq(E)))
            conv = norm < THR and ci > MIN_IT
            quiet = 0
            k = ci + 1
            while not (k % N == 0 or k == 1):
                quiet, k = quiet + 1, k + 1
            n_fused = BUDGET - it if conv else min(BUDGET - it, quiet + 1)
            f1, m1, d1, _, g = c.projective_match_gn(X.reshape(12), K_KITTI, 376, 1241, lcfg, diag, n_fused, 1.0, est, "circle", 50,
                                                     75.0, 0.8)
            assert g["done"] == n_fused
            phases.append((it, n_fused, len(f1)))
            poses.append(g["poses"])
            stats.append(g["stats"])
            prev = X  # the search call keeps the pose it was given; the n_fused - 1 calls that follow (none once converged) do too
            if not conv and n_fused >= 2:
                prev = g["poses"][n_fused - 2].astype(np.float32).reshape(3, 4)
            ci += 1 if conv else n_fused
            est, it = g["pose"], it + n_fused
        assert a["phases"] == phases
        assert np.array_equal(a["poses"], np.concatenate(poses)) and np.array_equal(a["stats"], np.concatenate(stats))
        assert np.array_equal(f, f1) and np.array_equal(mm, m1) and np.array_equal(d, d1)
        assert a["current_iteration"] == ci and a["has_converged"] == conv
        # argument validation
        with pytest.raises(Exception):
            c.projective_align(K_KITTI, 376, 1241, lcfg, diag, BUDGET, 1.0, guess, N, MIN_IT, has_converged=1)
        with pytest.raises(Exception):
            c.projective_align(K_KITTI, 376, 1241, lcfg, diag, 0, 1.0, guess, N, MIN_IT)
        # decisions left to the caller: nothing is applied, the state comes back as it went in
        _, _, _, b = c.projective_align(K_KITTI, 376, 1241, lcfg, diag, BUDGET, 1.0, guess, N, MIN_IT, THR, 0.99, 1, 10, shape="circle",
                                        radius=50, descriptor_distance=75.0, ratio=0.8)
        assert b["stop_reason"] == 2 and b["done"] == 0 and b["phases"] == [] and b["current_iteration"] == 0
        _, _, _, b = c.projective_align(K_KITTI, 376, 1241, lcfg, diag, BUDGET, 1.0, guess, N, MIN_IT, THR, 0.1, 0, 5000, shape="circle",
                                        radius=50, descriptor_distance=75.0, ratio=0.8)
        assert b["stop_reason"] == 3 and b["done"] == 0 and np.array_equal(b["pose"], guess)
    finally:
        c.close()

"""N2: SceneClipperProjective3D on the GPU -- pslam_scene_clip / pslam_scene_clip_dev against the CPU
oracle (bit-exact: fp32 projector arithmetic, map order, global indices, copied descriptors) and against the reference's
known survivor counts (tests/test_scene_clippers.cpp)."""
import numpy as np
import pytest

import oracle_lib as O
import scene_fixtures as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(oracle):
    from srrg2_proslam_b200 import capi
    c = capi.Context(device=0, max_images=2, max_rows=480, max_cols=1241, max_features=2048, max_raw_per_bin=8192)
    yield c
    c.close()


@pytest.fixture(scope="module")
def clouds(oracle):
    return {"icl_sparse": F.icl_sparse(), "icl_dense": (F.icl_dense(), None), "kitti": F.kitti_sparse()}


def check(ctx, xyz, desc, T, K, rows, cols, rmin, rmax, S=None, expect=None):
    from srrg2_proslam_b200 import capi
    g = ctx.scene_clip(xyz, capi.clip_cfg(K, rows, cols, T, rmin, rmax, S), desc=desc)
    oxyz, ouvz, oidx = O.scene_clip(xyz, T, K, rows, cols, rmin, rmax, sensor_in_robot=S)
    assert len(g["index"]) == len(oidx)
    if expect is not None:
        assert len(oidx) == expect
    assert np.array_equal(g["index"], oidx)
    assert np.array_equal(g["xyz"], oxyz) and np.array_equal(g["uvz"], ouvz)  # bit exact fp32
    if desc is not None:
        assert np.array_equal(g["desc"], desc[oidx])
    return g


@pytest.mark.parametrize("name,T,n_sparse,n_dense", F.ICL_CASES, ids=[c[0] for c in F.ICL_CASES])
def test_icl_cases(ctx, clouds, name, T, n_sparse, n_dense):  # tests/test_scene_clippers.cpp:7-391
    xyz, desc = clouds["icl_sparse"]
    check(ctx, xyz, desc, T, F.K_ICL, 480, 640, 0.1, 10.0, expect=n_sparse)
    check(ctx, clouds["icl_dense"][0], None, T, F.K_ICL, 480, 640, 0.1, 10.0, expect=n_dense)


@pytest.mark.parametrize("name,T,n", F.KITTI_CASES, ids=[c[0] for c in F.KITTI_CASES])
def test_kitti_cases(ctx, clouds, name, T, n):  # tests/test_scene_clippers.cpp:393-460
    xyz, desc = clouds["kitti"]
    g = check(ctx, xyz, desc, T, F.K_KITTI, 376, 1241, 0.1, 1000.0, expect=n)
    assert np.all(g["xyz"][:, 2] > 0)


def test_sensor_in_robot(ctx, clouds):
    xyz, desc = clouds["icl_sparse"]
    S = F.pose(F.rot_z(0.3), (0.1, -0.2, 0.05))
    check(ctx, xyz, desc, F.pose(F.rot_x(0.1), (0.05, 0, 0.5)), F.K_ICL, 480, 640, 0.1, 10.0, S=S)


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 100003, 1 << 20])
def test_synthetic_map(ctx, n):
    """random local map around the camera: about a third of the points is visible; tile boundaries (2048) included"""
    rng = np.random.default_rng(n)
    xyz = rng.uniform(-30, 30, (n, 3)).astype(np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    T = F.pose(F.rot_x(0.05) @ F.rot_z(-0.02), (0.3, -0.1, 1.5))
    g = check(ctx, xyz, desc, T, F.K_KITTI, 376, 1241, 0.1, 1000.0)
    assert np.all(np.diff(g["index"]) > 0)  # map order
    # idempotence: the survivors (already in the camera frame) clipped again from the identity all survive
    g2 = ctx_clip_identity(ctx, g["xyz"])
    assert len(g2["index"]) == len(g["index"]) and np.array_equal(g2["uvz"], g["uvz"])


def ctx_clip_identity(ctx, xyz):
    from srrg2_proslam_b200 import capi
    return ctx.scene_clip(xyz, capi.clip_cfg(F.K_KITTI, 376, 1241, np.eye(3, 4), 0.1, 1000.0))


def test_empty_none_visible_and_capacity(ctx):
    from srrg2_proslam_b200 import capi
    cfg = capi.clip_cfg(F.K_KITTI, 376, 1241, np.eye(3, 4), 0.1, 1000.0)
    assert len(ctx.scene_clip(np.zeros((0, 3), np.float32), cfg)["index"]) == 0
    behind = np.tile(np.array([[0, 0, -5]], np.float32), (5000, 1))
    assert len(ctx.scene_clip(behind, cfg)["index"]) == 0
    front = np.tile(np.array([[0, 0, 5]], np.float32), (5000, 1))
    assert len(ctx.scene_clip(front, cfg)["index"]) == 5000
    with pytest.raises(capi.PslamError) as e:
        ctx.scene_clip(front, cfg, capacity=100)
    assert e.value.code == capi.PSLAM_E_CAPACITY


def test_device_resident(ctx):
    """pslam_scene_clip_dev: map and outputs stay in HBM (torch tensors are plumbing), 4 M points"""
    import torch
    from srrg2_proslam_b200 import capi
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(1)
    xyz = (torch.rand((n, 3), generator=g, device="cuda") * 60 - 30).contiguous()
    desc = torch.randint(0, 2 ** 31 - 1, (n, 8), generator=g, device="cuda", dtype=torch.int32).contiguous()
    oxyz, ouvz = torch.empty((n, 3), device="cuda"), torch.empty((n, 3), device="cuda")
    oidx = torch.empty(n, dtype=torch.int32, device="cuda")
    odesc = torch.empty((n, 8), dtype=torch.int32, device="cuda")
    T = F.pose(F.rot_x(0.05), (0.3, -0.1, 1.5))
    cfg = capi.clip_cfg(F.K_KITTI, 376, 1241, T, 0.1, 1000.0)
    torch.cuda.synchronize()
    m, ms = ctx.scene_clip_dev(n, xyz.data_ptr(), desc.data_ptr(), cfg, oxyz.data_ptr(), ouvz.data_ptr(), oidx.data_ptr(),
                               odesc.data_ptr(), reps=3)
    exp_xyz, exp_uvz, exp_idx = O.scene_clip(xyz.cpu().numpy(), T, F.K_KITTI, 376, 1241, 0.1, 1000.0)
    assert m == len(exp_idx) and ms > 0
    assert np.array_equal(oidx[:m].cpu().numpy(), exp_idx)
    assert np.array_equal(oxyz[:m].cpu().numpy(), exp_xyz) and np.array_equal(ouvz[:m].cpu().numpy(), exp_uvz)
    assert torch.equal(odesc[:m], desc[torch.from_numpy(exp_idx).cuda().long()])
    # a map that does not start on a 16-byte boundary (row 1 of the tensor: +12 bytes) takes the non-TMA staging path
    n2 = 100001
    sub = xyz[1:1 + n2]
    assert sub.data_ptr() % 16 != 0
    m2, _ = ctx.scene_clip_dev(n2, sub.data_ptr(), 0, cfg, oxyz.data_ptr(), ouvz.data_ptr(), oidx.data_ptr(), 0, reps=1)
    e2 = O.scene_clip(sub.cpu().numpy(), T, F.K_KITTI, 376, 1241, 0.1, 1000.0)
    assert m2 == len(e2[2]) and np.array_equal(oidx[:m2].cpu().numpy(), e2[2]) and np.array_equal(oxyz[:m2].cpu().numpy(), e2[0])


def test_conf_clipper_module(oracle):
    """kitti.conf "clipper_stereo_projective" (SceneClipperProjective3D -> projector #5) instantiated by class name and
    driven through the SceneClipper_ contract; tests/test_scene_clippers.cpp:393-460 and the error texts of
    scene_clipper_projective_3d.cpp:12-20"""
    import pathlib
    from srrg2_proslam_b200 import plugin as P
    m = P.Manager(pathlib.Path(__file__).resolve().parent / "golden" / "configurations" / "kitti_hotpath.conf")
    cl = m.get("clipper_stereo_projective")
    assert cl.class_name == "SceneClipperProjective3D" and not cl.is_generic
    pr = cl.link("projector")
    pr.set_camera_matrix(F.K_KITTI)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    pr.set("range_min", 0.1).set("range_max", 1000.0)
    with pytest.raises(P.PluginError, match="missing global scene"):
        cl.clipper_compute()
    xyz, desc = F.kitti_sparse()
    cl.clipper_set_full_scene(xyz, desc)
    for name, T, n in F.KITTI_CASES:
        cl.clipper_set_robot_in_local_map(T)
        g = cl.clipper_compute()
        oxyz, ouvz, oidx = O.scene_clip(xyz, T, F.K_KITTI, 376, 1241, 0.1, 1000.0)
        assert g["status"] == 2 and len(g["index"]) == n
        assert np.array_equal(g["index"], oidx) and np.array_equal(g["xyz"], oxyz) and np.array_equal(g["uvz"], ouvz)
        assert np.array_equal(g["desc"], desc[oidx])
    # sensor_in_robot != identity: the clipped scene is expressed in the robot frame, the camera sits at robot * sensor
    S = F.pose(F.rot_z(0.2), (0.5, 0.1, -0.3))
    R = F.pose(t=(0, 0, 2))
    cl.clipper_set_sensor_in_robot(S)
    cl.clipper_set_robot_in_local_map(R)
    g = cl.clipper_compute()
    cam = O.pose_mul(R, S).astype(np.float32)
    oxyz, ouvz, oidx = O.scene_clip(xyz, cam, F.K_KITTI, 376, 1241, 0.1, 1000.0, sensor_in_robot=S)
    assert np.array_equal(g["index"], oidx) and np.allclose(g["xyz"], oxyz, atol=1e-5) and np.allclose(g["uvz"], ouvz, atol=1e-3)
    # empty scene: Status::Ready, nothing clipped (scene_clipper_projective_3d.cpp:21-29)
    cl.clipper_set_full_scene(np.zeros((0, 3), np.float32), np.zeros((0, 32), np.uint8))
    g = cl.clipper_compute()
    assert g["status"] == 1 and len(g["index"]) == 0

"""ctypes binding of the CPU oracle (oracle/libpslam_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(srrg2_proslam_b200/) never imports this module.
"""
import ctypes as C
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"

_lib = None


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        so = ORACLE_DIR / "libpslam_oracle.so"
        if not so.exists():
            build()
        _lib = C.CDLL(str(so))
        _lib.orc_pf_create.restype = C.c_void_p
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _img(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


def extract_cfg(threshold=10, nms=1, target=500, nh=3, nv=3):
    return np.array([threshold, nms, target, nh, nv], dtype=np.float32)


def load_gray(name):
    import cv2
    im = cv2.imread(str(GOLDEN / name), cv2.IMREAD_UNCHANGED)
    assert im is not None, name
    return im


def load_depth(name):
    """16-bit depth fixtures (re-encoded losslessly from the reference's .pgm by tools/make_golden.py)"""
    im = load_gray(name)
    assert im.dtype == np.uint16, name
    return im


def fast_detect(img, thr, nms=True, mask=None, cap=200000):
    img = _img(img)
    xy = np.zeros((cap, 2), np.float32)
    resp = np.zeros(cap, np.float32)
    m = None if mask is None else _img(mask)
    n = lib().orc_fast_detect(_p(img), img.shape[0], img.shape[1], img.shape[1], int(thr), int(nms),
                              _p(m), cap, _p(xy), _p(resp))
    assert n <= cap
    return xy[:n].copy(), resp[:n].copy()


def blur7(img):
    img = _img(img)
    out = np.zeros_like(img)
    lib().orc_blur7(_p(img), img.shape[0], img.shape[1], img.shape[1], _p(out))
    return out


def detect_binned(img, cfg, cap=200000):
    img = _img(img)
    xy = np.zeros((cap, 2), np.float32)
    resp = np.zeros(cap, np.float32)
    n = lib().orc_detect_binned(_p(img), img.shape[0], img.shape[1], img.shape[1], _p(cfg), cap,
                                _p(xy), _p(resp))
    assert n <= cap
    return xy[:n].copy(), resp[:n].copy()


def extract_binned(img, cfg, mask=None, cap=100000):
    """-> dict(xy [N,2] f32, response [N], intensity [N], desc [N,32] u8)"""
    img = _img(img)
    xy = np.zeros((cap, 2), np.float32)
    resp = np.zeros(cap, np.float32)
    inten = np.zeros(cap, np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    m = None if mask is None else _img(mask)
    n = lib().orc_extract_binned(_p(img), img.shape[0], img.shape[1], img.shape[1], _p(cfg), _p(m),
                                 cap, _p(xy), _p(resp), _p(inten), _p(desc))
    assert n <= cap
    return dict(xy=xy[:n].copy(), response=resp[:n].copy(), intensity=inten[:n].copy(),
                desc=desc[:n].copy())


def bin_lut(rows, cols, nh, nv, target):
    lut = np.zeros((rows, cols), np.int32)
    q = C.c_int64(0)
    lib().orc_bin_lut(rows, cols, nh, nv, target, _p(lut), C.byref(q))
    return lut, q.value


def std_sort(keys, descending=True):
    """permutation libstdc++ std::sort produces for these keys (payload = original index)"""
    keys = np.ascontiguousarray(keys, np.float32)
    perm = np.arange(len(keys), dtype=np.int32)
    fn = lib().orc_std_sort_desc if descending else lib().orc_std_sort_asc
    fn(len(keys), _p(keys), _p(perm))
    return perm


def hamming_matrix(df, dm):
    df = np.ascontiguousarray(df, np.uint8).reshape(-1, 32)
    dm = np.ascontiguousarray(dm, np.uint8).reshape(-1, 32)
    out = np.zeros((len(df), len(dm)), np.int32)
    lib().orc_hamming_matrix(len(df), _p(df), len(dm), _p(dm), _p(out))
    return out


def bf_best2(df, dm, threads=1):
    df = np.ascontiguousarray(df, np.uint8).reshape(-1, 32)
    dm = np.ascontiguousarray(dm, np.uint8).reshape(-1, 32)
    best = np.zeros(len(df), np.int32)
    second = np.zeros(len(df), np.int32)
    idx = np.zeros(len(df), np.int32)
    lib().orc_bf_best2(len(df), _p(df), len(dm), _p(dm), int(threads), _p(best), _p(second), _p(idx))
    return best, second, idx


def _corr_out(cap):
    return np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)


def match_epipolar(xy_f, desc_f, xy_m, desc_m, max_dist=50.0, ratio=0.9, max_disp=100, thickness=0):
    xy_f = np.ascontiguousarray(xy_f, np.float32).reshape(-1, 2)
    xy_m = np.ascontiguousarray(xy_m, np.float32).reshape(-1, 2)
    desc_f = np.ascontiguousarray(desc_f, np.uint8).reshape(-1, 32)
    desc_m = np.ascontiguousarray(desc_m, np.uint8).reshape(-1, 32)
    cap = max(len(xy_f), 1) * (2 * thickness + 1)
    fi, mi, d = _corr_out(cap)
    n = lib().orc_match_epipolar(len(xy_f), _p(xy_f), _p(desc_f), len(xy_m), _p(xy_m), _p(desc_m),
                                 C.c_float(max_dist), C.c_float(ratio), int(max_disp),
                                 int(thickness), cap, _p(fi), _p(mi), _p(d))
    assert n <= cap
    return fi[:n].copy(), mi[:n].copy(), d[:n].copy()


def match_bruteforce(desc_f, desc_m, max_dist=50.0, ratio=0.9):
    desc_f = np.ascontiguousarray(desc_f, np.uint8).reshape(-1, 32)
    desc_m = np.ascontiguousarray(desc_m, np.uint8).reshape(-1, 32)
    cap = max(min(len(desc_f), len(desc_m)), 1)
    fi, mi, d = _corr_out(cap)
    n = lib().orc_match_bruteforce(len(desc_f), _p(desc_f), len(desc_m), _p(desc_m),
                                   C.c_float(max_dist), C.c_float(ratio), cap, _p(fi), _p(mi), _p(d))
    assert n <= cap
    return fi[:n].copy(), mi[:n].copy(), d[:n].copy()


def stereo_adaptor(left, right, cfg, matcher="epipolar", max_dist=50.0, ratio=0.9, max_disp=100,
                   thickness=0, cap=20000):
    left, right = _img(left), _img(right)
    mc = np.array([max_dist, ratio, max_disp, thickness], np.float32)
    uvuv = np.zeros((cap, 4), np.float32)
    inten = np.zeros(cap, np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    nl, nr, nm = C.c_int(0), C.c_int(0), C.c_int(0)
    n = lib().orc_stereo_adaptor(_p(left), _p(right), left.shape[0], left.shape[1], left.shape[1],
                                 _p(cfg), 0 if matcher == "epipolar" else 1, _p(mc), cap, _p(uvuv),
                                 _p(inten), _p(desc), C.byref(nl), C.byref(nr), C.byref(nm))
    assert n <= cap
    return dict(uvuv=uvuv[:n].copy(), intensity=inten[:n].copy(), desc=desc[:n].copy(),
                n_left=nl.value, n_right=nr.value, n_matches=nm.value)


def stereo_frontend_batch(images, cfg, max_dist=50.0, ratio=0.9, max_disp=100, thickness=0, threads=1):
    """images: u8 [n_pairs, 2, rows, cols]; returns (counts int32[n_pairs], checksum uint64[n_pairs])"""
    images = np.ascontiguousarray(images, np.uint8)
    n, two, rows, cols = images.shape
    assert two == 2
    m = np.array([max_dist, ratio, max_disp, thickness], np.float32)
    counts = np.zeros(n, np.int32)
    chk = np.zeros(n, np.uint64)
    lib().orc_stereo_frontend_batch(_p(images), n, rows, cols, cols, C.c_longlong(rows * cols), _p(cfg), _p(m),
                                    int(threads), _p(counts), _p(chk))
    return counts, chk


def fnv1a_points(uvuv, intensity, desc):
    """the per-pair checksum of orc_stereo_frontend_batch, for results obtained elsewhere (vectorised)"""
    h = 1469598103934665603
    n = len(uvuv)
    if n == 0:
        return np.uint64(h)
    rec = np.concatenate([np.ascontiguousarray(uvuv, np.float32).view(np.uint8).reshape(n, 16),
                          np.ascontiguousarray(intensity, np.float32).view(np.uint8).reshape(n, 4),
                          np.ascontiguousarray(desc, np.uint8).reshape(n, 32)], axis=1).reshape(-1)
    M = (1 << 64) - 1
    for b in rec.tolist():
        h = ((h ^ b) * 1099511628211) & M
    return np.uint64(h)


def mono_depth_adaptor(img, depth, cfg, depth_scale=1.0, cap=20000):
    img = _img(img)
    is_float = depth.dtype == np.float32
    depth = np.ascontiguousarray(depth, np.float32 if is_float else np.uint16)
    uvd = np.zeros((cap, 3), np.float32)
    inten = np.zeros(cap, np.float32)
    desc = np.zeros((cap, 32), np.uint8)
    nf = C.c_int(0)
    n = lib().orc_mono_depth_adaptor(_p(img), img.shape[0], img.shape[1], img.shape[1], _p(depth),
                                     int(is_float), depth.shape[1], C.c_float(depth_scale), _p(cfg),
                                     cap, _p(uvd), _p(inten), _p(desc), C.byref(nf))
    return dict(uvd=uvd[:n].copy(), intensity=inten[:n].copy(), desc=desc[:n].copy(),
                n_features=nf.value)


def triangulate(uvuv, K, b_x, min_disparity=1.0, infinity_depth=None):
    uvuv = np.ascontiguousarray(uvuv, np.float32).reshape(-1, 4)
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    if infinity_depth is None:
        infinity_depth = float(np.sqrt(np.finfo(np.float32).max))
    xyz = np.zeros((len(uvuv), 3), np.float32)
    ninv = C.c_int(0)
    lib().orc_triangulate(len(uvuv), _p(uvuv), _p(K), C.c_float(b_x), C.c_float(min_disparity),
                          C.c_float(infinity_depth), _p(xyz), C.byref(ninv))
    return xyz, ninv.value


def project(xyz, pose12, K, rows, cols, rmin=0.1, rmax=1000.0):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    uvz = np.zeros((len(xyz), 3), np.float32)
    idx = np.zeros(len(xyz), np.int32)
    n = lib().orc_project(len(xyz), _p(xyz), _p(pose12), _p(K), rows, cols, C.c_float(rmin),
                          C.c_float(rmax), _p(uvz), _p(idx))
    return uvz[:n].copy(), idx[:n].copy()


def scene_clip(xyz, camera_in_map, K, rows, cols, rmin=0.1, rmax=1000.0, sensor_in_robot=None):
    """SceneClipperProjective3D::compute (mapping/scene_clipper_projective_3d.cpp:9-67)"""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    T = np.ascontiguousarray(camera_in_map, np.float32).reshape(12)
    S = None if sensor_in_robot is None else np.ascontiguousarray(sensor_in_robot, np.float32).reshape(12)
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    n = len(xyz)
    oxyz, ouvz, oidx = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
    m = lib().orc_scene_clip(n, _p(xyz), _p(T), _p(S) if S is not None else None, _p(K), rows, cols, C.c_float(rmin),
                             C.c_float(rmax), _p(oxyz), _p(ouvz), _p(oidx))
    return oxyz[:m].copy(), ouvz[:m].copy(), oidx[:m].copy()


EKF_KINDS = {"projective": 0, "projective_depth": 1, "stereo": 2}
EKF_DIMS = {"projective": 2, "projective_depth": 3, "stereo": 4}


def point_ekf(kind, cam6, world_in_sensor, Q, meas, Rm, state, cov):
    """PointEKFBase::compute (predict + correct) of one point in double; returns the new (state, covariance)"""
    E = EKF_DIMS[kind]
    a = lambda x, n: np.ascontiguousarray(x, np.float64).reshape(n)
    state, cov = a(state, 3).copy(), a(cov, 9).copy()
    rc = lib().orc_point_ekf(EKF_KINDS[kind], _p(a(cam6, 6)), _p(a(world_in_sensor, 12)), _p(a(Q, 9)), _p(a(meas, E)),
                             _p(a(Rm, E * E)), _p(state), _p(cov))
    assert rc == 0
    return state, cov.reshape(3, 3)


def landmarks_ekf_update(kind, K, baseline2, sensor_in_world, sensor_in_local_map, state_world, covariance, meas,
                         min_cov=0.01, max_cov_norm2=1.0, max_dist2=1.0):
    """LandmarkEstimatorEKF_::compute over n landmarks (mapping/landmarks/landmark_estimator_ekf_impl.cpp:17-82);
    returns (new state_world, new covariance, coords_in_local_map, inlier)"""
    E = EKF_DIMS[kind]
    st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
    n = len(st)
    cv = np.ascontiguousarray(covariance, np.float32).reshape(n, 9).copy()
    ms = np.ascontiguousarray(meas, np.float32).reshape(n, E)
    loc = np.zeros((n, 3), np.float32)
    inl = np.zeros(n, np.uint8)
    f32 = lambda x, k: np.ascontiguousarray(x, np.float32).reshape(k)
    b = np.ascontiguousarray(baseline2, np.float64).reshape(2)
    lib().orc_landmarks_ekf_update(n, EKF_KINDS[kind], _p(f32(K, 9)), _p(b), C.c_double(min_cov), C.c_double(max_cov_norm2),
                                   C.c_float(max_dist2), _p(f32(sensor_in_world, 12)), _p(f32(sensor_in_local_map, 12)),
                                   _p(st), _p(cv), _p(ms), _p(loc), _p(inl))
    return st, cv.reshape(n, 3, 3), loc, inl.astype(bool)


def landmarks_weighted_mean_update(sensor_in_world, sensor_in_local_map, state_world, n_opt, landmark_in_sensor, max_dist2=1.0):
    """LandmarkEstimatorWeightedMean_::compute over n landmarks (landmark_estimator_weighted_mean_impl.cpp:7-41)"""
    st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
    n = len(st)
    no = np.ascontiguousarray(n_opt, np.int32).reshape(n)
    ls = np.ascontiguousarray(landmark_in_sensor, np.float32).reshape(n, 3)
    loc, inl = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
    f32 = lambda x: np.ascontiguousarray(x, np.float32).reshape(12)
    lib().orc_landmarks_weighted_mean_update(n, C.c_float(max_dist2), _p(f32(sensor_in_world)), _p(f32(sensor_in_local_map)), _p(st),
                                             _p(no), _p(ls), _p(loc), _p(inl))
    return st, loc, inl.astype(bool)


def landmarks_smoother_update(K, frames_sensor_in_world, sensor_in_world, sensor_in_local_map, offsets, hist_frame, hist_uv,
                              hist_point_in_camera, state_world, n_opt, max_iterations=100, chi2_delta=1e-5, max_reproj2=100.0,
                              min_measurements=3, max_dist2=1.0):
    """LandmarkEstimatorPoseBasedSmoother_::compute over n landmarks (landmark_estimator_pose_based_smoother_impl.cpp:6-148);
    the histories (CSR) already contain the current measurement.  Returns (state_world, n_opt, coords_in_local_map, inlier)."""
    st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
    n = len(st)
    no = np.ascontiguousarray(n_opt, np.int32).reshape(n).copy()
    fr = np.ascontiguousarray(frames_sensor_in_world, np.float32).reshape(-1, 12)
    off = np.ascontiguousarray(offsets, np.int32).reshape(n + 1)
    hf = np.ascontiguousarray(hist_frame, np.int32)
    uv = np.ascontiguousarray(hist_uv, np.float32).reshape(len(hf), 2)
    pic = np.ascontiguousarray(hist_point_in_camera, np.float32).reshape(len(hf), 3)
    par = np.array([max_iterations, chi2_delta, max_reproj2, min_measurements, max_dist2], np.float32)
    loc, inl = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
    f32 = lambda x, k: np.ascontiguousarray(x, np.float32).reshape(k)
    lib().orc_landmarks_smoother_update(n, _p(f32(K, 9)), _p(par), len(fr), _p(fr), _p(f32(sensor_in_world, 12)),
                                        _p(f32(sensor_in_local_map, 12)), _p(off), _p(hf), _p(uv), _p(pic), _p(st), _p(no), _p(loc), _p(inl))
    return st, no, loc, inl.astype(bool)


MERGER_KINDS = {"base": 0, "stereo": 1, "depth": 2}


def _merger_cfg6(canvas_rows, canvas_cols, row_bins, col_bins, enable_binning, kind):
    return np.array([canvas_rows, canvas_cols, row_bins, col_bins, int(bool(enable_binning)), MERGER_KINDS[kind]], np.int32)


def merger_select_updates(measurements, corr_moving, corr_response, canvas_rows, canvas_cols, row_bins=10, col_bins=30,
                          max_distance_appearance=50.0, enable_binning=True, kind="stereo"):
    """MergerProjective_::compute update pass (merger_projective_impl.cpp:61-135), the sequential walk:
    (selected[n_corr] bool, blocked-bin bitmap words)"""
    m = np.ascontiguousarray(measurements, np.float32)
    mv = np.ascontiguousarray(corr_moving, np.int32).reshape(-1)
    rs = np.ascontiguousarray(corr_response, np.float32).reshape(len(mv))
    sel = np.zeros(max(len(mv), 1), np.uint8)
    occ = np.zeros(((row_bins + 1) * (col_bins + 1) + 31) // 32, np.uint32)
    k = lib().orc_merger_select_updates(_p(_merger_cfg6(canvas_rows, canvas_cols, row_bins, col_bins, enable_binning, kind)),
                                        C.c_float(max_distance_appearance), _p(m), m.shape[1], _p(mv), _p(rs), len(mv), _p(sel),
                                        _p(occ), len(occ))
    if k < 0:
        raise ValueError("measurement outside the bin grid")
    return sel[:len(mv)].astype(bool), occ


def merger_select_additions(measurements, occupied, canvas_rows, canvas_cols, row_bins=10, col_bins=30, enable_binning=True,
                            kind="stereo"):
    """MergerProjective_::_addPoints binning (merger_projective_impl.cpp:205-253): source measurement index of every entry of
    points_in_image_to_add, in its order"""
    m = np.ascontiguousarray(measurements, np.float32)
    win = np.zeros(max(len(m), 1), np.int32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint32)
    k = lib().orc_merger_select_additions(_p(_merger_cfg6(canvas_rows, canvas_cols, row_bins, col_bins, enable_binning, kind)), _p(m),
                                          m.shape[1], len(m), None if occ is None else _p(occ), _p(win))
    return win[:k].copy()


SHAPES = {"square": 0, "circle": 1, "rhombus": 2, "kdtree": 3}


class ProjectiveFinder:
    """Stateful oracle of CorrespondenceFinderProjective{Square,Circle,Rhombus}."""

    def __init__(self, K, rows, cols, shape="circle", max_desc_dist=50.0, ratio=0.9,
                 min_matching_ratio=0.25, min_desc_dist=25.0, desc_step=5.0, max_radius=100,
                 min_radius=10, radius_step=5, min_iterations=10, max_change_norm=1e-5,
                 iters_per_projection=25, range_min=0.1, range_max=1000.0):
        f = np.array([max_desc_dist, ratio, min_matching_ratio, min_desc_dist, desc_step, max_radius,
                      min_radius, radius_step, min_iterations, max_change_norm, iters_per_projection,
                      SHAPES[shape]], np.float32)
        p = np.concatenate([np.asarray(K, np.float32).reshape(9),
                            np.array([rows, cols, range_min, range_max], np.float32)])
        self.h = C.c_void_p(lib().orc_pf_create(_p(f), _p(p)))
        self.cap = 1

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pf_destroy(self.h)
            self.h = None

    def set_fixed(self, coords, desc):
        coords = np.ascontiguousarray(coords, np.float32)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.cap = max(self.cap, len(coords) + 1)
        lib().orc_pf_set_fixed(self.h, len(coords), _p(coords), coords.shape[1], _p(desc))

    def set_moving(self, xyz, desc):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.ncap = len(xyz) + 1
        lib().orc_pf_set_moving(self.h, len(xyz), _p(xyz), _p(desc))

    def set_estimate(self, pose12):
        pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
        lib().orc_pf_set_estimate(self.h, _p(pose12))

    def get_estimate(self):
        p = np.zeros(12, np.float32)
        lib().orc_pf_get_estimate(self.h, _p(p))
        return p

    def set_radius(self, r):
        lib().orc_pf_set_radius(self.h, int(r))

    def set_descriptor_distance(self, d):
        lib().orc_pf_set_descriptor_distance(self.h, C.c_float(d))

    def compute(self):
        fi, mi, d = _corr_out(self.cap)
        n = lib().orc_pf_compute(self.h, self.cap, _p(fi), _p(mi), _p(d))
        assert n <= self.cap
        return fi[:n].copy(), mi[:n].copy(), d[:n].copy()

    def state(self):
        s = np.zeros(6, np.float32)
        lib().orc_pf_state(self.h, _p(s))
        return dict(radius=int(s[0]), descriptor_distance=float(s[1]), iteration=int(s[2]),
                    converged=bool(s[3]), searches=int(s[4]), n_projected=int(s[5]))

    def candidates(self):
        out = np.zeros((self.ncap, 5), np.int32)
        n = lib().orc_pf_candidates(self.h, self.ncap, _p(out))
        return out[:n].copy()

    def lattice(self):
        out = np.zeros(self.cap, np.int32)
        n = lib().orc_pf_lattice(self.h, self.cap, _p(out))
        return out[:n].copy()


FACTORS = {"stereo": 0, "depth": 1, "mono": 2}
ROBUST = {"none": 0, "saturated": 1, "clamp": 2}


def linearize_cfg(kind, K, cols, rows, baseline=(0, 0, 0), mean_disparity=0.0,
                  robustifier="saturated", chi_threshold=25.0):
    return np.concatenate([[FACTORS[kind]], np.asarray(K, np.float64).reshape(9), [cols, rows],
                           np.asarray(baseline, np.float64), [mean_disparity, ROBUST[robustifier],
                                                              chi_threshold]]).astype(np.float64)


def linearize(lcfg, pose12, moving_xyz, fixed_meas, corr_fixed, corr_moving, info_diag, fp32=False, prior=None,
              want_status=False):
    """prior = (predicted pose12, information 6x6): the motion-model slice's pose-prior factor, summed into H, b.
    want_status: the returned dict carries "status" (per correspondence: 0 inlier, 1 kernelized, 2 suppressed)."""
    pose12 = np.ascontiguousarray(pose12, np.float64).reshape(12)
    moving_xyz = np.ascontiguousarray(moving_xyz, np.float64).reshape(-1, 3)
    fixed_meas = np.ascontiguousarray(fixed_meas, np.float64)
    cf = np.ascontiguousarray(corr_fixed, np.int32)
    cm = np.ascontiguousarray(corr_moving, np.int32)
    info = np.ascontiguousarray(info_diag, np.float64).reshape(-1, 3)
    assert len(info) == len(fixed_meas)
    H = np.zeros(36, np.float64)
    b = np.zeros(6, np.float64)
    st = np.zeros(5, np.float64)
    if fp32:
        assert prior is None and not want_status
        lib().orc_linearize_f32(_p(lcfg), _p(pose12), len(moving_xyz), _p(moving_xyz), len(fixed_meas), _p(fixed_meas),
                                fixed_meas.shape[1], len(cf), _p(cf), _p(cm), _p(info), _p(H), _p(b), _p(st))
        status = None
    else:
        pr = None if prior is None else np.concatenate([np.asarray(prior[0], np.float64).reshape(12),
                                                        np.asarray(prior[1], np.float64).reshape(36)])
        status = np.zeros(max(len(cf), 1), np.uint8) if want_status else None
        lib().orc_linearize_ex_f64(_p(lcfg), _p(pose12), len(moving_xyz), _p(moving_xyz), len(fixed_meas), _p(fixed_meas),
                                   fixed_meas.shape[1], len(cf), _p(cf), _p(cm), _p(info), _p(pr), _p(status), _p(H), _p(b), _p(st))
    out = dict(chi=st[0], inliers=int(st[1]), outliers=int(st[2]), suppressed=int(st[3]), prior_chi=st[4])
    if want_status:
        out["status"] = status[:len(cf)].copy()
    return H.reshape(6, 6), b, out


def gn_step(H, b, damping, pose12):
    H = np.ascontiguousarray(H, np.float64).reshape(36)
    b = np.ascontiguousarray(b, np.float64).reshape(6)
    pose = np.ascontiguousarray(pose12, np.float64).reshape(12).copy()
    dx = np.zeros(6, np.float64)
    rc = lib().orc_gn_step_f64(_p(H), _p(b), C.c_double(damping), _p(pose), _p(dx))
    return rc, pose, dx


def t2tnq(pose12):
    pose12 = np.ascontiguousarray(pose12, np.float64).reshape(12)
    v = np.zeros(6, np.float64)
    lib().orc_t2tnq_f64(_p(pose12), _p(v))
    return v


def pose_inverse(pose12):
    pose12 = np.ascontiguousarray(pose12, np.float64).reshape(12)
    o = np.zeros(12, np.float64)
    lib().orc_pose_inverse_f64(_p(pose12), _p(o))
    return o


def pose_mul(a, b):
    a = np.ascontiguousarray(a, np.float64).reshape(12)
    b = np.ascontiguousarray(b, np.float64).reshape(12)
    o = np.zeros(12, np.float64)
    lib().orc_pose_mul_f64(_p(a), _p(b), _p(o))
    return o


# ---- aligner loop (srrg2_slam_interfaces MultiAligner3DQR, external) --------------------------
# ORACLE restatement of one frame-to-map alignment, built from the oracle primitives above: per iteration
# finder.compute (state machine + window search + filter) -> setupFactor (per-correspondence information,
# aligner_slice_processor_projective.cpp:41-57; stereo mean disparity :78-88) -> linearise (fp64) -> GN step.
ALIGNER_STATUS = {"Fail": 0, "Success": 1, "NotEnoughCorrespondences": 2, "NotEnoughInliers": 3}


def mean_disparity_f32(fixed):
    """fp32 accumulation in cloud order (aligner_slice_processor_projective.cpp:80-85)"""
    acc = np.float32(0)
    for p in np.asarray(fixed, np.float32):
        acc = np.float32(acc + np.float32(p[0] - p[2]))
    return float(np.float32(acc / np.float32(len(fixed)))) if len(fixed) else 0.0


class BruteforceFinder:
    """CorrespondenceFinderDescriptorBasedBruteforce as an aligner's finder: the match does not depend on the estimate
    and is recomputed only when a cloud changed (change flags, bruteforce_impl.cpp:13-15,233-235)."""

    def __init__(self, max_dist=50.0, ratio=0.9):
        self.max_dist, self.ratio = max_dist, ratio
        self._f = self._m = self._result = None

    def set_fixed(self, coords, desc):
        self._f, self._result = np.ascontiguousarray(desc, np.uint8), None

    def set_moving(self, xyz, desc):
        self._m, self._result = np.ascontiguousarray(desc, np.uint8), None

    def set_estimate(self, pose12):
        pass

    def compute(self):
        if self._result is None:
            self._result = match_bruteforce(self._f, self._m, self.max_dist, self.ratio)
        return self._result


def constant_velocity_prediction(trajectory_chunk):
    """[upstream] MotionModelConstantVelocity3D over the aligner's "trajectory_chunk" slice (robot poses in the local map,
    oldest first): the last relative motion is applied once more.  Returns the predicted moving_in_fixed (local map in
    sensor) = (P_n * P_{n-1}^-1 * P_n)^-1; fewer than two poses: no motion (identity for an empty chunk)."""
    if len(trajectory_chunk) == 0:
        return np.eye(3, 4).reshape(12)
    last = np.asarray(trajectory_chunk[-1], np.float64).reshape(12)
    if len(trajectory_chunk) == 1:
        return pose_inverse(last)
    prev = np.asarray(trajectory_chunk[-2], np.float64).reshape(12)
    motion = pose_mul(pose_inverse(prev), last)
    return pose_inverse(pose_mul(last, motion))


def align(pf, kind, K, rows, cols, fixed, moving_xyz, diag, init_pose=None, n_opt=None, baseline=(0, 0, 0),
          inverse_depth_weighting=False, robustifier="saturated", chi_threshold=25.0, max_iterations=100,
          damping=0.0, min_num_inliers=6, min_num_correspondences=0, prior=None, enable_inlier_only_runs=False,
          keep_only_inlier_correspondences=False):
    """pf: a finder (O.ProjectiveFinder or O.BruteforceFinder) with fixed / moving already set.
    prior: (predicted moving_in_fixed pose12, 6x6 information) of the motion-model slice, or None.
    enable_inlier_only_runs / keep_only_inlier_correspondences (configurations/icl.conf:55-58) [upstream, restated as]:
    after a successful main loop the correspondences whose factor was an inlier in the last linearisation are frozen and,
    if at least min_num_inliers remain, `max_iterations` further GN iterations run on them alone (the finder is not
    consulted again); keep_only leaves exactly those correspondences in the result.
    Returns dict(status, pose, stats, corr[, inlier_run_stats])."""
    fixed = np.ascontiguousarray(fixed, np.float32)
    moving_xyz = np.ascontiguousarray(moving_xyz, np.float32).reshape(-1, 3)
    pose = np.eye(3, 4, dtype=np.float64).reshape(12) if init_pose is None else \
        np.asarray(init_pose, np.float32).astype(np.float64).reshape(12)
    md = mean_disparity_f32(fixed) if (kind == "stereo" and inverse_depth_weighting) else 0.0
    lcfg = linearize_cfg(kind, np.asarray(K, np.float32), cols, rows, np.asarray(baseline, np.float32), md,
                         robustifier, chi_threshold)
    diag = np.asarray(diag, np.float32)
    d3 = np.zeros(3, np.float32)
    d3[:len(diag)] = diag

    def information(fi, mi):
        info = np.zeros((len(fixed), 3), np.float64)
        for f, m in zip(fi, mi):
            d = d3.copy()
            n = 0 if n_opt is None else int(n_opt[m])
            if n > 2:
                d = (d * np.float32(1 + np.log(float(n)))).astype(np.float32)
            info[f] = d
        return info

    stats, corr, enough = [], (np.zeros(0, np.int32),) * 2 + (np.zeros(0, np.float32),), True
    last = None
    for it in range(max_iterations):
        pf.set_estimate(pose.astype(np.float32))
        corr = pf.compute()
        fi, mi, _ = corr
        if len(fi) < max(min_num_correspondences, 1):
            stats.append((len(fi), 0, 0, 0.0))
            enough = False
            break
        H, b, st = linearize(lcfg, pose, moving_xyz, fixed, fi, mi, information(fi, mi), prior=prior, want_status=True)
        rc, new_pose, _ = gn_step(H, b, damping, pose)
        last = st
        stats.append((len(fi), st["inliers"], st["outliers"], st["chi"]))
        if rc != 0:
            break
        pose = new_pose
    if not enough:
        status = ALIGNER_STATUS["NotEnoughCorrespondences"]
    else:
        status = ALIGNER_STATUS["Success"] if last and last["inliers"] >= min_num_inliers else ALIGNER_STATUS["NotEnoughInliers"]
    out = {"status": status, "pose": pose, "stats": np.asarray(stats, np.float64).reshape(-1, 4), "corr": corr}
    if status == ALIGNER_STATUS["Success"] and (enable_inlier_only_runs or keep_only_inlier_correspondences):
        keep = last["status"] == 0
        inl = tuple(np.ascontiguousarray(a[keep]) for a in corr)
        if enable_inlier_only_runs and int(keep.sum()) >= min_num_inliers:
            fi, mi, _ = inl
            info, rstats = information(fi, mi), []
            for it in range(max_iterations):
                H, b, st = linearize(lcfg, pose, moving_xyz, fixed, fi, mi, info, prior=prior, want_status=True)
                rc, new_pose, _ = gn_step(H, b, damping, pose)
                rstats.append((len(fi), st["inliers"], st["outliers"], st["chi"]))
                if rc != 0:
                    break
                pose = new_pose
                last = st
            out["pose"], out["inlier_run_stats"] = pose, np.asarray(rstats, np.float64).reshape(-1, 4)
            if keep_only_inlier_correspondences:  # inliers of the LAST linearisation of the inlier-only run
                k2 = last["status"] == 0
                inl = tuple(np.ascontiguousarray(a[k2]) for a in inl)
        if keep_only_inlier_correspondences:
            out["corr"] = inl
    return out

"""ctypes harness over libpslam_plugin.so (include/pslam_plugin.h): the host-side C++ mirror of the srrg2_proslam
plugin classes.  Used by tests/ only -- it lets the parity tests read like the reference's gtest files: load a
`.conf`, fetch a module by name, set PARAMs, hand it images / clouds, call compute()."""
import ctypes as C
import pathlib

import numpy as np

LIB_PATH = pathlib.Path(__file__).resolve().parent / "libpslam_plugin.so"
_lib = None


class PluginError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C srrg2_proslam_b200/host` "
                               "(or __graft_entry__.build())")
        _lib = C.CDLL(str(LIB_PATH))
        for name in ("psp_last_error", "psp_module_class_name", "psp_module_name", "psp_module_get_string"):
            getattr(_lib, name).restype = C.c_char_p
        for name in ("psp_manager_create", "psp_manager_at", "psp_manager_get_by_name", "psp_manager_create_module",
                     "psp_module_get_link"):
            getattr(_lib, name).restype = C.c_void_p
    return _lib


def _chk(rc):
    if rc < 0:
        raise PluginError(lib().psp_last_error().decode())
    return rc


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Module:
    """a Configurable owned by a Manager"""

    def __init__(self, handle, manager):
        self.h = C.c_void_p(handle)
        self.manager = manager  # keeps the owner alive

    @property
    def class_name(self):
        return lib().psp_module_class_name(self.h).decode()

    @property
    def name(self):
        return lib().psp_module_name(self.h).decode()

    @property
    def is_generic(self):
        return bool(lib().psp_module_is_generic(self.h))

    def has(self, param):
        return bool(lib().psp_module_has_param(self.h, param.encode()))

    def set(self, param, value):
        L = lib()
        if isinstance(value, Module) or value is None:
            _chk(L.psp_module_set_link(self.h, param.encode(), value.h if value is not None else None))
        elif isinstance(value, str):
            _chk(L.psp_module_set_string(self.h, param.encode(), value.encode()))
        elif isinstance(value, (list, tuple, np.ndarray)):
            v = np.asarray(value, np.float64)
            _chk(L.psp_module_set_numbers(self.h, param.encode(), len(v), _p(v)))
        else:
            _chk(L.psp_module_set_number(self.h, param.encode(), C.c_double(float(value))))
        return self

    def get(self, param):
        v = C.c_double()
        _chk(lib().psp_module_get_number(self.h, param.encode(), C.byref(v)))
        return v.value

    def get_string(self, param):
        return lib().psp_module_get_string(self.h, param.encode()).decode()

    def get_numbers(self, param):
        buf = np.zeros(64, np.float64)
        n = _chk(lib().psp_module_get_numbers(self.h, param.encode(), 64, _p(buf)))
        return buf[:n].copy()

    def link(self, param):
        h = lib().psp_module_get_link(self.h, param.encode())
        return Module(h, self.manager) if h else None

    # ---- IntensityFeatureExtractorBinned ------------------------------------------------------------------
    def extract(self, img, mask=None, cap=8192):
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols = img.shape
        xy = np.zeros((cap, 2), np.float32)
        inten = np.zeros(cap, np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        n = _chk(lib().psp_extractor_compute(self.h, _p(img), rows, cols, cols, _p(m), cap, _p(xy), _p(inten), _p(desc)))
        return {"xy": xy[:n].copy(), "intensity": inten[:n].copy(), "desc": desc[:n].copy()}

    def set_projections(self, coords, radius):
        coords = np.ascontiguousarray(coords, np.float32)
        _chk(lib().psp_extractor_set_projections(self.h, len(coords), coords.shape[1], _p(coords), int(radius)))

    def paint_tracking_mask(self, rows, cols, coords, radius):
        coords = np.ascontiguousarray(coords, np.float32)
        mask = np.zeros((rows, cols), np.uint8)
        _chk(lib().psp_extractor_paint_tracking_mask(self.h, rows, cols, len(coords), coords.shape[1], _p(coords), int(radius), _p(mask)))
        return mask

    def number_of_tracking_keypoints(self):
        return _chk(lib().psp_extractor_number_of_tracking_keypoints(self.h))

    # ---- RawDataPreprocessorStereoProjective ----------------------------------------------------------------
    def stereo_adaptor(self, left, right, cap=8192):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        rows, cols = left.shape
        uvuv = np.zeros((cap, 4), np.float32)
        inten = np.zeros(cap, np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        status = C.c_int(-1)
        n = _chk(lib().psp_stereo_adaptor_compute(self.h, _p(left), _p(right), rows, cols, cols, cap, _p(uvuv), _p(inten),
                                                  _p(desc), C.byref(status)))
        return {"uvuv": uvuv[:n].copy(), "intensity": inten[:n].copy(), "desc": desc[:n].copy(), "status": status.value}

    # ---- RawDataPreprocessorMonocularDepth ------------------------------------------------------------------
    def mono_depth_adaptor(self, img, depth, cap=8192):
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols = img.shape
        if depth.dtype == np.uint16:
            dtype = 0
        elif depth.dtype == np.float32:
            dtype = 1
        else:
            dtype = 7
        depth = np.ascontiguousarray(depth)
        uvz = np.zeros((cap, 3), np.float32)
        inten = np.zeros(cap, np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        status = C.c_int(-1)
        n = _chk(lib().psp_mono_depth_adaptor_compute(self.h, _p(img), rows, cols, cols, _p(depth), dtype, depth.shape[0],
                                                      depth.shape[1], depth.shape[1], cap, _p(uvz), _p(inten), _p(desc),
                                                      C.byref(status)))
        return {"uvz": uvz[:n].copy(), "intensity": inten[:n].copy(), "desc": desc[:n].copy(), "status": status.value}

    # ---- CorrespondenceFinder* ------------------------------------------------------------------------------
    def set_fixed(self, coords, desc):
        coords = np.ascontiguousarray(coords, np.float32)
        desc = np.ascontiguousarray(desc, np.uint8)
        _chk(lib().psp_finder_set_fixed(self.h, len(coords), coords.shape[1] if coords.ndim == 2 else 2, _p(coords), _p(desc)))

    def set_moving(self, coords, desc):
        coords = np.ascontiguousarray(coords, np.float32)
        desc = np.ascontiguousarray(desc, np.uint8)
        _chk(lib().psp_finder_set_moving(self.h, len(coords), coords.shape[1] if coords.ndim == 2 else 3, _p(coords), _p(desc)))

    def set_local_map_in_sensor(self, pose12):
        pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
        _chk(lib().psp_finder_set_local_map_in_sensor(self.h, _p(pose12)))

    def compute(self, cap=16384):
        f = np.zeros(cap, np.int32)
        m = np.zeros(cap, np.int32)
        d = np.zeros(cap, np.float32)
        n = _chk(lib().psp_finder_compute(self.h, cap, _p(f), _p(m), _p(d)))
        return f[:n].copy(), m[:n].copy(), d[:n].copy()

    def projective_state(self):
        r, it, cv, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        dd = C.c_float()
        _chk(lib().psp_projective_finder_state(self.h, C.byref(r), C.byref(dd), C.byref(it), C.byref(cv), C.byref(ns)))
        return {"radius": r.value, "descriptor_distance": dd.value, "iteration": it.value, "converged": bool(cv.value),
                "searches": ns.value}

    def set_projective_state(self, radius, descriptor_distance):
        _chk(lib().psp_projective_finder_set_state(self.h, int(radius), C.c_float(descriptor_distance)))

    def set_camera_matrix(self, K):
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        _chk(lib().psp_projector_set_camera_matrix(self.h, _p(K)))

    # ---- MultiAligner3DQR -----------------------------------------------------------------------------------
    def aligner_set_fixed(self, coords, desc):
        coords = np.ascontiguousarray(coords, np.float32)
        desc = np.ascontiguousarray(desc, np.uint8)
        _chk(lib().psp_aligner_set_fixed(self.h, len(coords), coords.shape[1], _p(coords), _p(desc)))

    def aligner_set_moving(self, xyz, desc, n_opt=None):
        xyz = np.ascontiguousarray(xyz, np.float32)
        desc = np.ascontiguousarray(desc, np.uint8)
        no = None if n_opt is None else np.ascontiguousarray(n_opt, np.int32)
        _chk(lib().psp_aligner_set_moving(self.h, len(xyz), _p(xyz), _p(desc), _p(no)))

    # ---- SceneClipperProjective3D ----------------------------------------------------------------------------------
    def clipper_set_full_scene(self, xyz, desc, intensity=None):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(len(xyz), 32)
        inten = None if intensity is None else np.ascontiguousarray(intensity, np.float32)
        self._scene_size = len(xyz)
        _chk(lib().psp_clipper_set_full_scene(self.h, len(xyz), _p(xyz), _p(inten), _p(desc)))

    def clipper_set_robot_in_local_map(self, pose12):
        pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
        _chk(lib().psp_clipper_set_robot_in_local_map(self.h, _p(pose12)))

    def clipper_set_sensor_in_robot(self, pose12):
        pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
        _chk(lib().psp_clipper_set_sensor_in_robot(self.h, _p(pose12)))

    def clipper_compute(self):
        cap = max(1, getattr(self, "_scene_size", 0))
        xyz, uvz = np.zeros((cap, 3), np.float32), np.zeros((cap, 3), np.float32)
        idx, desc = np.zeros(cap, np.int32), np.zeros((cap, 32), np.uint8)
        status = C.c_int(-1)
        n = _chk(lib().psp_clipper_compute(self.h, cap, _p(xyz), _p(uvz), _p(idx), _p(desc), C.byref(status)))
        return {"xyz": xyz[:n].copy(), "uvz": uvz[:n].copy(), "index": idx[:n].copy(), "desc": desc[:n].copy(),
                "status": status.value}

    # ---- point EKFs + LandmarkEstimatorEKF ---------------------------------------------------------------------------
    def filter_set_camera(self, K, baseline=(0.0, 0.0)):
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        _chk(lib().psp_point_filter_set_camera(self.h, _p(K), C.c_double(baseline[0]), C.c_double(baseline[1])))

    def estimator_set_transforms(self, measurement_in_world, measurement_in_scene):
        a = np.ascontiguousarray(measurement_in_world, np.float32).reshape(12)
        b = np.ascontiguousarray(measurement_in_scene, np.float32).reshape(12)
        _chk(lib().psp_landmark_estimator_set_transforms(self.h, _p(a), _p(b)))

    def estimator_compute_batch(self, state_world, covariance, meas):
        st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
        n = len(st)
        cv = np.ascontiguousarray(covariance, np.float32).reshape(n, 9).copy()
        ms = np.ascontiguousarray(meas, np.float32).reshape(n, -1)
        loc, inl = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
        k = _chk(lib().psp_landmark_estimator_compute_batch(self.h, n, _p(st), _p(cv), _p(ms), _p(loc), _p(inl)))
        return st, cv.reshape(n, 3, 3), loc, inl.astype(bool), k

    def estimator_weighted_mean_batch(self, state_world, n_opt, landmark_in_sensor):
        st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
        n = len(st)
        no = np.ascontiguousarray(n_opt, np.int32).reshape(n)
        ls = np.ascontiguousarray(landmark_in_sensor, np.float32).reshape(n, 3)
        loc, inl = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
        k = _chk(lib().psp_landmark_estimator_weighted_mean_batch(self.h, n, _p(st), _p(no), _p(ls), _p(loc), _p(inl)))
        return st, loc, inl.astype(bool), k

    def merger_select_updates(self, measurements, corr_moving, corr_response):
        """MergerProjective_::compute update pass: selected[n_corr] (the module keeps the blocked bins)"""
        m = np.ascontiguousarray(measurements, np.float32)
        mv = np.ascontiguousarray(corr_moving, np.int32).reshape(-1)
        rs = np.ascontiguousarray(corr_response, np.float32).reshape(len(mv))
        sel = np.zeros(max(len(mv), 1), np.uint8)
        _chk(lib().psp_merger_select_updates(self.h, _p(m), m.shape[1], len(m), _p(mv), _p(rs), len(mv), _p(sel)))
        return sel[:len(mv)].astype(bool)

    def merger_plan(self, measurements, corr_moving, corr_response):
        """both passes in one device round trip: (selected[n_corr] bool, addition winners)"""
        m = np.ascontiguousarray(measurements, np.float32)
        mv = np.ascontiguousarray(corr_moving, np.int32).reshape(-1)
        rs = np.ascontiguousarray(corr_response, np.float32).reshape(len(mv))
        sel, win, nw = np.zeros(max(len(mv), 1), np.uint8), np.zeros(max(len(m), 1), np.int32), C.c_int(0)
        _chk(lib().psp_merger_plan(self.h, _p(m), m.shape[1], len(m), _p(mv), _p(rs), len(mv), _p(sel), _p(win), C.byref(nw)))
        return sel[:len(mv)].astype(bool), win[:nw.value].copy()

    def merger_wants_additions(self, merged, n_meas, n_corr):
        return bool(_chk(lib().psp_merger_wants_additions(self.h, int(merged), int(n_meas), int(n_corr))))

    def merger_select_additions(self, measurements):
        m = np.ascontiguousarray(measurements, np.float32)
        win = np.zeros(max(len(m), 1), np.int32)
        k = _chk(lib().psp_merger_select_additions(self.h, _p(m), m.shape[1], len(m), _p(win)))
        return win[:k].copy()

    def smoother_set_camera_matrix(self, K):
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        _chk(lib().psp_landmark_smoother_set_camera_matrix(self.h, _p(K)))

    def smoother_compute_batch(self, frames_sensor_in_world, offsets, hist_frame, hist_uv, hist_point_in_camera, state_world, n_opt):
        st = np.ascontiguousarray(state_world, np.float32).reshape(-1, 3).copy()
        n = len(st)
        no = np.ascontiguousarray(n_opt, np.int32).reshape(n).copy()
        fr = np.ascontiguousarray(frames_sensor_in_world, np.float32).reshape(-1, 12)
        off = np.ascontiguousarray(offsets, np.int32).reshape(n + 1)
        hf = np.ascontiguousarray(hist_frame, np.int32)
        uv = np.ascontiguousarray(hist_uv, np.float32).reshape(len(hf), 2)
        pic = np.ascontiguousarray(hist_point_in_camera, np.float32).reshape(len(hf), 3)
        loc, inl = np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
        _chk(lib().psp_landmark_smoother_compute_batch(self.h, n, _p(st), _p(no), len(fr), _p(fr), _p(off), _p(hf), _p(uv), _p(pic),
                                                       _p(loc), _p(inl)))
        return st, no, loc, inl.astype(bool)

    def aligner_set_moving_in_fixed(self, pose12):
        pose12 = np.ascontiguousarray(pose12, np.float32).reshape(12)
        _chk(lib().psp_aligner_set_moving_in_fixed(self.h, _p(pose12)))

    def aligner_set_left_camera_in_right(self, t3):
        t3 = np.ascontiguousarray(t3, np.float32).reshape(3)
        _chk(lib().psp_aligner_set_left_camera_in_right(self.h, _p(t3)))

    def aligner_set_slice_processor(self, index, slice_module):
        _chk(lib().psp_aligner_set_slice_processor(self.h, int(index), slice_module.h))

    def aligner_num_slice_processors(self):
        return _chk(lib().psp_aligner_num_slice_processors(self.h))

    def aligner_set_trajectory_chunk(self, poses):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
        _chk(lib().psp_aligner_set_trajectory_chunk(self.h, len(poses), _p(poses)))

    def aligner_set_prior_information(self, information):
        information = np.ascontiguousarray(information, np.float64).reshape(36)
        _chk(lib().psp_aligner_set_prior_information(self.h, _p(information)))

    def aligner_run(self):
        """MultiAligner_::compute() alone (what a latency measurement should time); the per-iteration statistics and the
        correspondences are read afterwards with aligner_results()"""
        pose = np.zeros(12, np.float64)
        it, nc, ni = C.c_int(), C.c_int(), C.c_int()
        chi = C.c_double()
        status = _chk(lib().psp_aligner_compute(self.h, _p(pose), C.byref(it), C.byref(nc), C.byref(ni), C.byref(chi)))
        return {"status": status, "pose": pose, "iterations": it.value, "num_correspondences": nc.value,
                "num_inliers": ni.value, "chi": chi.value}

    def aligner_compute(self):
        return self.aligner_results(self.aligner_run())

    def aligner_results(self, run):
        status, pose = run["status"], run["pose"]
        it, nc, ni, chi = (C.c_int(run["iterations"]), C.c_int(run["num_correspondences"]), C.c_int(run["num_inliers"]),
                           C.c_double(run["chi"]))
        rows = np.zeros((max(it.value, 1), 4), np.float64)
        _chk(lib().psp_aligner_iteration_stats(self.h, len(rows), _p(rows)))
        f = np.zeros(16384, np.int32)
        m = np.zeros(16384, np.int32)
        d = np.zeros(16384, np.float32)
        n = _chk(lib().psp_aligner_correspondences(self.h, 16384, _p(f), _p(m), _p(d)))
        rrows = np.zeros((4096, 4), np.float64)
        nr = _chk(lib().psp_aligner_inlier_run_stats(self.h, len(rrows), _p(rrows)))
        return {"status": status, "pose": pose, "iterations": it.value, "num_correspondences": nc.value,
                "num_inliers": ni.value, "chi": chi.value, "stats": rows[:it.value], "inlier_run_stats": rrows[:nr].copy(),
                "corr": (f[:n].copy(), m[:n].copy(), d[:n].copy())}


class Manager:
    """srrg2_core::ConfigurableManager: read a .conf, look modules up by name"""

    def __init__(self, conf_path=None, text=None):
        self.h = C.c_void_p(lib().psp_manager_create())
        if conf_path is not None:
            self.read(conf_path)
        if text is not None:
            _chk(lib().psp_manager_read_string(self.h, text.encode()))

    def close(self):
        if self.h:
            lib().psp_manager_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def read(self, path):
        return _chk(lib().psp_manager_read(self.h, str(path).encode()))

    def write(self, path, names=()):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        _chk(lib().psp_manager_write(self.h, str(path).encode(), len(names), arr))

    def __len__(self):
        return lib().psp_manager_count(self.h)

    def modules(self):
        return [Module(lib().psp_manager_at(self.h, i), self) for i in range(len(self))]

    def get(self, name):
        h = lib().psp_manager_get_by_name(self.h, name.encode())
        if not h:
            raise KeyError(name)
        return Module(h, self)

    def create(self, class_name, name=""):
        h = lib().psp_manager_create_module(self.h, class_name.encode(), name.encode())
        return Module(h, self)


def is_registered(class_name):
    return bool(lib().psp_class_is_registered(class_name.encode()))

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native srrg2_proslam frontend.

Metric (BASELINE.json): frontend stereo frames/s on the frame-sharded config -- KITTI-shaped
(1241x376) synthetic stereo pairs, 4k target features/frame, kitti.conf matcher parameters.  One step =
one pass of detect -> select -> describe (L and R) -> epipolar stereo match over the rank's whole batch.

  value : images already resident in HBM, timed on the device (CUDA events on the context's stream)
  e2e   : the same batch through the host-pointer C-ABI call (pinned HOST images in, packed stereo
          measurement clouds out), host<->device copies inside the timed region (wall clock around
          synchronised calls)
  roofline     : dominant kernel, live per-kernel CUDA-event time (pslam_profile_*), algorithmic bytes
                 per image from SURVEY.md section 8(d) / DESIGN.md
  cpu_baseline : the CPU oracle (port of the reference path) on a bounded sample, all host threads
  --impl reference : the same oracle as the reference arm (the reference itself cannot be built here,
                 see DESIGN.md "Oracle"), rank 0 only.

N > 1: one process per GPU (torchrun), frames sharded by rank (weak scaling: every rank owns --pairs
pairs), no data-path collective; NCCL only for the barrier and the max-over-ranks of the timing.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ROWS, COLS = 376, 1241
THRESHOLD, TARGET = 15, 4000                 # kitti.conf detector threshold; config 4: 4k features / frame
MATCH = dict(max_dist=100.0, ratio=0.5, max_disp=100, thickness=0)   # kitti.conf epipolar finder (:484-501)
GEN_CHUNK = 250


_REAL_STDOUT = None


def quiet_stdout():
    """the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there) get stderr instead"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def clocks_sampler(stop, out, gpu_index):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    t.join(timeout=2)


def summarise_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for nme, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(nme)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_lib():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    O.build()
    return O


def run_reference(args, rank, world):
    """reference arm: the CPU implementation of the path on the host cores (oracle port), rank 0 only"""
    if rank != 0:
        return
    import torch
    from srrg2_proslam_b200 import synth
    O = oracle_lib()
    cores = os.cpu_count() or 1
    sample = args.ref_pairs
    imgs = synth.stereo_pairs(sample, ROWS, COLS, seed=args.seed, device="cpu").numpy()
    cfg = O.extract_cfg(THRESHOLD, 1, TARGET)
    for _ in range(args.warmup):
        O.stereo_frontend_batch(imgs, cfg, threads=cores, **MATCH)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        counts, _ = O.stereo_frontend_batch(imgs, cfg, threads=cores, **MATCH)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": "frontend_stereo_frames_per_s", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, sample),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} stereo pairs of the workload per step (CPU generator, same seed)"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mean_stereo_points_per_frame": float(np.mean(counts))}
    emit(line)


def workload_config(args, pairs):
    return {"workload": "frame-sharded stereo frontend: KITTI-shaped synthetic stereo pairs (BASELINE config 4)",
            "pairs_per_gpu": pairs, "image": f"{COLS}x{ROWS} u8", "target_features_per_frame": TARGET,
            "detector": f"FAST-9/16 thr {THRESHOLD} + NMS, 3x3 bins, ORB-256",
            "matcher": "epipolar max_dist 100 ratio 0.5 disparity<=100 thickness 0 (kitti.conf)",
            "stages": "detect+select+describe (L,R) -> epipolar stereo match -> stereo measurement cloud",
            "images": "seeded block-texture generator (srrg2_proslam_b200/synth.py), not photographs: ~17.7 % of the pixels pass the FAST "
                      "compass pre-test (natural KITTI frames: a few per cent), ~3 244 of the 4 000 targeted keypoints described per image",
            "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"frames sharded over {args.gpus} GPU(s)",
            "chunk_lanes": getattr(args, "lanes", 2)}


def tracking_lines(ctx, capi, stream, dev):
    """(1) per-frame latency of the conf-driven aligner (kitti.conf: projective finder + stereo factor + GN, 100
    iterations) on the KITTI 00 -> 01 pair of tests/golden through the plugin mirror; (2) H,b linearisation throughput
    on a batched synthetic (2^22 stereo correspondences) against the FP64 pipe."""
    import cv2
    import torch
    from srrg2_proslam_b200 import plugin as P
    out = {}
    G = ROOT / "tests" / "golden"
    K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)
    ld = lambda n: cv2.imread(str(G / n), cv2.IMREAD_UNCHANGED)
    small = capi.Context(device=ctx.device, max_images=2, max_rows=376, max_cols=1241,
                         max_features=2048, max_raw_per_bin=8192)
    try:
        e, mcfg = capi.extract_cfg(15, 1, 500), capi.match_cfg(50, 0.8, 100, 0)
        meas = [small.stereo_adaptor(ld(f"kitti_city_image_left_{i}.png"), ld(f"kitti_city_image_right_{i}.png"), e, mcfg) for i in (0, 1)]
        xyz, _, _ = small.triangulate(meas[0]["uvuv"], K, float(np.float32(718.856) * np.float32(0.537166)), 0.0)
        _, bf_moving, bf_response = small.match_bruteforce(meas[0]["desc"], meas[1]["desc"], capi.match_cfg(75, 0.8))
        # per-frame latency of the stereo adaptor (host images in, measurement cloud out: detect + describe L, R, match)
        L1, R1 = ld("kitti_city_image_left_1.png"), ld("kitti_city_image_right_1.png")
        e1k = capi.extract_cfg(15, 1, 1000)
        ta = []
        for rep in range(12):
            t0 = time.perf_counter()
            r1 = small.stereo_adaptor(L1, R1, e1k, capi.match_cfg(**MATCH))
            ta.append(time.perf_counter() - t0)
        out["adaptor"] = {"metric": "stereo_adaptor_ms_per_frame", "value": 1e3 * float(np.median(ta[2:])), "unit": "ms",
                          "stereo_points": int(len(r1["uvuv"])),
                          "config": "kitti.conf adaptor_stereo_projective on KITTI frame 01 of tests/golden (1241x376 pair, host pointers)"}
    finally:
        small.close()
    P.lib().psp_set_device(ctx.device)
    m = P.Manager(G / "configurations" / "kitti_hotpath.conf")
    al = m.get("aligner")
    sl = [x for x in m.modules() if x.class_name == "AlignerSliceProcessorProjectiveStereo"][0]
    pr = sl.link("projector")
    pr.set_camera_matrix(K)
    pr.set("canvas_rows", 376).set("canvas_cols", 1241)
    al.aligner_set_left_camera_in_right([-0.537166, 0, 0])
    times = []
    for rep in range(12):
        al.aligner_set_fixed(meas[1]["uvuv"], meas[1]["desc"])
        al.aligner_set_moving(xyz, meas[0]["desc"])
        al.aligner_set_moving_in_fixed(np.eye(3, 4, dtype=np.float32))
        t0 = time.perf_counter()
        r = al.aligner_run()  # MultiAligner3DQR::compute(): uploads, searches, solver iterations, download -- host pointers in and out
        times.append(time.perf_counter() - t0)
    out["aligner"] = {"metric": "aligner_ms_per_frame", "value": 1e3 * float(np.median(times[2:])), "unit": "ms",
                      "iterations": int(r["iterations"]), "correspondences": int(r["num_correspondences"]),
                      "inliers": int(r["num_inliers"]), "status": int(r["status"]),
                      "config": "kitti.conf aligner (MultiAligner3DQR -> AlignerSliceProcessorProjectiveStereo + "
                                "AlignerSliceMotionModel3D -> CorrespondenceFinderProjectiveCircle4D3D), KITTI 00 -> 01 of "
                                "tests/golden, identity guess, empty trajectory chunk"}
    # merger pass of the same frame pair (kitti.conf merger_ekf): binned update selection + binned additions, host pointers
    try:
        mg = m.get("merger_ekf")
        pm = mg.link("projector")
        pm.set_camera_matrix(K)
        pm.set("canvas_rows", 376).set("canvas_cols", 1241)
        mv, rs = bf_moving, bf_response  # frame 00 -> 01 correspondences (exhaustive matcher, 75 / 0.8)
        tm = []
        for rep in range(12):
            t0 = time.perf_counter()
            sel, win = mg.merger_plan(meas[1]["uvuv"], mv, rs)
            tm.append(time.perf_counter() - t0)
        O = oracle_lib()  # the same pass through the CPU restatement: checker + context number
        t0 = time.perf_counter()
        for rep in range(20):
            osel, occ = O.merger_select_updates(meas[1]["uvuv"], mv, rs, 376, 1241, 20, 60, 100.0, True, "stereo")
            owin = O.merger_select_additions(meas[1]["uvuv"], occ, 376, 1241, 20, 60, True, "stereo")
        t_cpu = (time.perf_counter() - t0) / 20
        out["merger"] = {"metric": "merger_binning_ms_per_frame", "value": 1e3 * float(np.median(tm[2:])), "unit": "ms",
                         "cpu_oracle_ms": 1e3 * t_cpu, "measurements": int(len(meas[1]["uvuv"])), "correspondences": int(len(mv)),
                         "updates": int(sel.sum()), "additions": int(len(win)),
                         "parity": bool(np.array_equal(sel, osel) and np.array_equal(win, owin)),
                         "config": "kitti.conf merger_ekf (MergerRigidStereoProjectiveEKF, 20 x 60 bins), KITTI 00 -> 01: "
                                   "pslam_merger_plan = one upload, two launches, one download per frame; latency bound (a few "
                                   "hundred items: the sequential CPU walk is as fast, the device version pays off inside a device-"
                                   "resident pipeline or batched)"}
    except Exception as e:
        out["merger"] = {"error": repr(e)}
    # batched H,b: one launch over 2^22 correspondences
    n = 1 << 22
    rng = np.random.default_rng(0)
    xyzb = np.stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(3, 40, n)], 1)
    h = xyzb @ K.reshape(3, 3).astype(np.float64).T
    measb = np.stack([h[:, 0] / h[:, 2], h[:, 1] / h[:, 2], (h[:, 0] - 386.1448) / h[:, 2], h[:, 1] / h[:, 2]], 1)
    measb[:, :3] += rng.normal(0, 0.5, (n, 3))
    idx = np.arange(n, dtype=np.int32)
    info = np.tile([1.0, 2.0, 1.0], (n, 1))
    cfg = ctx.linearize_cfg("stereo", K, 1241, 376, (-386.1448, 0, 0), 0.0, "saturated", 25.0)
    ms = ctx.linearize_timed(cfg, np.eye(3, 4).reshape(12), xyzb, measb, idx, idx, info, reps=10)
    flop = 250.0 * n  # ~250 fp64 FLOP per correspondence
    out["linearize"] = {"metric": "linearize_gcorr_per_s", "value": n / (ms * 1e-3) / 1e9, "unit": "GCorr/s", "ms": ms,
                        "correspondences": n, "fp64_tflops": flop / (ms * 1e-3) / 1e12,
                        "bytes_per_correspondence": 24 + 32 + 24 + 8,
                        "hbm_gbs": n * 88 / (ms * 1e-3) / 1e9}
    # the reference's own scalar: fp32 clouds in HBM, widened in registers (pslam_linearize_se3_f32) -- 3 + 4 + 3 floats + 2 indices
    try:
        ms32 = ctx.linearize_timed_f32(cfg, np.eye(3, 4).reshape(12), xyzb.astype(np.float32), measb.astype(np.float32), idx, idx,
                                       info.astype(np.float32), reps=10)
        out["linearize"]["f32_clouds"] = {"value": n / (ms32 * 1e-3) / 1e9, "unit": "GCorr/s", "ms": ms32,
                                          "bytes_per_correspondence": 12 + 16 + 12 + 8, "hbm_gbs": n * 48 / (ms32 * 1e-3) / 1e9}
    except Exception as e:
        out["linearize"]["f32_clouds"] = {"error": repr(e)}
    return out


def natural_images_line(capi, dev, work_images):
    """The same device-resident stereo frontend on NATURAL images: the five KITTI pairs of tests/golden (the reference's
    test_data) tiled to 2 000 pairs.  The synthetic batch of the headline is far richer in FAST candidates than a photograph
    (17.7 % of the pixels pass the compass pre-test there); this line shows what the kernels do on real texture."""
    import cv2
    import torch
    G = ROOT / "tests" / "golden"
    pairs = [np.stack([cv2.imread(str(G / f"kitti_city_image_{s_}_{i}.png"), cv2.IMREAD_UNCHANGED) for s_ in ("left", "right")])
             for i in range(5)]
    reps, P = 400, 2000
    batch = torch.from_numpy(np.stack(pairs)).to(dev).repeat(reps, 1, 1, 1).contiguous()  # [2000, 2, 376, 1241]
    ctx = capi.Context(device=dev.index or 0, max_images=2 * P, max_rows=ROWS, max_cols=COLS, max_features=4096,
                       max_raw_per_bin=8192, max_bins=9, work_images=work_images)
    try:
        ecfg, mcfg = capi.extract_cfg(THRESHOLD, 1, TARGET), capi.match_cfg(**MATCH)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        run = lambda: ctx.stereo_frontend_batch_dev(batch.data_ptr(), P, ROWS, COLS, COLS, ROWS * COLS, ecfg, mcfg)
        for _ in range(3):
            run()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            run()
        e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / 5
        counts = ctx.stereo_counts(P)
        feats = ctx.feature_counts(2 * P) if hasattr(ctx, "feature_counts") else None
        ctx.set_lanes(1)
        run()
        ctx.profile_enable(True)
        ctx.synchronize()
        run()
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        return {"metric": "frontend_stereo_frames_per_s", "value": P / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                "config": "5 KITTI stereo pairs of tests/golden (1241x376, the reference's test_data) tiled to 2 000 pairs, same detector / "
                          "matcher parameters as the headline",
                "mean_stereo_points_per_frame": float(counts.mean()),
                "mean_features_per_image": None if feats is None else float(np.mean(feats)),
                "kernels_us_per_image": {k: 1e3 * v[0] / (2 * P) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    finally:
        ctx.close()
        del batch


def scene_traffic():
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        return json.loads(tr.read_text()).get("scene_flags_kernel", {}).get("dram_bytes_per_launch")
    return None


def scene_clip_line(ctx, capi, dev, hbm_peak):
    """N2: SceneClipperProjective3D over a device-resident synthetic local map (16 M points with
    descriptors, about a third visible): one launch projects, clips and compacts in map order.  HBM bound."""
    import torch
    n = 1 << 24
    g = torch.Generator(device=dev).manual_seed(7)
    xyz = (torch.rand((n, 3), generator=g, device=dev) * 60 - 30).contiguous()
    desc = torch.randint(0, 2 ** 31 - 1, (n, 8), generator=g, device=dev, dtype=torch.int32)
    oxyz, ouvz = torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev)
    oidx = torch.empty(n, dtype=torch.int32, device=dev)
    odesc = torch.empty((n, 8), dtype=torch.int32, device=dev)
    K = np.array([718.856, 0, 607.193, 0, 718.856, 185.216, 0, 0, 1], np.float32)
    T = np.array([1, 0, 0, 0.3, 0, 1, 0, -0.1, 0, 0, 1, 1.5], np.float32)
    cfg = capi.clip_cfg(K, ROWS, COLS, T, 0.1, 1000.0)
    torch.cuda.synchronize()
    args = (n, xyz.data_ptr(), desc.data_ptr(), cfg, oxyz.data_ptr(), ouvz.data_ptr(), oidx.data_ptr(), odesc.data_ptr())
    ctx.scene_clip_dev(*args, reps=3)
    kept, ms = ctx.scene_clip_dev(*args, reps=20)
    # per-kernel live times (one CUDA event after every launch: a few microseconds of overhead land on each kernel)
    ctx.profile_enable(True)
    ctx.scene_clip_dev(*args, reps=20)
    prof = {k: 1e3 * t / c for k, (t, c) in ctx.profile_read().items() if k.startswith("scene_")}
    ctx.profile_enable(False)
    alg_all = n * 12 + kept * (12 + 12 + 12 + 4 + 32 + 32)  # map read once; survivors: xyz re-read, xyz / uvz / index written, descriptor copied
    alg_flags = n * 12 + n // 8                             # pass A: the map once, one validity bit per point out
    t_flags = prof.get("scene_flags_kernel", ms * 1e3) * 1e-6
    gbs = alg_flags / t_flags / 1e9
    return {"metric": "scene_clip_gpoints_per_s", "value": n / (ms * 1e-3) / 1e9, "unit": "GPoint/s", "ms": ms,
            "map_points": n, "survivors": int(kept), "alg_gbs_whole_pass": alg_all / (ms * 1e-3) / 1e9,
            "kernels_us": prof,
            "roofline": {"kernel": "scene_flags_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "alg_bytes_per_launch": alg_flags, "traffic": scene_traffic(),
                         "note": "time = event-to-event interval with profiling on (includes ~5 us of event overhead)"}}


def landmarks_ekf_line(ctx, capi, dev):
    """N3: LandmarkEstimatorStereoProjectiveEKF3D over a device-resident batch of 4 M landmarks (one thread
    per landmark, fp64 inside).  Per landmark: 64 B in (state, covariance, measurement), 61 B out; ~1.1 kFLOP fp64."""
    import torch
    n = 1 << 22
    g = torch.Generator(device=dev).manual_seed(11)
    truth = torch.stack([torch.rand(n, generator=g, device=dev) * 8 - 4, torch.rand(n, generator=g, device=dev) * 4 - 2,
                         torch.rand(n, generator=g, device=dev) * 30 + 5], 1)
    state = (truth + torch.randn((n, 3), generator=g, device=dev) * 0.1).contiguous()
    cov = torch.eye(3, device=dev).reshape(1, 9).repeat(n, 1).contiguous()
    fx, cx, cy, bx = 718.856, 607.193, 185.216, 386.1448
    u = fx * truth[:, 0] / truth[:, 2] + cx
    v = fx * truth[:, 1] / truth[:, 2] + cy
    meas = torch.stack([u, v, u - bx / truth[:, 2], v], 1).contiguous()
    local = torch.empty((n, 3), device=dev)
    inl = torch.empty(n, dtype=torch.uint8, device=dev)
    K = np.array([fx, 0, cx, 0, fx, cy, 0, 0, 1], np.float32)
    cfg = capi.ekf_cfg("stereo", K, (bx, 0.0), np.eye(3, 4), np.eye(3, 4), max_cov_norm2=4.0, max_dist2=1.0)
    torch.cuda.synchronize()
    args = (n, state.data_ptr(), cov.data_ptr(), meas.data_ptr(), cfg, local.data_ptr(), inl.data_ptr())
    ctx.landmarks_ekf_update_dev(*args, reps=2)
    cnt, ms = ctx.landmarks_ekf_update_dev(*args, reps=10)
    return {"metric": "landmark_ekf_updates_per_s", "value": n / (ms * 1e-3) / 1e9, "unit": "GLandmark/s", "ms": ms,
            "landmarks": n, "inliers": int(cnt), "hbm_gbs": n * (64 + 61) / (ms * 1e-3) / 1e9,
            "fp64_tflops": n * 1100.0 / (ms * 1e-3) / 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10000, help="stereo pairs per GPU per step")
    ap.add_argument("--work-images", type=int, default=768, help="pipeline chunk (images); two lanes: 384 -> 173.1 k, 768 -> 175.4 k, 1536 -> 178.7 k frames/s device-resident (one lane: 149.4 / 165.3 / 172.9 k)")
    ap.add_argument("--lanes", type=int, default=2, choices=[1, 2], help="chunk lanes of the batched stage-1 calls (pslam_set_lanes)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-pairs", type=int, default=4000, help="stereo pairs per GPU per end-to-end step (pinned host memory: 0.93 MB each)")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="cpu_baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--ref-pairs", type=int, default=256, help="--impl reference: pairs per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-tracking", action="store_true", help="skip the sequential-stage lines (aligner latency, H/b throughput)")
    args = ap.parse_args()
    if os.environ.get("PSLAM_BENCH_WATCHDOG"):  # debugging aid: dump all Python stacks if the run takes too long
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["PSLAM_BENCH_WATCHDOG"]), exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from srrg2_proslam_b200 import capi, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    # ---- synthetic workload, generated on the device (plumbing), one seed per chunk and rank ----
    images = torch.empty((P, 2, ROWS, COLS), dtype=torch.uint8, device=dev)
    for i, base in enumerate(range(0, P, GEN_CHUNK)):
        n = min(GEN_CHUNK, P - base)
        images[base:base + n] = synth.stereo_pairs(n, ROWS, COLS, seed=args.seed + 100003 * rank + i, device=dev)
    torch.cuda.synchronize()
    img_bytes = ROWS * COLS

    ctx = capi.Context(device=local, max_images=2 * P, max_rows=ROWS, max_cols=COLS, max_features=4096,
                       max_raw_per_bin=8192, max_bins=9, work_images=args.work_images)
    ecfg = capi.extract_cfg(THRESHOLD, 1, TARGET)
    mcfg = capi.match_cfg(**MATCH)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    ctx.set_lanes(args.lanes)
    def step_dev():
        ctx.stereo_frontend_batch_dev(images.data_ptr(), P, ROWS, COLS, COLS, img_bytes, ecfg, mcfg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # ---- device-resident leg -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_dev()
    ctx.synchronize()
    counts = ctx.stereo_counts(P)          # also raises on any capacity overflow
    clk_lines, stop = [], threading.Event()
    sampler = threading.Thread(target=clocks_sampler, args=(stop, clk_lines, local), daemon=True)
    sampler.start()
    time.sleep(0.3)
    ctx.profile_enable(True)
    barrier()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    gpu_launches = ctx.launches - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    prof_timed = ctx.profile_read()
    ctx.profile_enable(False)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * P * args.steps / (ms_total * 1e-3)
    # Per-kernel durations.  In the timed region the chunks of a batch alternate over two streams ("lanes"), so an event
    # interval there is the time a kernel SHARED the GPU with the other lane's kernels.  The roofline wants the kernel alone:
    # one more pass over the same batch right after the timed region, same kernels and launch shapes, one lane.
    ctx.set_lanes(1)
    step_dev()
    ctx.profile_enable(True)
    barrier()
    iso_steps = min(args.steps, 2)
    for _ in range(iso_steps):
        step_dev()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.set_lanes(args.lanes)

    # ---- end-to-end leg: pinned host images in, packed stereo clouds out -------------------------
    # E pairs per step (bounded so that 8 ranks do not pin 75 GB of host memory); same images, same call chain
    E = min(P, args.e2e_pairs)
    h_images = torch.empty((E, 2, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
    h_images.copy_(images[:E])
    torch.cuda.synchronize()
    total_pts = int(counts[:E].sum())
    cap_pts = int(total_pts * 1.05) + 1024
    out = {"offsets": np.zeros(E + 1, np.int64)}
    pin = {"uvuv": torch.empty((cap_pts, 4), dtype=torch.float32, pin_memory=True),
           "intensity": torch.empty((cap_pts,), dtype=torch.float32, pin_memory=True),
           "desc": torch.empty((cap_pts, 32), dtype=torch.uint8, pin_memory=True)}
    out.update({k: v.numpy() for k, v in pin.items()})
    want = ("uvuv", "intensity", "desc")

    def step_e2e():
        ctx.stereo_frontend_batch(h_images.data_ptr(), E, ROWS, COLS, COLS, img_bytes, ecfg, mcfg)
        return ctx.download_stereo_batch(E, cap_pts, want=want, out=out)

    for _ in range(2):
        res = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    ctx.synchronize()
    e2e_local_s = time.perf_counter() - t0  # this rank alone (download_stereo_batch has synchronised already)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * E * args.steps / float(e2e_s.item())

    # ---- the end-to-end leg's own roofline: raw pinned host -> device copy rate of THIS box with all N ranks copying at
    # the same time, same source buffer, same chunk size as the pipeline's uploads (work_images images per cudaMemcpyAsync)
    def h2d_probe():
        chunk = min(h_images.numel(), args.work_images * img_bytes)
        src = h_images.view(-1)
        n_chunks = max(1, min(8, src.numel() // chunk))
        dst = torch.empty(chunk, dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(device=dev)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for rep in range(3):  # first repetition = warm-up
            barrier()
            with torch.cuda.stream(side):
                p0.record(side)
                for i in range(n_chunks):
                    dst.copy_(src[i * chunk:(i + 1) * chunk], non_blocking=True)
                p1.record(side)
            side.synchronize()
            ms_ = torch.tensor([p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
            mine = float(ms_.item())
            if world > 1:
                dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
            if rep > 0 and (best is None or float(ms_.item()) < best[0]):
                best = (float(ms_.item()), mine)
        del dst
        agg = world * n_chunks * chunk / (best[0] * 1e-3) / 1e9
        return agg, n_chunks * chunk / (best[1] * 1e-3) / 1e9, chunk

    try:
        h2d_peak_gbs, h2d_rank_gbs, h2d_chunk = h2d_probe()
    except Exception as e:  # the probe must never cost the headline
        h2d_peak_gbs, h2d_rank_gbs, h2d_chunk = None, None, repr(e)
    # per-rank view of the end-to-end leg (explains spreads between runs at the same N: which rank / PCIe slot was slow)
    props = torch.cuda.get_device_properties(local)
    mine = torch.tensor([e2e_local_s, h2d_rank_gbs or 0.0, float(getattr(props, "pci_bus_id", -1))], device=dev, dtype=torch.float64)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[float(v) for v in t.tolist()] for t in per_rank]
    stop.set()
    sampler.join(timeout=3)
    n_pts = int(res["n"])
    h2d = 2 * E * img_bytes
    d2h = (E + 1) * 8 + n_pts * (16 + 4 + 32)

    # ---- roofline of the dominant kernel (live event times over the timed region) ----------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    n_img = 2 * P * args.steps
    mean_feat = float(ctx.feature_counts(2 * P).mean())
    # algorithmic bytes per image of stage 1: image read once + selected keypoints (x, y,
    # response, intensity: 16 B) + descriptors (32 B); per kernel: its own compulsory input + output
    alg = {
        "fast_blur_rows_kernel": ROWS * COLS + mean_feat * 48,   # the stage-1 figure is billed to its dominant kernel
        "bin_select_kernel": mean_feat * 2.2 * 8,  # ~2.2x raw keypoints per kept one: row-list read + raw-list write
        "assemble_features_kernel": mean_feat * (4 + 16),
        "orb_describe_kernel": mean_feat * (31 * 31 + 32),
        "epipolar_kernel": 2 * mean_feat * (32 + 8),
    }
    kernels = {}
    tot_kernel_ms = sum(v[0] for v in prof.values()) or 1.0
    n_img_timed = n_img
    n_img = n_img * iso_steps // args.steps  # images of the one-lane pass the per-kernel times come from
    for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        units = n_img if name != "epipolar_kernel" else n_img / 2
        a = alg.get(name)
        kernels[name] = {"ms_total": kms, "launches": cnt, "share": kms / tot_kernel_ms,
                         "us_per_image": 1e3 * kms / n_img,
                         "alg_gbs": (a * units / (kms * 1e-3) / 1e9) if a and kms > 0 else None}
    top = next(iter(kernels)) if kernels else None
    roofline = None
    if top:
        k = kernels[top]
        ach = k["alg_gbs"] or 0.0
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "avg_launch_ms": k["ms_total"] / max(k["launches"], 1),
                    "alg_bytes_per_image": alg.get(top), "share_of_kernel_time": k["share"]}
        tr = ROOT / "profiles" / "traffic.json"
        if tr.exists():
            t = json.loads(tr.read_text()).get(top)
            if t:
                # the capture was taken at `images_per_launch` images per launch: scale to this run's chunk size
                per_image = t.get("dram_bytes_per_launch", 0) / max(t.get("images_per_launch", args.work_images), 1)
                roofline["traffic"] = int(per_image * n_img / max(k["launches"], 1))
                roofline["traffic_source"] = t.get("source")
                # why the HBM fraction is small: the kernel is instruction-issue bound (same ncu capture)
                roofline["ncu"] = {k: t[k] for k in ("issue_active_pct", "alu_pipe_pct", "dram_throughput_pct", "registers", "ctas_per_sm") if k in t}
                # The kernel is integer work bound by instruction issue, not by HBM: second roofline against the issue slots
                # (4 warp instructions per clock per SM).  Instruction count from the committed ncu capture, time measured live.
                if t.get("warp_instructions_per_launch"):
                    wi = t["warp_instructions_per_launch"] / max(t.get("images_per_launch", args.work_images), 1)  # per image
                    sm_mhz_ = 1965.0
                    k_ms = k["ms_total"]
                    ach = wi * n_img / (k_ms * 1e-3) / 1e9
                    pk_issue = 148 * 4 * sm_mhz_ * 1e6 / 1e9
                    roofline["issue"] = {"bound": "issue", "achieved": ach, "peak": pk_issue, "unit": "G warp-instructions/s",
                                         "frac": ach / pk_issue, "warp_instructions_per_image": wi,
                                         "thread_instructions_per_pixel": wi * 32 / (ROWS * COLS),
                                         "peak_source": "148 SM x 4 schedulers x 1965 MHz",
                                         "note": "this ceiling only moves with the instruction count: thread_instructions_per_pixel is the tracked figure"}

    line = {"metric": "frontend_stereo_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, P),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pairs_per_gpu_per_step": E,
                    "api": "pslam_stereo_frontend_batch + pslam_download_stereo_batch (host pointers)",
                    "roofline": None if h2d_peak_gbs is None else {
                        "bound": "pcie-h2d", "achieved_gbs": e2e_value * h2d / E / 1e9, "peak_gbs": h2d_peak_gbs,
                        "frac": (e2e_value * h2d / E / 1e9) / h2d_peak_gbs, "unit": "GB/s",
                        "peak_source": f"measured in this job: {world} rank(s) copying concurrently from pinned host memory, "
                                       f"{h2d_chunk >> 20} MiB per cudaMemcpyAsync (the pipeline's chunk), max over ranks"},
                    "per_rank": [{"rank": i, "e2e_s": r_[0], "frames_per_s": E * args.steps / r_[0] if r_[0] > 0 else None,
                                  "h2d_probe_gbs": r_[1], "pci_bus_id": int(r_[2])} for i, r_ in enumerate(per_rank)]},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "kernels": kernels,
            "kernels_note": f"per-kernel times: {iso_steps} step(s) of the same batch right after the timed region with ONE chunk lane "
                            "(kernels strictly one after the other); in the timed region the chunks alternate over "
                            f"{args.lanes} lane(s) and kernels of neighbouring chunks overlap",
            "kernels_timed_region": {k_: {"stream_ms_total": v_[0], "launches": v_[1], "us_per_image": 1e3 * v_[0] / n_img_timed}
                                     for k_, v_ in sorted(prof_timed.items(), key=lambda kv: -kv[1][0])},
            "lanes": {"n": args.lanes, "sum_of_isolated_kernel_ms_per_step": tot_kernel_ms / iso_steps,
                      "ms_per_step": ms_total / args.steps},
            "clocks": summarise_clocks(clk_lines),
            "mean_features_per_image": mean_feat, "mean_stereo_points_per_frame": float(counts.mean())}

    # ---- secondary metric: Hamming GPair/s (BASELINE config 5, 64k x 64k) -------------------------
    # query rows sharded by rank, train set replicated; the per-row (best, second, argmin) table is assembled on
    # every rank by ONE all-gather (NCCL over NVLink) inside the timed region -> strong scaling
    if not args.no_hamming:
        nq = nt = 65536
        q, t = synth.hamming_sets(nq, nt, seed=args.seed)
        # N > 1: the C ABI's own multi-GPU entry point (pslam_bf_best2_sharded_dev): its NCCL communicator is created through
        # the ABI as well (the 128-byte id travels over torch.distributed), the all-gather runs on the context's stream
        dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
        table = torch.empty((3, nq), dtype=torch.int32, device=dev)
        comm, p2p = None, False
        if world > 1:
            # preferred: exchange fused into the merge kernel (peer tables over CUDA IPC + NVLink, no collective); fallback: the
            # NCCL all-gather entry point
            mine = None
            try:
                mine = ctx.p2p_table_export(nq)
            except Exception as e_:  # e.g. IPC not permitted in this container
                print(f"[bench] p2p table export failed ({e_!r})", file=sys.stderr)
            handles = [None] * world
            dist.all_gather_object(handles, mine)  # every rank takes part, whatever happened above
            if all(h is not None for h in handles):
                try:
                    ctx.p2p_table_import(rank, world, handles)
                    p2p = True
                except Exception as e_:
                    print(f"[bench] p2p table import failed ({e_!r}), using the NCCL entry point", file=sys.stderr)
            flag = torch.tensor([int(p2p)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            p2p = bool(flag.item())
            if not p2p:
                ids = [ctx.nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                comm = ctx.nccl_comm_create(ids[0], rank, world)
        def sweep():
            if p2p:
                ctx.bf_best2_sharded_p2p_dev(nq, dq.data_ptr(), nt, dt_.data_ptr(), table[0].data_ptr(), table[1].data_ptr(),
                                             table[2].data_ptr())
            elif world > 1:
                ctx.bf_best2_sharded_dev(comm, rank, world, nq, dq.data_ptr(), nt, dt_.data_ptr(), table[0].data_ptr(),
                                         table[1].data_ptr(), table[2].data_ptr())
            else:
                ctx.bf_best2_dev(nq, dq.data_ptr(), nt, dt_.data_ptr(), table[0].data_ptr(), table[1].data_ptr(), table[2].data_ptr())
        for _ in range(3):
            sweep()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        reps = 10
        for _ in range(reps):
            sweep()
        h1.record(stream)
        barrier()
        hms = torch.tensor([h0.elapsed_time(h1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(hms, op=dist.ReduceOp.MAX)
        gpairs = nq * nt * reps / (float(hms.item()) * 1e-3) / 1e9
        if comm is not None:
            ctx.nccl_comm_destroy(comm)
        if world > 1:
            barrier()
            ctx.p2p_table_release()  # no-op when nothing was exported
        sm_mhz = line["clocks"]["sm_mhz"] or 1965.0
        # Roofline of the sweep = the busier of the two integer pipes for the kernel's instruction mix per 256-bit pair
        # (SASS of bf_sweep_kernel<0>: 5 POPC on the XU pipe; 15 LOP3 + 1 IADD3 + 3 VIMNMX on the ALU pipe), with the pipe
        # rates MEASURED on this GPU by tools/int_peaks.cu (profiles/int_peaks.json), not taken from a guide.
        popc_rate, alu_rate, rate_src = 16.0, 64.0, "fallback 16 POPC / 64 ALU results per clk per SM"
        ip = ROOT / "profiles" / "int_peaks.json"
        if ip.exists():
            pk_ = json.loads(ip.read_text())
            ev = lambda d: next(v for k, v in d.items() if k.startswith("by_event_time"))
            popc_rate, alu_rate = ev(pk_["popc"]), min(ev(pk_["lop3"]), ev(pk_["iadd3"]))
            rate_src = f"measured {popc_rate:.2f} POPC and {alu_rate:.2f} ALU results per clk per SM (profiles/int_peaks.json)"
        POPC_PER_PAIR, ALU_PER_PAIR = 5, 19
        pairs_per_clk_sm = min(popc_rate / POPC_PER_PAIR, alu_rate / ALU_PER_PAIR)
        int_peak = 148 * pairs_per_clk_sm * sm_mhz * 1e6 / 1e9 * world
        plain_peak = 148 * popc_rate / 8 * sm_mhz * 1e6 / 1e9 * world
        line["hamming"] = {"metric": "hamming_best2_gpairs_per_s", "value": gpairs, "unit": "GPair/s", "scaling": "strong",
                           "config": {"workload": "64k x 64k 256-bit descriptors (BASELINE config 5): query rows sharded "
                                                  f"over {world} GPU(s), train set replicated, every rank ends with the full (best, second, argmin) table",
                                      "api": "pslam_bf_best2_dev" if world == 1 else (
                                          "pslam_bf_best2_sharded_p2p_dev (merge kernel stores into every rank's table over NVLink, flags; no collective)"
                                          if p2p else "pslam_bf_best2_sharded_dev (ncclAllGather on the context stream)")},
                           "ms_per_sweep": float(hms.item()) / reps,
                           "roofline": {"bound": "int (XU popc + ALU lop3, balanced)", "achieved": gpairs, "peak": int_peak, "unit": "GPair/s",
                                        "frac": gpairs / int_peak,
                                        "peak_source": f"{world} x 148 SM x min({popc_rate:.2f} / {POPC_PER_PAIR} POPC, {alu_rate:.2f} / "
                                                       f"{ALU_PER_PAIR} ALU) pairs per clk x {sm_mhz:.0f} MHz (sampled); {rate_src}",
                                        "plain_8_popc_roofline": plain_peak,
                                        "note": "the textbook form (8 XOR + 8 POPC per pair) is XU-bound at plain_8_popc_roofline; the kernel "
                                                "reduces the 8 XOR words with three carry-save adders first (5 POPC + 14 LOP3)"}}
        # the whole reference call on config 5: pslam_match_bruteforce = counting sweep + candidate sweep + sort + bijective resolve
        if rank == 0:
            try:
                tm0 = time.perf_counter()
                mfi, mmi, md = ctx.match_bruteforce(q, t, capi.match_cfg(50.0, 0.9))
                tm1 = time.perf_counter()
                mfi, mmi, md = ctx.match_bruteforce(q, t, capi.match_cfg(50.0, 0.9))
                tm2 = time.perf_counter()
                line["hamming"]["match_bruteforce"] = {"ms": 1e3 * min(tm1 - tm0, tm2 - tm1), "matches": int(len(mfi)),
                                                       "config": "pslam_match_bruteforce(64k x 64k, max_dist 50, ratio 0.9), host "
                                                                 "descriptors in, bijective correspondences out (wall clock)"}
            except Exception as e:
                line["hamming"]["match_bruteforce"] = {"error": repr(e)}

    # ---- sequential stage (latency in microseconds, FP64 throughput on a batched synthetic) ------
    if rank == 0 and not args.no_tracking:
        try:
            line["natural_images"] = natural_images_line(capi, dev, args.work_images)
        except Exception as e:
            line["natural_images"] = {"error": repr(e)}
    if rank == 0 and not args.no_tracking:
        try:
            line["tracking"] = tracking_lines(ctx, capi, stream, dev)
        except Exception as e:  # secondary lines must never cost the headline
            line["tracking"] = {"error": repr(e)}

    if rank == 0 and not args.no_tracking:
        try:
            line["scene_clip"] = scene_clip_line(ctx, capi, dev, hbm_peak)
        except Exception as e:
            line["scene_clip"] = {"error": repr(e)}
        try:
            line["landmark_ekf"] = landmarks_ekf_line(ctx, capi, dev)
        except Exception as e:
            line["landmark_ekf"] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample -----------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O = oracle_lib()
        cores = os.cpu_count() or 1
        ocfg = O.extract_cfg(THRESHOLD, 1, TARGET)
        probe = images[:cores].cpu().numpy()
        t0 = time.perf_counter()
        O.stereo_frontend_batch(probe, ocfg, threads=cores, **MATCH)
        per_round = time.perf_counter() - t0
        sample = args.cpu_pairs or int(min(P, max(cores, cores * round(12.0 / max(per_round, 1e-3)))))
        simgs = images[:sample].cpu().numpy()
        t0 = time.perf_counter()
        ccounts, cchk = O.stereo_frontend_batch(simgs, ocfg, threads=cores, **MATCH)
        cdt = time.perf_counter() - t0
        n1 = max(4, min(sample, int(round(4.0 / max(cdt / sample * cores, 1e-3)))))  # ~4 s on one thread
        t0 = time.perf_counter()
        O.stereo_frontend_batch(simgs[:n1], ocfg, threads=1, **MATCH)
        sdt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / cdt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} stereo pairs of the same batch, {cdt:.1f} s, all host threads",
                                "all_threads": sample / cdt, "single_thread": n1 / sdt,
                                "single_thread_sample": f"first {n1} pairs, {sdt:.1f} s",
                                "build": "oracle/Makefile: g++ -O3 -march=x86-64-v3 (AVX2), frame-level std::thread sharding",
                                "parity_counts_equal": bool(np.array_equal(ccounts, counts[:sample]))}
        # context: OpenCV's own SIMD FAST + ORB::compute + the epipolar match of the port, one thread,
        # on the same images -- what the reference's third-party arithmetic costs when it is OpenCV's build, not our port
        try:
            import cv2
            cv2.setNumThreads(1)
            fast = cv2.FastFeatureDetector_create(THRESHOLD, True)
            orb = cv2.ORB_create()
            nimg = min(16, sample)
            tf = td = 0.0
            nk = 0
            for i in range(nimg):
                for side in range(2):
                    im = np.ascontiguousarray(simgs[i, side])
                    t0 = time.perf_counter()
                    kps = fast.detect(im)
                    t1 = time.perf_counter()
                    kps = sorted(kps, key=lambda k: -k.response)[:int(line["mean_features_per_image"])]
                    t2 = time.perf_counter()
                    orb.compute(im, kps)
                    t3 = time.perf_counter()
                    tf += t1 - t0
                    td += t3 - t2
                    nk += len(kps)
            line["cpu_baseline"]["cv2_context"] = {"cv2": cv2.__version__, "threads": 1, "images": 2 * nimg,
                                                   "fast_ms_per_image": 1e3 * tf / (2 * nimg), "orb_compute_ms_per_image": 1e3 * td / (2 * nimg),
                                                   "keypoints_described_per_image": nk / (2 * nimg),
                                                   "stage1_frames_per_s_single_thread": 1.0 / (tf / nimg + td / nimg),
                                                   "note": "cv2.FastFeatureDetector + ORB.compute (SIMD builds) on the same images; not bit-identical to "
                                                           "the OpenCV-3 arithmetic the reference pins (the port is), reported for scale only"}
        except Exception as e:
            line["cpu_baseline"]["cv2_context"] = {"error": repr(e)}
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

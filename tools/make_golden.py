#!/usr/bin/env python3
"""Build tests/golden/ input fixtures from the reference's test_data (run in the build
container only; /root/reference does not exist on the GPU box).

Inputs are decoded exactly as the reference's fixtures do
(tests/fixtures.hpp:438,729-730,1067-1069: cv::imread(..., CV_LOAD_IMAGE_GRAYSCALE) for
intensity, CV_LOAD_IMAGE_ANYDEPTH for depth) and re-encoded losslessly as 8-bit gray PNG /
16-bit PNG so that every consumer sees the same pixels without needing a colour conversion.
"""
import hashlib, json, pathlib, shutil, sys
import cv2

REF = pathlib.Path("/root/reference/test_data")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

gray = [f"kitti/city/image_{s}_{i}.png" for i in range(5) for s in ("left", "right")]
gray += [f"kitti/highway/image_{s}_{i}.png" for i in (274, 275) for s in ("left", "right")]
gray += [f"icl/image_rgb_{i}.png" for i in (0, 1, 50)]
gray += ["scene_flow/image_left.png", "scene_flow/image_right.png"]
depth = [f"icl/image_depth_{i}.pgm" for i in (0, 1, 50)]
manifest = {}
for rel in gray:
    im = cv2.imread(str(REF / rel), cv2.IMREAD_GRAYSCALE)
    assert im is not None and im.ndim == 2, rel
    dst = OUT / rel.replace("/", "_")
    cv2.imwrite(str(dst), im, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert (cv2.imread(str(dst), cv2.IMREAD_UNCHANGED) == im).all()
    manifest[dst.name] = {"src": "test_data/" + rel, "shape": list(im.shape),
                          "sha256_pixels": hashlib.sha256(im.tobytes()).hexdigest()}
for rel in depth:
    im = cv2.imread(str(REF / rel), cv2.IMREAD_ANYDEPTH)
    assert im is not None and im.dtype.name == "uint16", rel
    dst = OUT / (rel.replace("/", "_").replace(".pgm", ".png"))
    cv2.imwrite(str(dst), im, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert (cv2.imread(str(dst), cv2.IMREAD_UNCHANGED) == im).all()
    manifest[dst.name] = {"src": "test_data/" + rel, "shape": list(im.shape), "dtype": "uint16",
                          "sha256_pixels": hashlib.sha256(im.tobytes()).hexdigest()}
shutil.copy(REF / "scene_flow/gt_stereo_matching_threshold-100.txt",
            OUT / "scene_flow_gt_stereo_matching_threshold-100.txt")
(OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")
print(len(manifest), "fixtures ->", OUT)

# ---- hot-path subsets of the shipped configurations -------------------------------------------------------
# The full files are read UNCHANGED from /root/reference/configurations by our BOSS reader
# (srrg2_proslam_b200/host/pslam_boss.cpp) and the modules on the frontend path -- with everything they link
# to -- are written back by our writer.  The GPU box has no /root/reference, so these derived files are what the
# `-m gpu` plugin tests load there; tests/test_plugin_cpu.py checks here that they agree with the originals.
sys.path.insert(0, str(OUT.parent.parent))
from srrg2_proslam_b200 import plugin as P  # noqa: E402

CONF = pathlib.Path("/root/reference/configurations")
HOT = {"kitti": ["adaptor_stereo_projective", "aligner", "cf_bruteforce", "clipper_stereo_projective", "landmark_estimator_ekf",
                 "landmark_estimator_weighted_mean", "landmark_estimator_smoother", "merger_triangulation", "merger_ekf"],
       "euroc": ["adaptor_stereo_projective", "aligner", "cf_bruteforce", "clipper_stereo_projective", "merger_triangulation",
                 "merger_ekf"],
       "icl": ["tracker_slice_processor_projective_depth", "aligner", "cf_bruteforce_2d", "cf_bruteforce_3d"]}
(OUT / "configurations").mkdir(exist_ok=True)
for name, roots in HOT.items():
    m = P.Manager(CONF / f"{name}.conf")
    roots = [r for r in roots if any(x.name == r for x in m.modules())]
    m.write(OUT / "configurations" / f"{name}_hotpath.conf", roots)
    print(name, "->", roots)

"""CPU-side checks: the C-ABI library loads and exports every symbol include/pslam_cuda.h declares,
fails loudly without a GPU (no CPU fallback), and the std::sort replay matches real libstdc++."""
import ctypes
import pathlib
import re
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", str(ROOT / "srrg2_proslam_b200" / "csrc"), "-j8", "-s"], check=True)
    from srrg2_proslam_b200 import capi
    return capi


def test_exports_every_declared_symbol(built):
    hdr = (ROOT / "include" / "pslam_cuda.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pslam_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    lib = built.lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_struct_sizes(built):
    assert b"sm_100a" in built.lib().pslam_version()
    assert ctypes.sizeof(built.Limits) == 28 and ctypes.sizeof(built.ExtractCfg) == 20
    assert ctypes.sizeof(built.MatchCfg) == 16


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.PslamError) as e:
        built.Context()
    assert e.value.code == built.PSLAM_E_CUDA


def test_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", str(built.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_sort_replay_matches_libstdcxx(tmp_path):
    exe = tmp_path / "sort_emul_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe),
                    str(ROOT / "tests" / "native" / "sort_emul_check.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout

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


def _build_native_caller(tmp_path):
    exe = tmp_path / "abi_smoke"
    lib_dir = ROOT / "srrg2_proslam_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "native" / "abi_smoke.c"), "-L", str(lib_dir), "-lpslam_cuda",
                    f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_headers_are_plain_c99_and_a_native_caller_links(built, tmp_path):
    """the drop-in boundary is a C ABI: both public headers compile as C99 (-pedantic) and a C program links against the
    library without Python / C++ / torch; without a GPU it must be refused with PSLAM_E_CUDA (tests/native/abi_smoke.c)"""
    src = tmp_path / "hdr.c"
    src.write_text('#include "pslam_cuda.h"\n#include "pslam_plugin.h"\nint main(void) { return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"),
                    str(src)], check=True)
    exe = _build_native_caller(tmp_path)
    import torch
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "NO_DEVICE" in r.stdout, r.stdout + r.stderr

"""CPU checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every symbol
include/svo_b200.h declares; without a device the product fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparse-voxel-octree-raycasting_b200", "libsvo_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.dirname(LIB)], check=True)
    return ctypes.CDLL(LIB)


def declared_symbols():
    names = set()
    for h in ("svo_b200.h", "svo_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(svo_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_ocl_surface():
    names = declared_symbols()
    for must in ("svo_init", "svo_exit", "svo_get_kernel", "svo_malloc", "svo_copy_to_host", "svo_begin", "svo_param",
                 "svo_end", "svo_begin_all_kernels", "svo_end_all_kernels", "svo_memcpy", "svo_memset", "svo_round_up",
                 "svo_frame_fused"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_mirror_matches_the_header():
    """The ctypes mirror binds every frame-level entry point the header declares and uses the header's flag values."""
    text = open(os.path.join(ROOT, "include", "svo_b200.h")).read()
    flags = {k: int(v) for k, v in re.findall(r"\b(SVO_FRAME_[A-Z0-9_]+)\s*=\s*(\d+)", re.sub(r"/\*.*?\*/", "", text, flags=re.S))}
    assert flags == {"SVO_FRAME_PINGPONG": 1, "SVO_FRAME_CACHE_ROTATION": 2, "SVO_FRAME_TEX_RGB24": 4}
    mirror = open(os.path.join(ROOT, "sparse-voxel-octree-raycasting_b200", "ocl.py")).read()
    for name, value in flags.items():
        assert re.search(rf"^{name[4:]}\s*=\s*{value}\b", mirror, flags=re.M), name
    for sym in ("svo_frame_fused", "svo_frame_deferred_count", "svo_frame_early_count", "svo_present_async", "svo_present_rgb24_async",
                "svo_present_wait", "svo_raycast_batch", "svo_debug_set"):
        assert f'"{sym}"' in mirror, sym


def test_round_up(lib):
    lib.svo_round_up.restype = ctypes.c_size_t
    assert lib.svo_round_up(16, 240) == 240 and lib.svo_round_up(16, 540) == 544 and lib.svo_round_up(256, 1) == 256


def test_no_cpu_fallback_without_gpu():
    """On a machine without a CUDA device svo_init must fail (non-zero), never succeed silently."""
    code = (
        "import ctypes,sys; l=ctypes.CDLL(%r); l.svo_set_error_mode(1); n=l.svo_device_count(); "
        "rc=l.svo_init(0); sys.exit(0 if (n>0 and rc==0) or (n==0 and rc!=0) else 3)" % LIB)
    assert subprocess.run([sys.executable, "-c", code], stderr=subprocess.DEVNULL).returncode == 0


def test_product_does_not_import_oracle():
    """The product package and the C/CUDA sources never import, link or load anything under oracle/."""
    pkg = os.path.dirname(LIB)
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|svo_oracle|libsvo_ref|orc_[a-z]|ref_[a-z]+\()")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(src), (f, bad.search(src).group(0))

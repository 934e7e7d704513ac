"""CPU-side checks of the drop-in boundary: the C-ABI library builds (nvcc cross-compiles sm_100a
without a GPU), loads, exports every symbol include/cover_b200.h declares, and refuses to work
without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cover_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cover_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import cover_b200
    path = cover_b200.build()
    return ctypes.CDLL(path)


def test_header_declares_the_expected_entries():
    names = declared_functions()
    for must in ("cover_area_is_free", "cover_area_accumulate_cost", "cover_outline_is_free", "cover_ray_is_free",
                 "cover_cells_is_free", "cover_footprints_is_free", "cover_split_is_free", "cover_map_create", "cover_ctx_create"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cover_abi_version() == 2


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings): compile it with gcc -std=c99."""
    src = '#include "cover_b200.h"\nint main(void){cover_poses p; cover_results r; (void)p; (void)r; return COVER_OK;}\n'
    proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                          input=src, text=True, capture_output=True)
    assert proc.returncode == 0, proc.stderr


def test_sass_is_sm100a_only(lib):
    import cover_b200
    out = subprocess.run(["cuobjdump", "-lelf", cover_b200.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import cover_b200
    with pytest.raises(cover_b200.CoverError) as err:
        cover_b200.Context(0)
    assert err.value.code == -2  # COVER_E_CUDA


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cover_b200/ or include/ may reference it."""
    for base in ("cover_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    for needle in ("import oracle", "from oracle", "pyoracle", "liboracle", "cover_oracle", "oracle/"):
                        assert needle not in text, (os.path.join(dirpath, f), needle)


def test_cpp_host_layer_compiles_and_links(lib, tmp_path):
    """include/cover_b200.hpp (the C++ mirror of the reference's surface) builds against the C ABI."""
    import cover_b200
    exe = tmp_path / "host_api_test"
    libdir = os.path.dirname(cover_b200.library_path())
    proc = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-o", str(exe), "-L", libdir, "-lcover_b200",
                           f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr


def test_plain_c_client_compiles_and_links(lib, tmp_path):
    """tests/c/abi_smoke.c: the ABI consumed from C99 (no C++ in the client)."""
    import cover_b200
    libdir = os.path.dirname(cover_b200.library_path())
    proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                           "-o", str(tmp_path / "abi_smoke"), "-L", libdir, "-lcover_b200", f"-Wl,-rpath,{libdir}", "-lm"], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr

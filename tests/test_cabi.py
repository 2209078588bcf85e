"""The C-ABI library loads on a GPU-less box and exports every symbol include/b2d.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from ray_lightning_b200 import _b2d


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b2d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_b2d.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _b2d.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.b2d_version() == 110


def test_handle_blob_size_matches_header():
    txt = open(os.path.join(ROOT, "include", "b2d.h")).read()
    assert int(re.search(r"#define B2D_HANDLE_BYTES (\d+)", txt).group(1)) == _b2d.HANDLE_BYTES
    assert int(re.search(r"#define B2D_MAX_WORLD (\d+)", txt).group(1)) == _b2d.MAX_WORLD


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_b2d.B2DError) as ei:
        _b2d.Context(0, 1, 0, 1 << 20)
    assert "CUDA" in str(ei.value)
    from ray_lightning_b200.comm import Communicator
    with pytest.raises(_b2d.B2DUnavailableError):
        Communicator(0, 1, 0, 1 << 20)


def test_bad_arguments_are_rejected_before_cuda():
    lib = _b2d.load()
    out = ctypes.c_void_p()
    assert lib.b2d_ctx_create(3, 2, 0, 1 << 20, 0, ctypes.byref(out)) == -1   # rank >= world
    assert lib.b2d_ctx_create(0, 9, 0, 1 << 20, 0, ctypes.byref(out)) == -1   # world > B2D_MAX_WORLD
    assert b"world" in lib.b2d_last_error(None)
    assert lib.b2d_ctx_destroy(None) == 0


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ray_lightning_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src


def test_header_is_plain_c_and_usable_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99, and a C program must be able to dlopen the
    library and drive it (examples_c/b2d_probe.c) — no Python, no torch, no C++ in the signatures."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "b2d.h")
    subprocess.run(["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Werror", "-x", "c", hdr], check=True)
    exe = str(tmp_path / "b2d_probe")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples_c", "b2d_probe.c"), "-o", exe, "-ldl"], check=True)
    out = subprocess.run([exe, _b2d.lib_path()], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "libb2d version 110" in out.stdout

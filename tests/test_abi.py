"""The C-ABI library loads and exports every symbol include/scoary_b200.h declares; without a
GPU it refuses to create a context (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scoary_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from scoary_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libscoary_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert lib.sb_version() == 100


def test_stats_struct_matches_header():
    from scoary_b200 import _lib
    text = open(os.path.join(ROOT, "include", "scoary_b200.h")).read()
    body = text[text.index("typedef struct {"):text.index("} sb_stats_t;")]
    fields = re.findall(r"\b(?:int64_t|int32_t|double)\s+([a-z0-9_]+)\s*;", body)
    assert fields == [f for f, _ in _lib.SbStats._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scoary_b200 import _lib
    from scoary_b200.engine import Engine, EngineError
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.sb_create(0, ctypes.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b"no CUDA device" in lib.sb_last_error(None) or b"CUDA" in lib.sb_last_error(None)
    with pytest.raises(EngineError):
        Engine(0)


def test_oracle_is_not_imported_by_the_product():
    """Nothing under scoary_b200/ may import or load the oracle."""
    pkg = os.path.join(ROOT, "scoary_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libscoary_oracle" not in src, f


def test_header_is_plain_c_and_links(tmp_path):
    """include/scoary_b200.h compiles as C and a C program can drive the library."""
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.join(ROOT, "scoary_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", lib_dir, "-lscoary_b200",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr

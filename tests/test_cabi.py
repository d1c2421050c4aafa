"""The C-ABI library builds, loads, and exports every symbol include/aesmc_b200.h declares
(no compute calls: runs on the CPU-only box)."""
import os
import re

from aesmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aesmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aesmc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_binding_covers_header(built):
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_metadata_calls(built):
    lib = _lib.load()
    assert lib.aesmc_version() >= 100
    assert _lib.max_particles_single_cta() >= 8192
    assert _lib.launch_count() >= 0


def test_bad_arguments_rejected_without_gpu(built):
    lib = _lib.load()
    # null pointers are rejected by argument validation before any CUDA call
    rc = lib.aesmc_smc_step_f32(None, None, None, None, 4, 8, None, None, None, None, None, 1, None, 0, None)
    assert rc == _lib.ERR_BAD_ARG
    assert "aesmc_smc_step_f32" in _lib.last_error()
    rc = lib.aesmc_gather_bytes(None, None, 0, 1, 1, 4, None, None, None)
    assert rc == _lib.ERR_BAD_ARG


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        return
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs

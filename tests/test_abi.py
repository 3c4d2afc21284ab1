"""CPU: the C-ABI library loads, exports every symbol include/spvd_b200.h declares, and the ctypes
signature table covers exactly that set (no compute calls: there is no GPU here)."""
import os
import re
import subprocess

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "spvd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(spvd_[a-z0-9_]+)\s*\(", src))


def test_library_builds_and_exports_every_declared_symbol():
    from spvd_b200 import build
    path = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (spvd_[a-z0-9_]+)", out))
    declared = header_functions()
    assert declared, "no declarations parsed from the header"
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert exported <= declared, f"exported but not declared: {sorted(exported - declared)}"


def test_ctypes_table_matches_header():
    from spvd_b200 import _C
    assert set(_C.SIGNATURES) == header_functions()
    assert _C.lib.spvd_abi_version() == 1
    assert _C.lib.spvd_last_error() == b"" or isinstance(_C.lib.spvd_last_error(), bytes)


def test_argument_validation_without_a_gpu():
    """Host-side argument checks return SPVD_ERR_INVALID and a message before any CUDA call."""
    from spvd_b200 import _C
    rc = _C.lib.spvd_hash_query(None, 5, 0, 7, None, None, 16, None, None)
    assert rc == 1 and b"spvd_hash_query" in _C.lib.spvd_last_error()
    assert _C.lib.spvd_conv_umma_supported(27, 64, 96) == 1
    assert _C.lib.spvd_conv_umma_supported(27, 3, 32) == 0
    assert _C.lib.spvd_unique_workspace_bytes(65536) > 2 * 65536 * 8


def test_sass_has_blackwell_tensor_core_and_bulk_copy_instructions():
    from spvd_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out or "UTCMMA" in out      # tcgen05.mma
    assert "LDTM" in out                             # tcgen05.ld
    assert "UBLKCP" in out                           # cp.async.bulk
    assert "LDGSTS" in out                           # cp.async row gather

"""The drop-in boundary: libsimbackend_cuda.so exports exactly what the reference host binds
(crates/ppf-cts-solver/src/backend.rs extern block; cpp/main/main.cu), include/ppf_abi.h mirrors the
reference's repr(C) layouts, and the library loads without a GPU (no compute calls here)."""
import ctypes
import os
import subprocess

import pytest

from conftest import PKG, REPO


def test_backend_exports_the_20_reference_symbols(ppf):
    sym = ppf.exported_symbols(ppf.LIB_OURS)
    assert len(ppf.ABI_SYMBOLS) >= 20  # 19 functions + the fatal-state data symbols + the print_rust hook
    missing = [s for s in ppf.ABI_SYMBOLS if s not in sym]
    assert not missing, missing
    for s in ppf.EXT_SYMBOLS:
        assert s in sym


def test_every_symbol_declared_in_the_header_is_exported(ppf):
    import re

    hdr = open(os.path.join(REPO, "include", "ppf_abi.h")).read()
    body = hdr[hdr.index('extern "C" {'):]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    names = set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", body))
    names -= {"sizeof", "offsetof", "_Static_assert", "static_assert", "PPF_STATIC_ASSERT"}
    sym = ppf.exported_symbols(ppf.LIB_OURS)
    assert len(names) >= 24, names
    assert not [n for n in names if n not in sym], [n for n in names if n not in sym]


def test_backend_exports_nothing_else_from_the_oracle_or_torch(ppf):
    sym = ppf.exported_symbols(ppf.LIB_OURS)
    assert not [s for s in sym if s.startswith("oracle_") or "torch" in s.lower() or "c10" in s]
    # the product library must not link the oracle or any CPU fallback
    out = subprocess.run(["ldd", ppf.LIB_OURS], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out


def test_library_loads_without_a_gpu(ppf):
    lib = ctypes.CDLL(ppf.LIB_OURS)
    lib.ppf_fatal_code.restype = ctypes.c_ubyte
    assert lib.ppf_fatal_code() == 0


def test_header_layouts_match_the_reference_sizes(tmp_path):
    """DataSet 1456 B, ParamSet 136 B, StepResult 16 B (data.hpp; static_asserts in the header), in C and C++."""
    src = tmp_path / "t.c"
    src.write_text('#include "ppf_abi.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(PpfDataSet), sizeof(PpfParamSet), sizeof(PpfStepResult), sizeof(PpfConstraint));return 0;}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[:3] == ["1456", "136", "16"]
    srcpp = tmp_path / "t.cpp"
    srcpp.write_text('#include "ppf_abi.h"\nint main(){return sizeof(PpfDataSet) == 1456 ? 0 : 1;}\n')
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(REPO, "include"), str(srcpp), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_product_has_no_cpu_fallback(ppf, tmp_path):
    """A missing CUDA backend is a hard error in the host API, never a silent oracle / CPU path."""
    with pytest.raises(RuntimeError):
        ppf.Session("sheet", n=8, lib=str(tmp_path / "libsimbackend_cuda.so"))
    src = open(os.path.join(PKG, "ppf_b200.py")).read()
    assert "LIB_ORACLE" in src and "lib=LIB_OURS" in src  # the default backend is the CUDA library

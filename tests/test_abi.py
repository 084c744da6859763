"""The C-ABI library loads and exports every symbol include/euc_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import euc_b200
from euc_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "euc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\*|int|uint64_t)\s+(euc_\w+)\s*\(", src, flags=re.M)))


def test_header_and_ctypes_table_agree():
    assert header_functions() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(euc_b200.LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(euc_b200.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} not exported"
    lib.euc_abi_version.restype = ctypes.c_int
    assert lib.euc_abi_version() == abi.ABI_VERSION


def test_struct_layouts_match_header():
    assert ctypes.sizeof(abi.SamplerDesc) == 24
    assert ctypes.sizeof(abi.PipelineDesc) == 48 + 16 + 2 * 24
    assert ctypes.sizeof(abi.BatchDraw) == 16 and ctypes.sizeof(abi.RenderStats) == 24
    assert euc_b200.VERTEX_PN.itemsize == 24 and euc_b200.VERTEX_P4UV.itemsize == 32
    assert euc_b200.VERTEX_P4C4.itemsize == 32 and euc_b200.VERTEX_VOXEL.itemsize == 32


def test_no_cpu_fallback_without_device():
    """euc_init must fail (not fall back) when no CUDA device is visible."""
    import subprocess, sys
    code = ("import os; os.environ['CUDA_VISIBLE_DEVICES']='';"
            "import sys; sys.path.insert(0, %r); import euc_b200\n"
            "try:\n    euc_b200.Context(0); print('CREATED')\nexcept euc_b200.EucError as ex:\n    print('RAISED', ex.code)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300).stdout
    assert "RAISED" in out and "CREATED" not in out


def test_product_does_not_import_oracle():
    """Nothing under euc_b200/ may import, include, link or load anything from oracle/."""
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#include\s*[\"<][^\">]*oracle)|(libeuc_oracle)|(oracle/)", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "euc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                assert not bad.search(open(os.path.join(dirpath, f)).read()), os.path.join(dirpath, f)

"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads, exports
every symbol include/psp_b200.h declares, and refuses to compute without a GPU
(no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psp_[a-z0-9_]+)\s*\(", text)) - {"psp_callback"})


def test_header_symbols_are_exported(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in psp_b200.h but not exported"
    assert sorted(native.SYMBOLS) == syms


def test_library_contains_sm100a_code_and_tma(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # cp.async.bulk (TMA) feeding the SpMV


def test_no_cpu_fallback(native):
    if native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(native.PspError) as e:
        native.Matrix.csr(1, 1, [1.0], [0], [0, 1])
    assert e.value.status == -1      # PSP_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_oracle_is_not_imported_by_the_product():
    import pysparse_b200  # noqa: F401
    pkg = os.path.join(ROOT, "pysparse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "psp_oracle" not in src, f


def test_hot_kernels_keep_their_register_budget(native):
    """The persistent PCG kernel runs at 64 registers (4 CTAs per SM).  When ptxas had to spill
    ~200 bytes per thread into its loops it was 12 % slower (621 vs 550 us per iteration at
    256^3) -- slower than the per-phase path it replaces -- so a change that brings the spills
    back must fail here, not show up as a benchmark regression.  The SpMV kernels must not
    spill at all."""
    out = subprocess.run(["cuobjdump", "-res-usage", native.LIB_PATH], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
    persist = {k: v for k, v in usage.items() if "k_pcg_persist" in k}
    assert len(persist) == 2, "k_pcg_persist<0> / <1> not found in the library"
    for k, (reg, stack) in persist.items():
        if "ILi1E" in k:      # diagonal-run format: loads kept in flight from registers, no spills at all
            assert reg <= 64 and stack == 0, (k, reg, stack)
        else:                 # CSR row blocks through the TMA ring
            assert reg <= 64 and stack <= 32, (k, reg, stack)
    spmv = [v for k, v in usage.items() if "k_spmv_tma" in k or "k_spmv_runs" in k]
    assert len(spmv) >= 4
    for reg, stack in spmv:
        assert reg <= 64 and stack == 0, (reg, stack)

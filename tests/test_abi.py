"""CPU-only: librsgpu.so loads, exports every symbol include/rsgpu.h declares, the ctypes table covers all of them,
and the product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import ringsnark_b200 as rs
    from ringsnark_b200.capi import SIGNATURES
    lib = rs.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rsgpu.h but not exported by librsgpu.so"
        assert n in SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(SIGNATURES) == names


def test_no_cpu_fallback():
    import torch
    import ringsnark_b200 as rs
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rs.RsgError) as ei:
        rs.Context(128, [33550337], 256, [1073738753])
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_argument_validation_precedes_device_use():
    import ringsnark_b200 as rs
    lib = rs.load_library()
    h = ctypes.c_void_p()
    q = (ctypes.c_uint64 * 1)(33550337)
    Q = (ctypes.c_uint64 * 1)(1073738753)
    # N_E not a power of two / modulus not 1 mod 2N -> RSG_ERR_UNSUPPORTED / RSG_ERR_ARG, like std::invalid_argument upstream
    assert lib.rsg_context_create(ctypes.byref(h), 128, 1, q, 300, 1, Q, 0) == -5
    bad = (ctypes.c_uint64 * 1)(1073738755)
    assert lib.rsg_context_create(ctypes.byref(h), 128, 1, q, 256, 1, bad, 0) == -1
    assert b"prime" in lib.rsg_last_error()

"""ringsnark_b200 -- B200-native prover backend for zkFHE/ringSNARK.

The product is `librsgpu.so` (CUDA kernels for sm_100a behind the C ABI of include/rsgpu.h) plus the C++ backend
header under ringsnark_b200/cpp/ that gives the reference's templates a RingElem/EncodingElem pair over that ABI.
This Python package is only the thin ctypes mirror used by tests/ and bench.py.  There is no CPU fallback: importing
works anywhere, computing needs a CUDA device.
"""
from .capi import RsgError, lib_path, load_library  # noqa: F401
from .backend import (AUX_POLY, Context, Crs, Groth16ProvingKey, R1cs, RingVec, TERM_GENERAL, TERM_ONE,  # noqa: F401
                      TERM_SKIP)

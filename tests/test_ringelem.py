"""CPU tier: the drop-in backend's own RingElem (ringsnark_b200/cpp/ringsnark/seal_gpu/seal_ring.hpp: scalar / polynomial variant,
SEAL's modular routines restated, SealPoly::is_zero / is_equal byte-count quirks) against the reference's
ringsnark::seal::RingElem on every operator and operand pairing -- oracle/ringelem_check.cpp, built where /root/reference exists."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ringelem_host_arithmetic_matches_reference():
    exe = os.path.join(ROOT, "oracle", "_ref", "ringelem_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ringelem_check not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-1500:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["checks"] > 2000

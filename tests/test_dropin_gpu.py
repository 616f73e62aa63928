"""Drop-in boundary on the GPU: oracle/_ref/dropin_harness instantiates the UNMODIFIED reference templates
(groth16 / rinocchio generator, prover, verifier) over ringsnark::seal_gpu::{RingElem, EncodingElem}
(ringsnark_b200/cpp/ringsnark/seal_gpu/seal_ring.hpp -> librsgpu.so) and over the reference's own SEAL backend in one
process, on the same CRS / assignment / prover randomness; proofs must be word-identical and the reference verifier
must accept the GPU proof (constant-free circuits).  The binary is prebuilt where /root/reference exists and travels
to the GPU box."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "dropin_harness")


def run(case, seed, which="both", timeout=900, env=None):
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/dropin_harness not built (needs /root/reference at build time)")
    out = subprocess.run([HARNESS, case, str(seed), which], capture_output=True, text=True, timeout=timeout,
                         env=dict(os.environ, **(env or {})))
    assert out.stdout.strip(), out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    return out.returncode, res, out.stderr


@pytest.mark.parametrize("case", ["tiny_fast", "tiny_slow", "tiny_full", "tiny_quirks"])
def test_tiny_cases_both_systems(case):
    rc, res, err = run(case, 11)
    assert rc == 0 and res["ok"], (res, err[-1500:])
    assert res["groth16"]["bit_exact"] == [1, 1, 1]
    assert res["instance_map"]["bit_exact"] == 1      # generator / verifier ran the device instance map (8(f) rank 3)
    assert res["rinocchio"]["bit_exact"] == [1] * 9
    assert res["groth16"]["verified"] == res["groth16"]["verified_ref"]
    assert res["rinocchio"]["verified"] == res["rinocchio"]["verified_ref"]
    if case in ("tiny_fast", "tiny_slow"):     # constant-free circuits with noise budget to spare: accepted
        assert res["groth16"]["verified"] and res["rinocchio"]["verified"]


@pytest.mark.parametrize("case,which", [("c1", "both"), ("c3p", "groth16"), ("c4s", "both")])
def test_reference_sized_cases(case, which):
    rc, res, err = run(case, 23, which)
    assert rc == 0 and res["ok"], (res, err[-1500:])
    if which in ("groth16", "both"):
        assert res["groth16"]["bit_exact"] == [1, 1, 1] and res["groth16"]["verified"]
        assert res["instance_map"]["bit_exact"] == 1
    if which == "both":
        assert res["rinocchio"]["bit_exact"] == [1] * 9 and res["rinocchio"]["verified"]


@pytest.mark.parametrize("witness", ["dense", "fast"])
def test_medium_circuit_both_witness_paths(witness):
    """C4 parameters at n = 129 (the largest circuit the reference proves in seconds): the unmodified prover template over
    the GPU backend, with the dense and with the quasi-linear witness map, against the reference's proof on the same CRS."""
    rc, res, err = run("c4m", 31, "groth16", env={"RSG_WITNESS": witness})
    assert rc == 0 and res["ok"], (res, err[-1500:])
    assert res["groth16"]["bit_exact"] == [1, 1, 1] and res["instance_map"]["bit_exact"] == 1
    assert res["groth16"]["verified"] == res["groth16"]["verified_ref"]

"""CPU tier: the host-side polynomial routines behind the witness-map constants (csrc/host_poly.hpp: blocked transform
products, the vanishing polynomial by divide and conquer, rev(Z)^-1 by Newton iteration) against their naive forms --
tools/host_poly_check.cpp, compiled here with g++ (no CUDA)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_poly_routines(tmp_path):
    exe = str(tmp_path / "host_poly_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "host_poly_check.cpp")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stdout + res.stderr

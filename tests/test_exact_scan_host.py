"""The exact-cumsum arithmetic (bayesssm_b200/csrc/bssm_exact.cuh) compiled for the host and checked
against the sequential double cumsum of src/resampling.cpp:25 -- no GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exact_scan_logic_matches_sequential_sum(tmp_path):
    exe = tmp_path / "host_exact_scan"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_exact_scan.cpp")], check=True)
    r = subprocess.run([str(exe), "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches=0" in r.stdout

"""The C++ host layer (include/chrom.hpp, libchrom_host.so): the mirror of chrom-rs's Solver API a
compiled host links against.  CPU part: the reference's own unit tests of the boundary types,
restated in tests/cpp/test_host_api.cpp, and the exported C++ symbols."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(os.path.join(BIN, "test_host_api")):
        import __graft_entry__ as g
        g.build()
    return BIN


def test_host_api_unit_tests(built):
    r = subprocess.run([os.path.join(built, "test_host_api")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


def test_host_library_exports_solver_api():
    so = os.path.join(ROOT, "chromatography_b200", "libchrom_host.so")
    out = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    for sym in ("chrom::CudaSolverBase::solve(", "chrom::CudaSolverBase::solve_batch(", "chrom::LangmuirSingle::LangmuirSingle(",
                "chrom::LangmuirMulti::create(", "chrom::TemporalInjection::evaluate(", "chrom::DomainBoundaries::temporal(",
                "chrom::SolverConfiguration::time_evolution(", "chrom::validate_state_message", "chrom::rsf(",
                "chrom::sample_indices("):
        assert sym in out, sym


def test_solve_without_gpu_fails_loudly(built):
    """No CPU fallback: on a box without a GPU the C++ solver returns the backend's error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(built, "chrom_solve_dump"), "single", "rk4", "fast", "6", "100", "0", "/dev/null"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "CUDA backend unavailable (no CPU fallback)" in r.stderr

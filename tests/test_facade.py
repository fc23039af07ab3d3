"""The C++ drop-in facade (viennacl-dev_b200/include/viennacl/...) is exercised by plain-g++ programs that restate the
reference's own tests/tutorials (tests/src/sparse.cpp, self_assign.cpp, examples/tutorial/iterative*.cpp,
wrap-cuda-buffer.cu) -- see viennacl-dev_b200/facade_tests/.  CPU: they build without nvcc and refuse to run without a
device (no CPU fallback).  GPU: they pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACADE = os.path.join(ROOT, "viennacl-dev_b200", "lib", "facade")
PROGS = ["sparse_prod", "iterative", "wrap_cuda_buffer", "matrix_free", "matrix_market", "api_surface", "sparse_prod_float", "iterative_float"]


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "viennacl-dev_b200", "facade_tests")])


def test_facade_programs_build_with_plain_gxx(pkg):
    if not pkg.library_available():
        pkg.build_library()
    _build()
    for p in PROGS + ["bench_sparse_solver"]:
        assert os.access(os.path.join(FACADE, p), os.X_OK), p


def test_facade_refuses_to_run_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    _build()
    r = subprocess.run([os.path.join(FACADE, "wrap_cuda_buffer")], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("prog", PROGS)
def test_facade_program_passes_on_gpu(prog):
    exe = os.path.join(FACADE, prog)
    if not os.access(exe, os.X_OK):
        _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "COMPLETED SUCCESSFULLY" in r.stdout


@pytest.mark.gpu
def test_plain_c_example_passes_on_gpu():
    """examples/c_api_cg.c: the C-ABI used from C99 (gcc -std=c99 -pedantic), double and float entry points."""
    exe = os.path.join(FACADE, "c_api_cg")
    if not os.access(exe, os.X_OK):
        _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:]); print(r.stderr[-2000:])
    assert r.returncode == 0 and "C EXAMPLE COMPLETED SUCCESSFULLY" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_plain_c_example_builds_and_refuses_without_device(pkg):
    import torch
    if not pkg.library_available():
        pkg.build_library()
    _build()
    exe = os.path.join(FACADE, "c_api_cg")
    assert os.access(exe, os.X_OK)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_matrix_market_host_parsing(pkg):
    """viennacl/io/matrix_market.hpp is host code: general / symmetric / pattern headers, index bases, malformed input."""
    if not pkg.library_available():
        pkg.build_library()
    _build()
    r = subprocess.run([os.path.join(FACADE, "matrix_market"), "host-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "COMPLETED SUCCESSFULLY" in r.stdout, r.stdout + r.stderr

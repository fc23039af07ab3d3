"""CPU tests of host-side logic that otherwise only runs behind the GPU drivers: the host half of a pipelined GMRES restart cycle
(csrc/gmres_host.cuh, shared by the single-domain and the row-partitioned driver) against a numpy restatement of gmres.hpp:306-352."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libgmres_host_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(ROOT, "viennacl-dev_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host_shims", "gmres_host_shim.cpp"), "-o", out])
    L = C.CDLL(out)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.gmres_cycle_host_shim.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, ip, ip]
    L.gmres_cycle_host_shim.restype = C.c_int
    return L


def reference_cycle(k, R, xi, tol, rho, rho_0, norm_rhs, per_iteration_stop):
    """gmres.hpp:306-352 (+ :579-584 for the per-iteration stop of the preconditioned path), restated.  R: k x k, column-major, R[i + j*k]."""
    kk = k
    for i in range(k):
        if abs(R[i + i * k]) < tol * R[0]:
            kk = i
            break
    iters, conv = 0, False
    for i in range(kk):
        iters += 1
        if xi[i] >= rho or xi[i] <= -rho:
            kk = i
            break
        rho *= math.sin(math.acos(xi[i] / rho))
        if per_iteration_stop and abs(rho * rho_0 / norm_rhs) < tol:
            kk = i + 1; conv = True
            break
    eta = list(xi)
    for i in range(kk - 1, -1, -1):
        for j in range(i + 1, kk):
            eta[i] -= R[i + j * k] * eta[j]
        eta[i] /= R[i + i * k]
    coef = [rho_0 * eta[i] for i in range(kk)] + [0.0] * (k - kk)
    return kk, coef, rho, iters, conv


@pytest.mark.parametrize("case", ["regular", "loss_of_orthogonality", "xi_exceeds_rho", "per_iteration_stop"])
def test_gmres_cycle_host(shim, case):
    rng = np.random.default_rng(11)
    k = 12
    R = np.zeros(k * k)
    for j in range(k):
        for i in range(j + 1):
            R[i + j * k] = rng.uniform(0.5, 1.5) if i == j else rng.uniform(-0.3, 0.3)
    xi = rng.uniform(-0.2, 0.2, k)
    tol, rho, rho_0, norm_rhs, stop = 1e-8, 1.0, 3.5, 7.0, False
    if case == "loss_of_orthogonality":
        R[5 + 5 * k] = 1e-12                     # |R_55| < tol * R_00: Krylov space truncated to 5
    elif case == "xi_exceeds_rho":
        xi[4] = 2.0                              # |xi_4| >= rho: truncated to 4 (gmres.hpp:324)
    elif case == "per_iteration_stop":
        xi = np.full(k, 0.9999999999); xi[0] = 0.9999999999
        stop, tol = True, 1e-3
    exp_kk, exp_coef, exp_rho, exp_it, exp_conv = reference_cycle(k, R, list(xi), tol, rho, rho_0, norm_rhs, stop)
    coef = np.zeros(k); rho_io = C.c_double(rho); it, cv = C.c_int(0), C.c_int(0)
    kk = shim.gmres_cycle_host_shim(k, R.ctypes.data_as(C.POINTER(C.c_double)), xi.ctypes.data_as(C.POINTER(C.c_double)), tol, rho_0, norm_rhs, int(stop),
                                    C.byref(rho_io), coef.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it), C.byref(cv))
    assert kk == exp_kk and it.value == exp_it and bool(cv.value) == exp_conv
    assert rho_io.value == exp_rho
    assert np.array_equal(coef[:kk], np.array(exp_coef[:kk]))
    if case == "loss_of_orthogonality":
        assert kk == 5
    if case == "xi_exceeds_rho":
        assert kk == 4
    if case == "per_iteration_stop":
        assert cv.value == 1 and kk < k

"""CPU tests of the boundary: the C-ABI library loads, exports every symbol
the header declares, fails loudly without a GPU, and its host-side math
(interpolation, Magnus moments, Chebyshev coefficients) is right."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy.interpolate import make_interp_spline
from scipy.special import jv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "pulser_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    raw = C.CDLL(os.path.join(ROOT, "pulser_b200", "libpulser_b200.so"))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    from pulser_b200 import _lib

    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.pb200_version() >= 100


def test_no_gpu_means_loud_failure(lib):
    """The product path has no CPU fallback."""
    from pulser_b200 import engine, workloads as W
    from pulser_b200._lib import PB200Error

    if engine.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(PB200Error, match="no CUDA device"):
        engine.DevicePlan(W.config_c1())


def test_bad_arguments_rejected(lib):
    from pulser_b200._lib import PlanDesc

    h = C.c_void_p()
    d = PlanDesc()
    d.n_qudits = 0
    assert lib.pb200_plan_create(C.byref(h), C.byref(d)) == -1
    assert b"n_qudits" in lib.pb200_last_error()
    assert lib.pb200_plan_destroy(None) == 0


@pytest.mark.parametrize("n,order", [(4, 3), (5, 3), (64, 3), (64, 1), (4001, 3)])
def test_interpolant_equals_scipy(lib, n, order):
    rng = np.random.default_rng(n)
    x = np.arange(n) / 1000.0
    y = rng.normal(size=n) + 1j * rng.normal(size=n)
    tq = np.ascontiguousarray(np.concatenate([rng.uniform(x[0], x[-1], 200), x[:3], x[-3:]]))
    out = np.zeros(2 * len(tq))
    assert lib.pb200_host_interpolate(P(x), P(y.view(np.float64)), n, order, P(tq), len(tq), P(out)) == 0
    ref = make_interp_spline(x, y, k=order)(tq)
    assert np.max(np.abs(out.view(np.complex128) - ref)) < 1e-12


def test_interpolant_nonuniform_grid(lib):
    """sampling_rate < 1 gives a non-uniform grid (hamiltonian.py:87-95)."""
    full = np.arange(1001) / 1000.0
    idx = np.linspace(0, 1000, int(0.37 * 1001), dtype=int)
    x = np.ascontiguousarray(full[idx])
    y = np.sin(7 * x) * np.exp(1j * 3 * x)
    tq = np.ascontiguousarray(np.linspace(0, 1, 333))
    out = np.zeros(2 * len(tq))
    assert lib.pb200_host_interpolate(P(x), P(y.view(np.float64)), len(x), 3, P(tq), len(tq), P(out)) == 0
    assert np.max(np.abs(out.view(np.complex128) - make_interp_spline(x, y, k=3)(tq))) < 1e-11


def test_order0_is_previous_value(lib):
    x = np.arange(6, dtype=float)
    y = (np.arange(6) ** 2).astype(complex)
    tq = np.array([0.0, 0.5, 1.0, 1.999, 4.2, 5.0])
    out = np.zeros(12)
    assert lib.pb200_host_interpolate(P(x), P(y.view(np.float64)), 6, 0, P(tq), 6, P(out)) == 0
    np.testing.assert_allclose(out[::2], [0, 0, 1, 1, 16, 25])


@pytest.mark.parametrize("a,b", [(0.003, 0.007), (0.0031, 0.0124), (0.0, 0.063), (0.0105, 0.0107)])
def test_magnus_moments_exact(lib, a, b):
    rng = np.random.default_rng(3)
    n = 64
    x = np.arange(n) / 1000.0
    y = rng.normal(size=n) + 1j * rng.normal(size=n)
    out = np.zeros(4)
    assert lib.pb200_host_moments(P(x), P(y.view(np.float64)), n, 3, a, b, P(out)) == 0
    sp = make_interp_spline(x, y, k=3)
    A1, A2 = sp.antiderivative(1), sp.antiderivative(2)
    tm, h = 0.5 * (a + b), b - a
    B0 = A1(b) - A1(a)
    B1 = ((b - tm) * A1(b) - (a - tm) * A1(a) - (A2(b) - A2(a))) / h
    assert abs(complex(out[0], out[1]) - B0) < 1e-15
    assert abs(complex(out[2], out[3]) - B1) < 1e-16


@pytest.mark.parametrize("rho", [1e-6, 0.3, 2.0, 17.0, 120.0])
def test_chebyshev_coefficients(lib, rho):
    out = np.zeros(2000)
    cnt = C.c_int32()
    assert lib.pb200_host_chebyshev(rho, 1e-13, P(out), 1000, C.byref(cnt)) == 0
    a = out[: 2 * cnt.value].view(np.complex128)
    ref = np.array([(2 if j else 1) * (-1j) ** j * jv(j, rho) for j in range(cnt.value)])
    assert np.max(np.abs(a - ref)) < 1e-13
    xs = np.linspace(-1, 1, 41)
    assert np.max(np.abs(np.polynomial.chebyshev.chebval(xs, a) - np.exp(-1j * rho * xs))) < 1e-12

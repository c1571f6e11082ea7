"""CPU tests of the host side of the time-dependent Taylor propagator (integrator 3): the polynomial fit of the
QobjEvo spline on a step and the a-priori Taylor order, through the C ABI, and -- with exactly these two pieces --
the recurrence the CUDA stage kernel executes, restated in numpy on a 6-atom C2-shaped sequence and compared with the
DOP853 oracle (the solver call it replaces: pulser_simulation/simulation.py:729-735)."""
import ctypes as C

import numpy as np
import pytest
from scipy.interpolate import make_interp_spline

from pulser_b200 import workloads as W

dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


def fit(lib, x, y, a, h, p, order=3):
    c = np.zeros(p + 1)
    r = C.c_double()
    assert lib.pb200_host_taylor_fit(P(x), P(np.ascontiguousarray(y, dtype=float)), len(x), order, a, h, p, P(c), C.byref(r)) == 0
    return c, r.value


def test_fit_reproduces_polynomial_data(lib):
    """samples of a cubic: the not-a-knot spline IS that cubic, and the fit returns its shifted coefficients"""
    x = np.arange(200) / 1000.0
    y = 3.0 - 40.0 * x + 900.0 * x**2 - 5000.0 * x**3
    a, h = 0.0231, 0.0517
    c, r = fit(lib, x, y, a, h, 3)
    u = np.linspace(0, 1, 33)
    t = a + h * u
    exact = 3.0 - 40.0 * t + 900.0 * t**2 - 5000.0 * t**3
    assert np.max(np.abs(np.polynomial.polynomial.polyval(u, c) - exact)) < 1e-12
    assert r < 1e-12
    # a linear ramp is matched by degree 1 to rounding (what makes C2 / C5 cost one history term per order)
    c1, r1 = fit(lib, x, 5.0 - 120.0 * x, 0.05, 0.08, 1)
    assert r1 < 1e-13 and abs(c1[1] + 120.0 * 0.08) < 1e-12


@pytest.mark.parametrize("p", [2, 5, 8])
def test_fit_residual_is_honest(lib, p):
    """smooth curved samples: the reported residual bounds the true distance to the spline, and it falls with p"""
    x = np.arange(400) / 1000.0
    y = 7.0 * np.sin(9.0 * x) ** 2
    spl = make_interp_spline(x, y, k=3)
    a, h = 0.1003, 0.060
    c, r = fit(lib, x, y, a, h, p)
    u = np.linspace(0, 1, 2001)
    true = np.max(np.abs(np.polynomial.polynomial.polyval(u, c) - spl(a + h * u)))
    assert true <= 1.5 * r + 1e-15
    # degree 8 sits on the floor set by the spline itself: it deviates from any smooth function by ~ h^4 f''''/384
    assert r < {2: 5e-2, 5: 1e-5, 8: 3e-9}[p]


def test_order_matches_exponential_series(lib):
    """constant generator |H| <= m: the majorant is exp(h m) and K is where its Taylor tail drops below tol"""
    from math import factorial

    for rho in (0.3, 3.0, 10.0):
        k = C.c_int32()
        tail = C.c_double()
        m = np.array([rho])
        assert lib.pb200_host_taylor_order(1.0, P(m), 0, 1e-12, C.byref(k), C.byref(tail)) == 0
        K = k.value
        t_K = sum(rho**j / factorial(j) for j in range(K + 1, K + 120))
        t_Km1 = t_K + rho**K / factorial(K)
        assert t_K <= 1e-12 < t_Km1
        assert abs(tail.value - t_K) < 1e-3 * t_K + 1e-30


def test_recurrence_with_library_host_math_vs_oracle(lib):
    """(k+1) chi_{k+1} = -i h sum_j H_j chi_{k-j} with H_j from pb200_host_taylor_fit and K from
    pb200_host_taylor_order, in numpy (the arithmetic of stage_d2_taylor_kernel), against the DOP853 oracle."""
    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    n = 6
    spec = W.config_c2(n=n, seed=20)
    t = np.asarray(spec.sampling_times)
    om_s = np.real(spec.drives[0].coef[0]).copy()
    th_s = spec.drives[0].det[0].copy()
    U = spec.pair_matrix()
    D = 1 << n
    idx = np.arange(D)
    bits = np.array([(idx >> (n - 1 - k)) & 1 for k in range(n)])
    nr = (bits == 0).astype(float)           # eigenbasis ["r", "g"]: digit 0 = r = the `from` state
    cnt = nr.sum(axis=0)
    dint = sum(U[i, j] * nr[i] * nr[j] for i in range(n) for j in range(i + 1, n))

    def xtot(v):
        vt = v.reshape([2] * n)
        return sum(np.flip(vt, axis=k) for k in range(n)).reshape(-1)

    psi = evolve.all_ground_state(spec).astype(complex)
    tol = 1e-8
    rate = tol / t[-1]
    a_i, applies = 0, 0
    nt = len(t)
    while a_i < nt - 1:
        L = min(80, nt - 1 - a_i)
        while True:   # longest step both splines fit with degree <= 8 (degree 3 on a single interval is exact)
            a, h = t[a_i], t[a_i + L] - t[a_i]
            allow = 0.15 * rate / n
            fo = ft = None
            for p in range(9):
                c, r = fit(lib, t, om_s, a, h, p)
                if r <= allow or (L == 1 and p >= 3):
                    fo = c
                    break
            for p in range(9):
                c, r = fit(lib, t, th_s, a, h, p)
                if r <= allow or (L == 1 and p >= 3):
                    ft = c
                    break
            if fo is not None and ft is not None:
                break
            L = max(1, L // 2)
        p = max(len(fo), len(ft)) - 1
        co = np.zeros(p + 1); co[: len(fo)] = fo
        ct = np.zeros(p + 1); ct[: len(ft)] = ft
        d0 = dint - ct[0] * cnt
        lo, hi = d0.min() - abs(co[0]) * n, d0.max() + abs(co[0]) * n
        gam = np.array([0.5 * (lo + hi)] + [-ct[j] * n / 2 for j in range(1, p + 1)])
        mj = np.array([0.5 * (hi - lo)] + [abs(ct[j]) * n / 2 + abs(co[j]) * n for j in range(1, p + 1)])
        k_out, tail = C.c_int32(), C.c_double()
        assert lib.pb200_host_taylor_order(h, P(mj), p, max(1e-15, 0.1 * rate * h), C.byref(k_out), C.byref(tail)) == 0
        K = k_out.value
        chis, Gs, acc = [psi], [], psi.copy()
        for k in range(K):
            Gs.append(xtot(chis[k]))
            new = (d0 - gam[0]) * chis[k] + co[0] * Gs[k]
            for j in range(1, min(p, k) + 1):
                new += (-ct[j] * cnt - gam[j]) * chis[k - j] + co[j] * Gs[k - j]
            chis.append(-1j * h / (k + 1) * new)
            acc += chis[-1]
        psi = acc * np.exp(-1j * h * sum(gam[j] / (j + 1) for j in range(p + 1)))
        applies += K
        a_i += L
    ref = evolve.sesolve(OracleHamiltonian.from_spec(spec), evolve.all_ground_state(spec), [0.0, t[-1]],
                         rtol=1e-13, atol=1e-15)[-1]
    assert np.max(np.abs(psi - ref)) < 1e-8
    assert applies < 1.5 * (nt - 1)      # ~0.6 H-applies per ns at this size (the Magnus path needs ~4)


# ---------------------------------------------------------------------------------------------------------------------
# separable structure of noise-trajectory batches (what lets C4 batches run on integrator 3)
def _separable(lib, coef, det):
    B, N, nt = det.shape
    ok = C.c_int32(-1)
    a = np.zeros((B, N), dtype=np.complex128)
    c = np.zeros((B, N))
    m = np.zeros(nt)
    coef = np.ascontiguousarray(coef, dtype=np.complex128)
    det = np.ascontiguousarray(det, dtype=np.float64)
    assert lib.pb200_host_taylor_separable(P(coef.view(np.float64)), P(det), B, N, nt, C.byref(ok), P(a.view(np.float64)),
                                           P(c), P(m)) == 0
    return ok.value, a, c, m


def test_c4_batches_are_separable(lib):
    """every device batch of the striped C4 run (bench.py: 1024 trajectories over 8 ranks, batches of 64) has the form
    coef = a_k x one row, det = det_00 + c_k x slot mask; a = amplitude fluctuation x waist factor, c = doppler shifts.
    (A plain double sum in the least-squares fit used to reject 3 batches in 8 on rounding.)"""
    from pulser_b200 import parallel

    for rank in (0, 5):
        mine = set(parallel.stripe(1024, rank, 8))
        chunk = []
        for _, spec in W.config_c4_stream(1024, keep=mine):
            chunk.append(spec)
            if len(chunk) == 64:
                coef = np.array([s.drives[0].coef for s in chunk])
                det = np.array([s.drives[0].det for s in chunk])
                ok, a, c, m = _separable(lib, coef, det)
                assert ok == 1
                # reconstruct the tables from the factors
                b, k, _ = np.unravel_index(np.argmax(np.abs(coef)), coef.shape)
                assert np.max(np.abs(coef - a[:, :, None] * coef[b, k][None, None, :])) < 1e-12
                assert np.max(np.abs(det - det[0, 0][None, None, :] - c[:, :, None] * m[None, None, :])) < 1e-11
                assert np.max(np.abs(np.abs(m[:-1]) - 1.0)) < 1e-12 and m[-1] == 0.0   # slot mask: 0 on the padded sample
                assert np.all(np.abs(a.imag) < 1e-15) and 0.6 < a.real.min() and a.real.max() == 1.0
                chunk = []


def test_non_separable_tables_are_refused(lib):
    """per-qubit time shapes (a Local pulse on one atom) or moving phases are not of that form"""
    specs = W.config_c4(3)
    coef = np.array([s.drives[0].coef for s in specs])
    det = np.array([s.drives[0].det for s in specs])
    assert _separable(lib, coef, det)[0] == 1
    bad = coef.copy()
    bad[1, 4, 1000:1500] *= 1.0 + 1e-9            # one qubit's amplitude changes shape
    assert _separable(lib, bad, det)[0] == 0
    bad = det.copy()
    bad[2, 7, 2000:] += 1e-6                       # a second detuning shape
    assert _separable(lib, coef, bad)[0] == 0
    # a uniform single trajectory is trivially separable: a = 1, c = 0
    spec = W.config_c2(n=6)
    ok, a, c, m = _separable(lib, spec.drives[0].coef[None], spec.drives[0].det[None])
    assert ok == 1 and np.all(a == 1.0) and np.all(c == 0.0)

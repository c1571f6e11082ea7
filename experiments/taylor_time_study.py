"""Experiment (CPU, numpy): time-dependent Taylor propagator for Pulser-shaped Hamiltonians.

H(t) = Dint - delta(t) N + omega(t) Xtot   (global drive, constant phase), delta / omega piecewise cubic
(the not-a-knot spline QuTiP's QobjEvo builds, hamiltonian.py:436).  On a step [a, a+h] the two coefficient
functions are replaced by polynomials of degree p in u = (t-a)/h (Chebyshev interpolation of the spline; the
residual is measured and enters the error budget), H(u) = sum_j H_j u^j, and the solution is the Taylor series
psi(u) = sum_k chi_k u^k with   (k+1) chi_{k+1} = -i h sum_{j<=min(p,k)} H_j chi_{k-j}:
ONE gather (Xtot chi_k) per order, the history terms are own-element linear combinations of stored
chi_{k-j}, G_{k-j} = Xtot chi_{k-j}.  No Magnus commutator error at all: the step is bounded by the spectral
half-width (rho = W h <~ 10, fp64 cancellation) and by the polynomial fit only.

Usage: python experiments/taylor_time_study.py [n] [rho_target] [pmax]
"""
from __future__ import annotations

import sys
import time

import numpy as np
from scipy.interpolate import CubicSpline

sys.path.insert(0, ".")
from pulser_b200 import workloads  # noqa: E402


def build(n, seed=None):
    spec = workloads.config_c2(n, seed)
    t = np.asarray(spec.sampling_times)
    drv = spec.drives[0]
    om = CubicSpline(t, np.real(drv.coef[0]))  # = Omega/2 (phase 0)
    de = CubicSpline(t, drv.det[0])
    U = spec.pair_matrix()
    D = 1 << n
    idx = np.arange(D)
    bits = np.array([(idx >> (n - 1 - k)) & 1 for k in range(n)])
    # eigenbasis ["r","g"]: digit 0 = r
    nr = (bits == 0).astype(float)
    dint = np.zeros(D)
    for i in range(n):
        for j in range(i + 1, n):
            dint += U[i, j] * nr[i] * nr[j]
    cnt = nr.sum(axis=0)
    return spec, t, om, de, dint, cnt


def xtot(psi, n):
    pt = psi.reshape([2] * n)
    out = np.zeros_like(pt)
    for k in range(n):
        out += np.flip(pt, axis=k)
    return out.reshape(-1)


def poly_fit(fn, a, h, p):
    """monomial coefficients in u in [0,1] of the degree-p Chebyshev interpolant of fn on [a, a+h]; residual."""
    k = np.arange(p + 1)
    x = np.cos(np.pi * (2 * k + 1) / (2 * (p + 1)))  # nodes in [-1,1]
    u = 0.5 * (x + 1.0)
    V = np.vander(u, p + 1, increasing=True)
    c = np.linalg.solve(V, fn(a + h * u))
    ug = np.linspace(0, 1, max(64, int(h * 1000 * 16)))
    res = np.max(np.abs(np.polynomial.polynomial.polyval(ug, c) - fn(a + h * ug)))
    return c, res


def taylor_step(psi, a, h, om, de, dint, cnt, n, p_om, p_de, tol, kmax=200):
    """one step; returns new psi, number of gathers."""
    co, _ = poly_fit(om, a, h, p_om)
    cd, _ = poly_fit(de, a, h, p_de)
    # centre: H0 spectrum bounds
    d0 = dint - cd[0] * cnt
    lo = d0.min() - abs(co[0]) * n
    hi = d0.max() + abs(co[0]) * n
    c0 = 0.5 * (lo + hi)
    # higher-order centres: -cd[j] * n/2
    cen = np.zeros(max(len(co), len(cd)))
    cen[0] = c0
    for j in range(1, len(cd)):
        cen[j] = -cd[j] * n / 2
    chis = [psi]
    Gs = []
    acc = psi.copy()
    small = 0
    k = 0
    while k < kmax:
        G = xtot(chis[k], n)
        Gs.append(G)
        new = (d0 - cen[0]) * chis[k] + co[0] * G
        for j in range(1, min(max(p_om, p_de), k) + 1):
            if j <= p_de:
                new += (-cd[j] * cnt - cen[j]) * chis[k - j]
            if j <= p_om:
                new += co[j] * Gs[k - j]
        new *= -1j * h / (k + 1)
        chis.append(new)
        acc += new
        k += 1
        nn = np.linalg.norm(new)
        if nn < tol:
            small += 1
            if small >= 2:
                break
        else:
            small = 0
    # phase of the centre: exp(-i h int_0^1 sum_j cen_j u^j du)
    ph = sum(cen[j] / (j + 1) for j in range(len(cen)))
    return acc * np.exp(-1j * h * ph), k, 0.5 * (hi - lo)


def reference(psi0, t, om, de, dint, cnt, n, t_end, sub=4):
    """tight reference: CF4 with exact-ish moments via Gauss quadrature on each 1/sub-ns piece, expm via Taylor."""
    from scipy.integrate import solve_ivp

    def rhs(tt, y):
        return -1j * ((dint - de(tt) * cnt) * y + om(tt) * xtot(y, n))

    sol = solve_ivp(rhs, (0.0, t_end), psi0, method="DOP853", rtol=1e-12, atol=1e-14, max_step=2e-3)
    return sol.y[:, -1], sol.nfev


def kinks(t, fns, tol=1e-6):
    """knot indices where the samples' third difference is not ~0 (non-smooth)."""
    ks = set()
    for fn in fns:
        y = fn(t)
        sc = np.max(np.abs(y)) + 1e-300
        d3 = y[3:] - 3 * y[2:-1] + 3 * y[1:-2] - y[:-3]
        for i in np.nonzero(np.abs(d3) > tol * sc)[0]:
            ks.update([i + 1, i + 2])
    return sorted(ks)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rho_t = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    pmax = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    spec, t, om, de, dint, cnt = build(n)
    D = 1 << n
    psi0 = np.zeros(D, complex)
    psi0[D - 1] = 1.0  # all ground (digit 1 = g)
    T = t[-1]
    tol_total = 1e-8
    fit_budget = 0.2 * tol_total / T  # per unit time
    # half-width per unit time (max over the sequence)
    W = 0.5 * ((dint - de(t)[:, None] * cnt).max(axis=1) - (dint - de(t)[:, None] * cnt).min(axis=1)) + np.abs(om(t)) * n
    print("n", n, "half-width range rad/us", W.min(), W.max())
    kk = kinks(t, [om, de])
    print("non-smooth knots:", kk[:12], "...", len(kk))
    psi = psi0.copy()
    a_i = 0
    nt = len(t)
    gathers = 0
    steps = 0
    log = []
    t0 = time.time()
    while a_i < nt - 1:
        # longest step with rho <= rho_t whose fit residual passes, with degree <= pmax
        Wi = W[a_i]
        L = max(1, int(rho_t / (Wi * 1e-3)))
        L = min(L, nt - 1 - a_i)
        while True:
            a = t[a_i]
            h = t[a_i + L] - a
            ok = False
            for p in range(1, pmax + 1):
                _, r1 = poly_fit(om, a, h, p)
                _, r2 = poly_fit(de, a, h, p)
                if (r1 * n + r2 * n / 2) <= fit_budget:
                    ok = True
                    break
            if ok or L == 1:
                if not ok:
                    p = 3
                break
            L = max(1, L // 2)
        psi, k, hw = taylor_step(psi, a, h, om, de, dint, cnt, n, p, p, 0.05 * tol_total * h / T)
        gathers += k
        steps += 1
        log.append((a_i, L, p, k))
        a_i += L
    el = time.time() - t0
    ref, nfev = reference(psi0, t, om, de, dint, cnt, n, T)
    err = np.linalg.norm(psi - ref)
    print(f"steps {steps} gathers {gathers} per ns {gathers / (nt - 1):.3f} err {err:.3e} norm-1 {abs(np.vdot(psi, psi) - 1):.2e} (ref nfev {nfev}) {el:.1f}s")
    short = [x for x in log if x[1] < 8]
    print("short steps:", len(short), "gathers in short steps", sum(x[3] for x in short))
    print("sample of log:", log[:6], log[len(log) // 2: len(log) // 2 + 3])


if __name__ == "__main__":
    main()

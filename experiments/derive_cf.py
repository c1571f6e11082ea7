"""Derivation of commutator-free Magnus schemes in the truncated free algebra.

A(x) = a1 P0(x) + a2 P1(x) + a3 P2(x) (+ a4 P3 ...), x in [-1, 1] (Legendre), generator a_n has weight n
(for a smooth A the n-th Legendre coefficient is O(h^n)).  The exact flow U = Texp(int A) and any product of
exponentials prod_i exp(sum_n f[i][n] a_n) are elements of the tensor algebra truncated at weight W; the scheme has
order p iff log(prod) - log(U) vanishes up to weight p.  No reference to a published coefficient table: the
coefficients are solved for here and checked by the convergence order of the resulting propagator
(experiments/magnus_step_study.py, tests/test_cf_schemes_cpu.py).

Conventions used by the library (csrc/spline.hpp):  with t in [a, b], h = b - a, x = 2 (t - mid) / h:
    a1 = int A dt                       (= h * mean)
    a2 = 3 * int x A dt     = 6 * B1    (B1 = (1/h) int (t - mid) A dt)
    a3 = 5 * int P2(x) A dt
so that A(t) = (1/h) (a1 P0 + a2 P1 + a3 P2 + ...).
"""
from __future__ import annotations

import itertools
import sys
from fractions import Fraction

import numpy as np
from numpy.polynomial import legendre as L
from numpy.polynomial import polynomial as Pn


# ----------------------------------------------------------------- truncated tensor algebra
class Alg:
    """Elements: dict word(tuple of generator indices 1..G) -> float; weight(word) = sum of indices."""

    def __init__(self, W: int, G: int = 3):
        self.W, self.G = W, G
        self.words = [()]
        frontier = [()]
        while frontier:
            nxt = []
            for w in frontier:
                for g in range(1, G + 1):
                    if sum(w) + g <= W:
                        nxt.append(w + (g,))
            self.words += nxt
            frontier = nxt
        self.index = {w: i for i, w in enumerate(self.words)}
        self.n = len(self.words)
        # multiplication table: pairs (i, j, k) with words[i] + words[j] = words[k]
        ii, jj, kk = [], [], []
        for i, wi in enumerate(self.words):
            for j, wj in enumerate(self.words):
                if sum(wi) + sum(wj) <= W:
                    ii.append(i); jj.append(j); kk.append(self.index[wi + wj])
        self.ii, self.jj, self.kk = np.array(ii), np.array(jj), np.array(kk)
        self.weight = np.array([sum(w) for w in self.words])

    def mul(self, x, y):
        out = np.zeros(self.n, dtype=np.result_type(x, y))
        np.add.at(out, self.kk, x[self.ii] * y[self.jj])
        return out

    def one(self):
        e = np.zeros(self.n); e[0] = 1.0
        return e

    def gen(self, coeffs):
        """sum_n coeffs[n-1] a_n"""
        e = np.zeros(self.n)
        for n, c in enumerate(coeffs, start=1):
            if n <= self.G and n <= self.W:
                e[self.index[(n,)]] = c
        return e

    def exp(self, x):
        out = self.one(); term = self.one()
        for k in range(1, self.W + 1):
            term = self.mul(term, x) / k
            out = out + term
        return out

    def log(self, u):
        x = u - self.one()
        out = np.zeros(self.n); term = self.one()
        for k in range(1, self.W + 1):
            term = self.mul(term, x)
            out = out + ((-1) ** (k + 1)) * term / k
        return out


def exact_flow(alg: Alg):
    """U(1) for U' = A(x(t)) U on t in [0, 1], x = 2t - 1, A = sum_n a_n P_{n-1}(x): Picard iteration with
    polynomial-in-t coefficients per word (exact up to the truncation weight)."""
    G = alg.G
    # coefficient polynomials (in t) of each generator: P_{n-1}(2t - 1)
    gen_poly = []
    for n in range(1, G + 1):
        c = np.zeros(n); c[n - 1] = 1.0
        p = L.leg2poly(c)              # polynomial in x
        # substitute x = 2t - 1
        q = np.zeros(1)
        for k, ck in enumerate(p):
            q = Pn.polyadd(q, ck * Pn.polypow(np.array([-1.0, 2.0]), k))
        gen_poly.append(q)
    # U(t) = sum_w c_w(t) w ; c_() = 1 ; c_{g w}(t) = int_0^t p_g(s) c_w(s) ds   (A U: generator prepended)
    coef = {(): np.array([1.0])}
    for w in alg.words[1:]:
        g, rest = w[0], w[1:]
        integrand = Pn.polymul(gen_poly[g - 1], coef[rest])
        coef[w] = Pn.polyint(integrand)
    u = np.zeros(alg.n)
    for w, p in coef.items():
        u[alg.index[w]] = Pn.polyval(1.0, p)
    return u


def scheme_product(alg: Alg, F: np.ndarray):
    """prod over rows i = s-1 .. 0 ordering:  U = exp(row s-1) ... exp(row 0)  (row 0 acts first)."""
    u = alg.one()
    for row in F:
        u = alg.mul(alg.exp(alg.gen(row)), u)
    return u


def symmetric_rows(params: np.ndarray, s: int, G: int = 3):
    """Time-symmetric scheme: row s-1-i = row i with the sign (-1)^(n+1) on a_n."""
    half = s // 2
    rows = np.zeros((s, G))
    p = params.reshape(-1)
    k = 0
    for i in range(half):
        rows[i] = p[k:k + G]; k += G
        rows[s - 1 - i] = rows[i] * np.array([(-1) ** n for n in range(G)])
    if s % 2:
        mid = np.zeros(G)
        for n in range(0, G, 2):      # only a1, a3 survive in the middle row
            mid[n] = p[k]; k += 1
        rows[half] = mid
    return rows


def n_params(s: int, G: int = 3):
    return (s // 2) * G + ((G + 1) // 2 if s % 2 else 0)


def residual(alg: Alg, target_log, F, order):
    d = alg.log(scheme_product(alg, F)) - target_log
    return d[(alg.weight <= order) & (alg.weight > 0)]


def solve(s: int, order: int, G: int = 3, tries: int = 200, seed: int = 0, err_weight: int | None = None):
    from scipy.optimize import least_squares

    alg = Alg(order, G)
    tgt = alg.log(exact_flow(alg))
    algE = Alg(order + 1, G)
    tgtE = algE.log(exact_flow(algE))
    rng = np.random.default_rng(seed)
    sols = []
    for _ in range(tries):
        x0 = rng.normal(0, 0.4, size=n_params(s, G))
        fun = lambda x: residual(alg, tgt, symmetric_rows(x, s, G), order)
        r = least_squares(fun, x0, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
        if np.max(np.abs(r.fun)) < 1e-12:
            F = symmetric_rows(r.x, s, G)
            e = residual(algE, tgtE, F, order + 1)
            errn = float(np.linalg.norm(e))
            cover = float(np.sum(np.abs(F[:, 0])))
            key = tuple(np.round(F.reshape(-1), 6))
            if not any(k == key for k, *_ in sols):
                sols.append((key, F, errn, cover))
    sols.sort(key=lambda t: t[2])
    return sols


if __name__ == "__main__":
    s = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    G = int(sys.argv[3]) if len(sys.argv) > 3 else (2 if order <= 4 else 3)
    sols = solve(s, order, G, tries=int(sys.argv[4]) if len(sys.argv) > 4 else 100)
    print(f"{len(sols)} distinct solutions for s={s} order={order} G={G}")
    np.set_printoptions(precision=16, linewidth=200)
    for key, F, errn, cover in sols[:8]:
        print(f"-- leading error norm {errn:.3e}  sum|f_i1| {cover:.4f}")
        print(F)

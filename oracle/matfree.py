"""ORACLE (test infrastructure, never on the product path).

Matrix-free numpy restatement of ``H(t) @ psi`` for sizes where the
Kronecker-product assembly of ``oracle/ref_hamiltonian.py`` (the literal
restatement of ``pulser_simulation/hamiltonian.py:145-200, 246-439``) is too
slow.  ``tests/test_oracle_cpu.py`` checks it against that literal
restatement on small systems; it follows SURVEY.md Appendix A.3:

    (H psi)[s] = diag(s) psi[s] + sum_k c_k^(+/-) psi[s with digit_k swapped]
    diag(s) = - sum_k det_k [s_k = from] + sum_{i<j} U_ij [s_i = r][s_j = r]
"""
from __future__ import annotations

from typing import Any

import numpy as np
from scipy.interpolate import make_interp_spline

# (to, from) eigenstates of the drive operator c(t)|to><from| + h.c. and of the detuning -det(t)|from><from| per
# addressed basis, restated here from the reference (hamiltonian.py:340-352: "ground-rydberg" -> sigma_gr with
# the detuning on |r>, "digital" -> sigma_hg with the detuning on |g>, "XY" -> sigma_ud with the detuning on |d>)
# so that the checker does not borrow the table of the code it checks (tests/test_oracle_cpu.py compares the two).
BASIS_ROLES = {
    "ground-rydberg": ("g", "r"),
    "digital": ("h", "g"),
    "XY": ("u", "d"),
}


class MatFreeHamiltonian:
    def __init__(self, spec: Any, order: int = 3) -> None:
        self.spec = spec
        n, d = spec.n_qudits, spec.dim
        self.n, self.d = n, d
        D = d**n
        idx = np.arange(D)
        # digits[k] = digit of qudit k (qudit 0 most significant)
        self.digits = np.array(
            [(idx // d ** (n - 1 - k)) % d for k in range(n)], dtype=np.int8
        )
        self.dint = np.zeros(D)
        self.xy = None
        if spec.has_interaction() and spec.interaction_type == "XY":
            self.xy = (spec.xy_matrix(), spec.eigenbasis.index("u"), spec.eigenbasis.index("d"))
        # XY + SLM mask (hamiltonian.py:399-424): pairs touching a masked qudit carry the interpolated 0/1
        # coefficient of the unmasked term; the other pairs have weight c + (1 - c) = 1
        self.slm_fn = None
        self.touched = np.zeros((n, n), dtype=bool)
        self.dint_c = np.zeros(D)
        slm_c = spec.slm_coefficient() if hasattr(spec, "slm_coefficient") else None
        if slm_c is not None:
            if order == 0:
                raise NotImplementedError
            self.slm_fn = make_interp_spline(spec.sampling_times, slm_c, k=order)
            m = np.zeros(n, dtype=bool)
            m[list(spec.slm_targets)] = True
            self.touched = m[:, None] | m[None, :]
        if spec.has_interaction():
            r = spec.eigenbasis.index("u" if spec.interaction_type == "XY" else "r")
            U = spec.pair_matrix()
            nr = (self.digits == r).astype(float)
            for i in range(n):
                for j in range(i + 1, n):
                    if U[i, j] != 0.0:
                        if self.touched[i, j]:
                            self.dint_c += U[i, j] * nr[i] * nr[j]
                        else:
                            self.dint += U[i, j] * nr[i] * nr[j]
        t = spec.sampling_times
        self.fns = []
        for drv in spec.drives:
            to, frm = BASIS_ROLES[drv.basis]
            if order == 0:
                raise NotImplementedError
            cf = make_interp_spline(t, drv.coef.T, k=order)
            df = make_interp_spline(t, drv.det.T, k=order)
            self.fns.append(
                (spec.eigenbasis.index(to), spec.eigenbasis.index(frm), cf, df)
            )

    def apply(self, t_us: float, psi: np.ndarray) -> np.ndarray:
        n, d = self.n, self.d
        psi = np.asarray(psi, dtype=complex)
        cm = float(self.slm_fn(t_us)) if self.slm_fn is not None else 1.0
        out = (self.dint + cm * self.dint_c) * psi
        pt = psi.reshape([d] * n)
        ot = out.reshape([d] * n)
        for to, frm, cf, df in self.fns:
            c = cf(t_us)
            dt = df(t_us)
            for k in range(n):
                s_to = [slice(None)] * n
                s_fr = [slice(None)] * n
                s_to[k] = to
                s_fr[k] = frm
                s_to, s_fr = tuple(s_to), tuple(s_fr)
                # c |to><from| + conj(c) |from><to| - det |from><from|
                ot[s_to] += c[k] * pt[s_fr]
                ot[s_fr] += np.conj(c[k]) * pt[s_to] - dt[k] * pt[s_fr]
        if self.xy is not None:  # U_xy (|u d><d u| + h.c.) on every pair
            U, iu, idn = self.xy
            for i in range(n):
                for j in range(i + 1, n):
                    if U[i, j] == 0.0:
                        continue
                    a = [slice(None)] * n
                    b = [slice(None)] * n
                    a[i], a[j] = iu, idn
                    b[i], b[j] = idn, iu
                    a, b = tuple(a), tuple(b)
                    u = U[i, j] * (cm if self.touched[i, j] else 1.0)
                    ot[a] += u * pt[b]
                    ot[b] += u * pt[a]
        return out

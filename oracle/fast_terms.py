"""ORACLE (test infrastructure, never on the product path).

Index-arithmetic construction of the *same* CSR terms that
``oracle/ref_hamiltonian.py`` builds with Kronecker products (and that the
reference builds with ``qutip.tensor``, hamiltonian.py:145-200), for d = 2 and
a global ground-rydberg drive, so that the CPU baseline can be timed at
N = 20 without minutes of Kronecker assembly.  The operators (and therefore
the per-RHS cost: 5 CSR products) are identical;
``tests/test_oracle_cpu.py`` checks equality on small registers.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import scipy.sparse as sp

from .ref_hamiltonian import OracleHamiltonian


def global_ising_hamiltonian(spec: Any) -> OracleHamiltonian:
    assert spec.dim == 2 and len(spec.drives) == 1 and spec.drives[0].uniform
    assert spec.drives[0].basis == "ground-rydberg" and spec.eigenbasis == ["r", "g"]
    n = spec.n_qudits
    D = 1 << n
    idx = np.arange(D, dtype=np.int64)
    # digit 0 = r, 1 = g; qubit k <-> bit n-1-k
    nr = np.zeros((n, D), dtype=np.int8)
    for k in range(n):
        nr[k] = 1 - ((idx >> (n - 1 - k)) & 1)
    U = spec.pair_matrix()
    diag = np.zeros(D)
    for i in range(n):
        for j in range(i + 1, n):
            if U[i, j] != 0.0:
                diag += 0.5 * U[i, j] * (nr[i] * nr[j])
    terms = []
    if spec.has_interaction():
        terms.append((sp.diags(diag.astype(complex), format="csr"), None))
    drv = spec.drives[0]
    # sum_k |g><r|_k : row has g (bit 1), column has r (bit 0)
    rows, cols = [], []
    for k in range(n):
        p = n - 1 - k
        r = idx[((idx >> p) & 1) == 1]
        rows.append(r)
        cols.append(r ^ (1 << p))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    sig_gr = sp.csr_matrix((np.ones(len(rows), dtype=complex), (rows, cols)), shape=(D, D))
    sig_rr = sp.diags(nr.sum(axis=0).astype(complex), format="csr")
    coeffs = [drv.coef[0], -0.5 * drv.det[0]]
    qobj = list(terms)
    for op, c in zip((sig_gr, sig_rr), coeffs):
        if np.any(c != 0):
            qobj.append((op, c))
    full = list(qobj)
    for a, c in qobj:  # ham + ham.dag(), hamiltonian.py:437
        full.append((sp.csr_matrix(a.conj().T), None if c is None else np.conj(c)))
    return OracleHamiltonian(n, spec.eigenbasis, spec.sampling_times, full, [])

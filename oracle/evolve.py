"""ORACLE (test infrastructure, never on the product path).

CPU restatement of the solver call the reference makes at
``pulser_simulation/simulation.py:729-735``:
``qutip.sesolve / mesolve(H: QobjEvo, psi0, tlist, c_ops, options)``.

QuTiP is third-party and absent here (qutip>=5,<6); its published algorithm
is restated: the Schroedinger / Lindblad right-hand side
``f(t, psi) = -i sum_k c_k(t) (A_k @ psi)`` with CSR ``A_k`` and interpolated
``c_k``, integrated by an ODE solver.  Two integrators are offered:

* ``method="dop853"`` with tight tolerances -- the PARITY oracle;
* ``method="zvode-adams"`` with QuTiP's default options (``method="adams"``,
  ``atol=1e-8``, ``rtol=1e-6``) and pulser's ``max_step`` / ``nsteps``
  (simulation.py:768-780) -- the stand-in for the reference's own CPU run,
  used as ``cpu_baseline`` in bench.py.

PARITY PIN STATUS: see oracle/ref_hamiltonian.py (evolved states: parity
unpinned at 1e-8; pinned to the reference's loose golden vectors only).
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import scipy.sparse as sp
from scipy.integrate import ode, solve_ivp

from .ref_hamiltonian import OracleHamiltonian


class _Rhs:
    """-i H(t) psi as a sum of CSR products (the QuTiP hot loop)."""

    def __init__(self, ham: OracleHamiltonian, order: int) -> None:
        self.ops = [sp.csr_matrix(a) for a, _ in ham.terms]
        self.fns = ham.coefficient_functions(order)
        self.n_calls = 0
        # constant terms merged once (QobjEvo.compress does the same)
        const = [a for a, f in zip(self.ops, self.fns) if f is None]
        self.const = sum(const[1:], const[0]) if const else None
        self.td = [(a, f) for a, f in zip(self.ops, self.fns) if f is not None]
        self.t_lo = ham.sampling_times[0]
        self.t_hi = ham.sampling_times[-1]

    def apply_h(self, t: float, psi: np.ndarray) -> np.ndarray:
        self.n_calls += 1
        t = min(max(t, self.t_lo), self.t_hi)
        out = self.const @ psi if self.const is not None else np.zeros_like(psi)
        for a, f in self.td:
            out = out + complex(f(t)) * (a @ psi)
        return out

    def __call__(self, t: float, psi: np.ndarray) -> np.ndarray:
        return -1j * self.apply_h(t, psi)


def sesolve(
    ham: OracleHamiltonian,
    psi0: np.ndarray,
    eval_times: Sequence[float],
    order: int = 3,
    method: str = "dop853",
    rtol: float = 1e-12,
    atol: float = 1e-14,
    max_step: float | None = 1e-3,
    nsteps: int = 10**9,
    return_stats: bool = False,
) -> Any:
    """States at ``eval_times`` (microseconds); ``normalize_output=False``
    as the reference forces (simulation.py:720-721)."""
    rhs = _Rhs(ham, order)
    eval_times = np.asarray(eval_times, dtype=float)
    psi0 = np.asarray(psi0, dtype=complex).ravel()
    states = [psi0.copy()]
    if method == "dop853":
        # integrate segment-wise so that the step never straddles an
        # evaluation time (dense output would be lower order)
        y = psi0.copy()
        for t0, t1 in zip(eval_times[:-1], eval_times[1:]):
            if t1 > t0:
                sol = solve_ivp(
                    rhs,
                    (t0, t1),
                    y,
                    method="DOP853",
                    rtol=rtol,
                    atol=atol,
                    max_step=max_step if max_step else np.inf,
                )
                assert sol.success, sol.message
                y = sol.y[:, -1]
            states.append(y.copy())
    elif method == "zvode-adams":
        r = ode(rhs)
        r.set_integrator(
            "zvode",
            method="adams",
            atol=atol,
            rtol=rtol,
            max_step=max_step or 0.0,
            nsteps=nsteps,
        )
        r.set_initial_value(psi0, eval_times[0])
        for t1 in eval_times[1:]:
            if t1 > r.t:
                r.integrate(t1)
                assert r.successful()
            states.append(np.array(r.y))
    else:
        raise ValueError(method)
    if return_stats:
        return states, {"rhs_calls": rhs.n_calls}
    return states


def mesolve(
    ham: OracleHamiltonian,
    rho0: np.ndarray,
    eval_times: Sequence[float],
    order: int = 3,
    rtol: float = 1e-11,
    atol: float = 1e-13,
    max_step: float | None = 1e-3,
) -> list[np.ndarray]:
    """Lindblad master equation on the dense density matrix (small N only).

    ``d rho/dt = -i[H, rho] + sum_L (L rho L^+ - 1/2 {L^+ L, rho})`` -- what
    ``qutip.mesolve`` integrates for ``c_ops`` (simulation.py:724-735).
    """
    rhs = _Rhs(ham, order)
    D = ham.dim**ham.n_qudits
    rho0 = np.asarray(rho0, dtype=complex)
    if rho0.ndim == 1 or rho0.shape[-1] == 1:
        v = rho0.ravel()
        rho0 = np.outer(v, v.conj())
    cs = [sp.csr_matrix(c) for c in ham.collapse_ops]
    cdc = sum((c.conj().T @ c for c in cs), sp.csr_matrix((D, D), dtype=complex))

    def f(t: float, y: np.ndarray) -> np.ndarray:
        rho = y.reshape(D, D)
        hr = rhs.apply_h(t, rho)  # H rho (CSR @ dense matrix)
        out = -1j * (hr - hr.conj().T)  # valid for Hermitian rho
        for c in cs:
            out = out + c @ rho @ c.conj().T
        anti = cdc @ rho
        out = out - 0.5 * (anti + anti.conj().T)
        return out.ravel()

    eval_times = np.asarray(eval_times, dtype=float)
    y = rho0.ravel().copy()
    out = [rho0.copy()]
    for t0, t1 in zip(eval_times[:-1], eval_times[1:]):
        if t1 > t0:
            sol = solve_ivp(
                f, (t0, t1), y, method="DOP853", rtol=rtol, atol=atol,
                max_step=max_step if max_step else np.inf,
            )
            assert sol.success, sol.message
            y = sol.y[:, -1]
        out.append(y.reshape(D, D).copy())
    return out


def all_ground_state(ham_or_spec: Any) -> np.ndarray:
    """simulation.py:498-505: every qudit in ``g`` (``u`` for XY)."""
    eig = list(ham_or_spec.eigenbasis)
    d = len(eig)
    n = ham_or_spec.n_qudits
    g = eig.index("u") if "u" in eig and "g" not in eig else eig.index("g")
    idx = 0
    for _ in range(n):
        idx = idx * d + g
    psi = np.zeros(d**n, dtype=complex)
    psi[idx] = 1.0
    return psi

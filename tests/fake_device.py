"""TEST-ONLY stand-ins for the device plans, backed by the CPU oracle, so that
the host logic of ``B200Emulator`` (evaluation times, trajectories, sampling,
result wrapping) can be exercised where there is no GPU.  Never imported by the
product."""
from __future__ import annotations

import numpy as np

from oracle import evolve
from oracle.ref_hamiltonian import OracleHamiltonian


class FakeDevicePlan:
    calls = 0

    def __init__(self, specs, interp_order=3, device=0):
        from pulser_b200.spec import HamiltonianSpec

        self.specs = [specs] if isinstance(specs, HamiltonianSpec) else list(specs)
        self.hams = [OracleHamiltonian.from_spec(s) for s in self.specs]
        self.order = interp_order
        # the coefficients are smooth between two samples: one sampling interval bounds the integrator step (the
        # oracle's default, 1 ns, is QuTiP's max_step and makes coarse-sampled tests 10-100x slower for nothing)
        self.max_step = float(np.min(np.diff(self.specs[0].sampling_times)))
        self.states = None
        FakeDevicePlan.calls += 1

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def set_state(self, psi):
        psi = np.asarray(psi, dtype=complex).reshape(-1)
        self.states = [psi.copy() for _ in self.hams]

    def propagate(self, t0, t1, **opts):
        self.states = [
            evolve.sesolve(h, s, [t0, t1], order=self.order, rtol=1e-10, atol=1e-12, max_step=self.max_step)[-1]
            for h, s in zip(self.hams, self.states)
        ]
        return {"n_steps": 1, "n_applies": 1, "n_launches": 0, "max_rho": 0.0}

    def get_state(self):
        return np.stack(self.states)

    def sample(self, n_samples, one_state, traj=0):
        from pulser_b200.results import B200Result, StateVector

        spec = self.specs[traj]
        meas = {"r": "ground-rydberg", "h": "digital", "d": "XY"}[one_state]
        matching = meas in spec.basis_name
        res = B200Result(tuple(range(spec.n_qudits)), meas, StateVector(self.states[traj]), matching)
        return res.get_samples(n_samples)

    # --- device reductions used by DeviceStateView (numpy stand-ins) ---
    @property
    def n(self):
        return self.specs[0].n_qudits

    def norm2(self):
        return np.array([np.vdot(s, s).real for s in self.states])

    def _digits(self):
        spec = self.specs[0]
        idx = np.arange(spec.hilbert_dim)
        return [(idx // spec.dim ** (spec.n_qudits - 1 - k)) % spec.dim for k in range(spec.n_qudits)]

    def occupation(self, digit, traj0=0, count=None):
        return np.stack([np.diagonal(c) for c in self.correlation(digit, traj0, count)])

    def correlation(self, digit, traj0=0, count=None):
        count = len(self.states) - traj0 if count is None else count
        dg = self._digits()
        out = np.zeros((count, self.n, self.n))
        for c in range(count):
            p = np.abs(self.states[traj0 + c]) ** 2
            for i in range(self.n):
                for j in range(self.n):
                    out[c, i, j] = p[(dg[i] == digit) & (dg[j] == digit)].sum()
        return out

    def energy(self, t_us):
        hs = [h.matrix_at(t_us, self.order) @ s for h, s in zip(self.hams, self.states)]
        return (np.array([np.vdot(s, w).real for s, w in zip(self.states, hs)]),
                np.array([np.vdot(w, w).real for w in hs]))

    def overlap(self, phi, traj0=0, count=None):
        count = len(self.states) - traj0 if count is None else count
        return np.array([np.vdot(phi, self.states[traj0 + c]) for c in range(count)])

    def apply_h(self, t_us, vec, traj=0):
        return self.hams[traj].matrix_at(t_us, self.order) @ np.asarray(vec, dtype=complex)

    @property
    def n_traj(self):
        return len(self.hams)

    def copy_state_from(self, other, src_traj=0, dst_traj=0):
        if self.states is None:
            self.states = [None] * len(self.hams)
        self.states[dst_traj] = np.array(other.states[src_traj], dtype=complex)

    def set_collapse(self, ops, seed=0):
        raise NotImplementedError("the oracle-backed fake has no Monte-Carlo wave-function path")


class FakeLindbladPlan(FakeDevicePlan):
    def set_state(self, psi):
        psi = np.asarray(psi, dtype=complex).reshape(-1)
        self.states = [np.outer(psi, psi.conj()) for _ in self.hams]

    def propagate(self, t0, t1, **opts):
        self.states = [
            evolve.mesolve(h, s, [t0, t1], order=self.order, rtol=1e-9, atol=1e-11, max_step=self.max_step)[-1]
            for h, s in zip(self.hams, self.states)
        ]
        return {"n_steps": 1, "n_applies": 1, "n_launches": 0, "max_rho": 0.0}

    def get_rho(self):
        return np.stack(self.states)

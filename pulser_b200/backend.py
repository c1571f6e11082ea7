"""``B200Backend``: the ``pulser.backend.EmulatorBackend`` plugin (seam S2).

Mirrors ``pulser_simulation.qutip_backend.QutipBackendV2``
(``pulser-simulation/pulser_simulation/qutip_backend.py:121-325``) with
``B200Config`` / ``B200State`` / ``B200Operator`` in place of
``QutipConfig`` / ``QutipState`` / ``QutipOperator``
(``qutip_config.py:28-192``, ``qutip_state.py:35-281``, ``qutip_op.py:30-259``).
States come from the CUDA path (``B200Emulator``); the Hamiltonian handed to the
observables applies ``H(t)`` on the device (``pb200_apply_h``), so ``Energy`` &
co. never build a matrix.
"""
from __future__ import annotations

import math
import warnings
from collections import Counter, defaultdict
from typing import Any, Literal, Mapping, Sequence, Type

import numpy as np
import scipy.sparse as sp

from ._compat import ensure_pulser

if not ensure_pulser():  # pragma: no cover
    raise ImportError("pulser_b200.backend needs pulser-core")

import pulser  # noqa: E402
from pulser.backend.abc import EmulatorBackend  # noqa: E402
from pulser.backend.config import EmulationConfig  # noqa: E402
from pulser.backend.default_observables import BitStrings, StateResult  # noqa: E402
from pulser.backend.operator import Operator  # noqa: E402
from pulser.backend.results import Results  # noqa: E402
from pulser.backend.state import State  # noqa: E402

from .emulator import B200Emulator, Solver, _has_stochastic_noise  # noqa: E402
from .results import multinomial  # noqa: E402


class B200State(State[complex, float]):
    """A state vector or density matrix as a numpy array (``QutipState`` mirror)."""

    def __init__(self, state: np.ndarray, *, eigenstates: Sequence[str]):
        super().__init__(eigenstates=eigenstates)
        arr = np.asarray(state.full() if hasattr(state, "full") else state, dtype=np.complex128)
        if arr.ndim == 2 and arr.shape[1] == 1:
            arr = arr.reshape(-1)
        if arr.ndim == 2 and arr.shape[0] == 1:
            arr = arr.reshape(-1).conj()  # a bra
        if arr.ndim not in (1, 2) or (arr.ndim == 2 and arr.shape[0] != arr.shape[1]):
            raise TypeError("'state' must be a ket (1-D) or a density matrix (square 2-D).")
        self._state = arr
        n = math.log(arr.shape[0], self.qudit_dim)
        if not np.isclose(n, round(n)):
            raise ValueError(
                f"A state with shape {arr.shape} is incompatible with "
                f"a system of {self.qudit_dim}-level qudits."
            )

    @property
    def is_ket(self) -> bool:
        return self._state.ndim == 1

    @property
    def n_qudits(self) -> int:
        return round(math.log(self._state.shape[0], self.qudit_dim))

    def to_array(self) -> np.ndarray:
        return self._state.copy()

    def overlap(self, other: "B200State") -> float:
        if not isinstance(other, B200State):
            raise TypeError(
                "'B200State.overlap()' expects another 'B200State', not " f"{type(other)}."
            )
        if self.n_qudits != other.n_qudits or self.qudit_dim != other.qudit_dim:
            raise ValueError(
                "Can't calculate the overlap between a state with "
                f"{self.n_qudits} {self.qudit_dim}-dimensional qudits and "
                f"another with {other.n_qudits} {other.qudit_dim}-dimensional "
                "qudits."
            )
        if self.eigenstates != other.eigenstates:
            msg = (
                "Can't calculate the overlap between states with eigenstates "
                f"{self.eigenstates} and {other.eigenstates}."
            )
            if set(self.eigenstates) != set(other.eigenstates):
                raise ValueError(msg)
            raise NotImplementedError(msg)
        a, b = self._state, other._state
        if a.ndim == 1 and b.ndim == 1:
            return float(np.abs(np.vdot(a, b)) ** 2)
        if a.ndim == 1:
            return float(np.vdot(a, b @ a).real)
        if b.ndim == 1:
            return float(np.vdot(b, a @ b).real)
        return float(np.trace(a @ b).real)

    def probabilities(self, *, cutoff: float = 1e-12) -> dict[str, float]:
        probs = (np.abs(self._state) ** 2) if self.is_ket else np.abs(np.diagonal(self._state))
        non_zero = np.argwhere(probs > cutoff).flatten()
        probs = probs[non_zero]
        probs = probs / np.sum(probs)
        return dict(zip(map(self.get_basis_state_from_index, non_zero), probs))

    def bitstring_probabilities(self, *, one_state: str | None = None, cutoff: float = 1e-12) -> Mapping[str, float]:
        one_state = one_state or self.infer_one_state()
        zero_states = set(self.eigenstates) - {one_state}
        probs = self.probabilities(cutoff=cutoff)
        out: dict[str, float] = defaultdict(float)
        for state_str, p in probs.items():
            bitstring = state_str.replace(one_state, "1")
            for s_ in zero_states:
                bitstring = bitstring.replace(s_, "0")
            out[bitstring] += p
        return dict(out)

    def sample(self, *, num_shots: int, one_state: str | None = None, p_false_pos: float = 0.0,
               p_false_neg: float = 0.0) -> Counter:
        """qutip_state.py:169-218 (same RNG consumption)."""
        bitstring_probs = self.bitstring_probabilities(one_state=one_state, cutoff=1 / (1000 * num_shots))
        bitstrings = np.array(list(bitstring_probs))
        probs = np.array(list(map(float, bitstring_probs.values())))
        indices = multinomial(num_shots, probs)
        if p_false_pos == 0.0 and p_false_neg == 0.0:
            return Counter(bitstrings[indices].tolist())
        bitstr_arr = np.array([list(bs) for bs in bitstrings[indices]], dtype=int)
        flip_probs = np.where(bitstr_arr == 1, p_false_neg, p_false_pos)
        flips = np.random.uniform(size=flip_probs.shape) < flip_probs
        new_counts: Counter = Counter(map(tuple, bitstr_arr ^ flips))
        return Counter({"".join(map(str, k)): v for k, v in new_counts.items()})

    @classmethod
    def _from_state_amplitudes(cls, *, eigenstates: Sequence[str], n_qudits: int,
                               amplitudes: Mapping[str, complex]):
        d = len(eigenstates)
        vec = np.zeros(d**n_qudits, dtype=np.complex128)
        amps = {k: complex(v) for k, v in amplitudes.items()}
        for basis_state, amp in amps.items():
            idx = 0
            for s in basis_state:
                idx = idx * d + eigenstates.index(s)
            vec[idx] += amp
        return cls(vec, eigenstates=eigenstates), amps

    def __repr__(self) -> str:
        return f"B200State(eigenstates={self.eigenstates}, shape={self._state.shape})"

    def __eq__(self, other: Any) -> bool:
        return (
            isinstance(other, B200State)
            and self.eigenstates == other.eigenstates
            and self._state.shape == other._state.shape
            and bool(np.allclose(self._state, other._state, atol=1e-12, rtol=0))
        )


class B200Operator(Operator[complex, complex, B200State]):
    """An operator as a scipy sparse matrix (``QutipOperator`` mirror)."""

    def __init__(self, operator: Any, eigenstates: Sequence[str]):
        super().__init__()
        B200State._validate_eigenstates(eigenstates)
        self._eigenstates = eigenstates
        mat = sp.csr_matrix(operator.full() if hasattr(operator, "full") else operator, dtype=np.complex128)
        if mat.shape[0] != mat.shape[1]:
            raise TypeError("'operator' must be a square matrix.")
        self._operator = mat

    @property
    def eigenstates(self) -> tuple[str, ...]:
        return tuple(self._eigenstates)

    def to_array(self) -> np.ndarray:
        return self._operator.toarray()

    def _validate_other(self, other: Any, expected_type: Type, op_name: str) -> None:
        if not isinstance(other, expected_type):
            raise TypeError(f"'{op_name}' expects a '{expected_type.__name__}' instance, not {type(other)}.")
        if self.eigenstates != other.eigenstates:
            msg = (
                f"Can't apply {op_name} between a {self.__class__.__name__} "
                f"with eigenstates {self.eigenstates} and a "
                f"{other.__class__.__name__} with {other.eigenstates}."
            )
            if set(self.eigenstates) != set(other.eigenstates):
                raise ValueError(msg)
            raise NotImplementedError(msg)

    def _matvec(self, arr: np.ndarray) -> np.ndarray:
        return self._operator @ arr

    def apply_to(self, state: B200State, /) -> B200State:
        self._validate_other(state, B200State, "B200Operator.apply_to()")
        out = self._matvec(state._state)
        if not state.is_ket:
            out = self._matvec(out.conj().T).conj().T  # O rho O^+
        return type(state)(out, eigenstates=state.eigenstates)

    @property
    def _isherm(self) -> bool:
        if not hasattr(self, "_herm_cache"):
            m = self._operator
            self._herm_cache = bool(abs(m - m.getH()).max() < 1e-12) if m.nnz else True
        return self._herm_cache

    def expect(self, state: B200State, /) -> complex:
        """``qutip.expect`` semantics: a real number for a Hermitian operator."""
        self._validate_other(state, B200State, "B200Operator.expect()")
        if state.is_ket:
            val = complex(np.vdot(state._state, self._matvec(state._state)))
        else:
            val = complex(np.trace(self._matvec(state._state)))
        return val.real if self._isherm else val

    def __add__(self, other: "B200Operator", /) -> "B200Operator":
        self._validate_other(other, B200Operator, "__add__")
        return B200Operator(self._operator + other._operator, eigenstates=self.eigenstates)

    def __rmul__(self, scalar: complex) -> "B200Operator":
        return B200Operator(complex(scalar) * self._operator, eigenstates=self.eigenstates)

    def __matmul__(self, other: "B200Operator") -> "B200Operator":
        self._validate_other(other, B200Operator, "__matmul__")
        return B200Operator(self._operator @ other._operator, eigenstates=self.eigenstates)

    @classmethod
    def _from_operator_repr(cls, *, eigenstates: Sequence[str], n_qudits: int, operations: Any):
        """qutip_op.py:150-220 with scipy.sparse Kronecker products."""
        d = len(eigenstates)

        def qudit_op(qop: Mapping[str, complex]) -> sp.csr_matrix:
            m = sp.lil_matrix((d, d), dtype=np.complex128)
            for proj_str, coeff in qop.items():
                m[eigenstates.index(proj_str[0]), eigenstates.index(proj_str[1])] += complex(coeff)
            return m.tocsr()

        full = sp.csr_matrix((d**n_qudits, d**n_qudits), dtype=np.complex128)
        reconstructed = []
        for coeff, tensor_op in operations:
            factors = [sp.identity(d, format="csr", dtype=np.complex128) for _ in range(n_qudits)]
            re_tensor = []
            for qop, inds in tensor_op:
                for ind in inds:
                    factors[ind] = qudit_op(qop)
                re_tensor.append(({k: complex(v) for k, v in qop.items()}, set(inds)))
            term = factors[0]
            for f in factors[1:]:
                term = sp.kron(term, f, format="csr")
            full = full + complex(coeff) * term
            reconstructed.append((complex(coeff), re_tensor))
        return B200Operator(full, eigenstates=eigenstates), reconstructed

    def __repr__(self) -> str:
        return f"B200Operator(eigenstates={self.eigenstates}, shape={self._operator.shape})"

    def __eq__(self, other: Any) -> bool:
        return (
            isinstance(other, B200Operator)
            and self.eigenstates == other.eigenstates
            and abs(self._operator - other._operator).max() < 1e-12
        )


class DeviceHamiltonian(B200Operator):
    """``H(t)`` of the noiseless sequence, applied matrix-free on the GPU.

    Stands for the ``QutipOperator(noiseless H(t))`` the reference hands to every
    observable (``qutip_backend.py:258-264``)."""

    def __init__(self, plan: Any, t_us: float, eigenstates: Sequence[str]):
        Operator.__init__(self)
        self._eigenstates = eigenstates
        self._plan = plan
        self._t = t_us
        self._operator = None  # never materialised
        self._herm_cache = True

    def _matvec(self, arr: np.ndarray) -> np.ndarray:
        if arr.ndim == 1:
            return self._plan.apply_h(self._t, arr)
        return np.stack([self._plan.apply_h(self._t, np.ascontiguousarray(arr[:, j])) for j in range(arr.shape[1])], axis=1)

    def __add__(self, other, /):  # pragma: no cover - not needed by the default observables
        raise NotImplementedError("DeviceHamiltonian is matrix-free")

    __rmul__ = __matmul__ = __add__


class B200Config(EmulationConfig[B200State]):
    """``QutipConfig`` mirror (``qutip_config.py:28-192``): same options."""

    _enforce_expected_kwargs = True
    sampling_rate: float
    _state_type = B200State
    _operator_type = B200Operator
    solver: Solver

    def __init__(self, *, sampling_rate: float = 1.0,
                 solver: Solver | Literal["default", "MasterEquation", "MonteCarlo"] = Solver.DEFAULT,
                 print_progress: bool = False, progress_bar: bool = False, **backend_options: Any):
        if backend_options.setdefault("interaction_matrix") is not None:
            raise NotImplementedError("'B200Backend' does not handle custom interaction matrices.")
        if not (0 < sampling_rate <= 1.0):
            raise ValueError(
                f"The sampling rate (`sampling_rate` = {sampling_rate}) must"
                " be greater than 0 and less than or equal to 1."
            )
        initial_state = backend_options.setdefault("initial_state")
        if initial_state is not None and not isinstance(initial_state, B200State):
            raise TypeError(
                "If provided, `initial_state` must be an instance of "
                f"`B200State`, not {type(initial_state)}."
            )
        try:
            solver = Solver(solver)
        except ValueError:
            allowed_str = ", ".join(s.value for s in Solver)
            raise ValueError(f"Invalid solver '{solver}'. Allowed solvers are: {allowed_str}.")
        super().__init__(sampling_rate=sampling_rate, solver=Solver(solver), print_progress=print_progress,
                         progress_bar=progress_bar, **backend_options)

    def _expected_kwargs(self) -> set[str]:
        return super()._expected_kwargs() | {"sampling_rate", "solver", "print_progress", "progress_bar"}

    def _get_legacy_evaluation_times(self, total_duration_ns: int):
        """qutip_config.py:169-192: relative observable times -> microseconds."""
        extra: set[float] = set()
        if self.callbacks:
            return "Full"
        for obs in self.observables:
            if obs.evaluation_times is not None:
                extra.update(obs.evaluation_times)
        rel = self.default_evaluation_times
        if isinstance(rel, str) and rel == "Full":
            if not extra:
                return "Full"
            idx = np.linspace(0, total_duration_ns - 1, int(self.sampling_rate * total_duration_ns), dtype=int)
            rel = idx / total_duration_ns
        rel = np.union1d(np.asarray(rel, dtype=float), np.array(sorted(extra), dtype=float))
        return np.asarray(rel) * total_duration_ns * 1e-3


class B200Backend(EmulatorBackend):
    """Emulate a sequence on a B200 through the generic ``pulser.backend`` API."""

    default_config = B200Config(observables=[BitStrings(evaluation_times=[1.0]), StateResult()])
    _config: B200Config

    def __init__(self, sequence: pulser.Sequence, *, config: EmulationConfig | None = None,
                 mimic_qpu: bool = False) -> None:
        super().__init__(sequence, config=config, mimic_qpu=mimic_qpu)
        noise_model = None
        if self._config.prefer_device_noise_model:
            noise_model = sequence.device.noise_model
        noise_model = noise_model or self._config.noise_model
        self._sim_obj = B200Emulator.from_sequence(
            sequence,
            sampling_rate=self._config.sampling_rate,
            noise_model=noise_model,
            with_modulation=self._config.with_modulation,
            solver=self._config.solver,
            n_trajectories=self._config.n_trajectories,
        )
        self._sim_obj.set_evaluation_times(
            self._config._get_legacy_evaluation_times(self._sim_obj.total_duration_ns)
        )
        if self._config.initial_state is not None:
            self._sim_obj.set_initial_state(self._config.initial_state.to_array())

    def _replay(self, plan: Any, coherent: Any, res: Results) -> None:
        """Feed the stored states to callbacks / observables (qutip_backend.py:254-280)."""
        sim, config = self._sim_obj, self._config
        eig = sim._hamiltonian_data.basis_data.eigenbasis
        for r in coherent:
            t = float(r.evaluation_time)
            raw = r.state.full()
            raw = raw.reshape(-1) if r.state.isket else raw
            nrm = np.linalg.norm(raw) if r.state.isket else np.trace(raw).real
            state = B200State(raw / nrm, eigenstates=eig)
            ham = DeviceHamiltonian(plan, t * res.total_duration / 1000, eig)
            for callback in config.callbacks:
                callback(config=config, t=t, state=state, hamiltonian=ham, result=res)
            for obs in config.observables:
                obs(config=config, t=t, state=state, hamiltonian=ham, result=res)

    def run(self) -> Results:
        from . import engine

        sim = self._sim_obj
        opts = {"print_progress": self._config.print_progress, "progress_bar": self._config.progress_bar}
        atom_order = tuple(sim._register.qubit_ids)
        with engine.DevicePlan(sim._noiseless_spec(), sim._interp_order, sim._gpu) as hplan:
            if not _has_stochastic_noise(sim.noise_model):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", DeprecationWarning)
                    single = sim.run(**opts)
                res = Results(atom_order=atom_order, total_duration=sim.total_duration_ns)
                self._replay(hplan, single, res)
                return res
            results: list[Results] = []
            sim._validate_options({})
            sim._check_supported()
            for cleanres, reps in sim._noisy_runs(print_progress=self._config.print_progress, batch=0,
                                                  opts={"max_step": 0, "cheb_tol": 0.0, "refine_window": -1, "tol": 0.0}):
                for _ in range(reps):
                    res = Results(atom_order=atom_order, total_duration=sim.total_duration_ns)
                    self._replay(hplan, cleanres, res)
                    results.append(res)
            return Results.aggregate(results)

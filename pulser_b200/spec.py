"""Plain-array description of one Hamiltonian: the input of the C-ABI.

``HamiltonianSpec`` is the qutip-free image of the arguments of
``pulser_simulation.hamiltonian.Hamiltonian.__init__``
(reference ``pulser-simulation/pulser_simulation/hamiltonian.py:45-81``):
sampled amp/det/phase per addressing/basis/qubit
(``SequenceSamples.to_nested_dict()``, ``pulser-core/pulser/sampler/samples.py:524-621``),
the trajectory's interaction matrix and bad atoms
(``pulser-core/pulser/_hamiltonian_data/noise_trajectory.py``), the basis
(``basis_data.py``) and the per-qubit collapse operators
(``lindblad_data.py`` + ``hamiltonian.py:97-124``).

Folding rule used here (exact, because QobjEvo coefficient interpolation is
linear in the sample arrays and every reference term is a single-qubit
operator): for each addressed basis the Global coefficient array and the
Local arrays of qubit *k* are summed into one per-qubit table
``coef[basis][k, :] = 0.5 * amp * exp(-1j * phase)`` and
``det[basis][k, :]`` (``hamiltonian.py:349-352, 370-375``).  ``uniform`` marks
tables whose rows are all identical (pure Global drive) so that the device can
use the cheaper global-drive kernel.
"""
from __future__ import annotations

import dataclasses
import io
from typing import Any, Mapping, Sequence

import numpy as np

# Order in which pulser ranks eigenstates
# (reference pulser-core/pulser/channels/base_channel.py:49-57).
STATES_RANK = ("u", "d", "r", "g", "h", "x")

# (target, source) eigenstates of the drive operator sigma_{target source} and
# the detuning projector sigma_{source source} per addressed basis
# (reference hamiltonian.py:340-345).
BASIS_ROLES = {
    "ground-rydberg": ("g", "r"),
    "digital": ("h", "g"),
    "XY": ("u", "d"),
}


@dataclasses.dataclass
class DriveTable:
    """Per-qubit drive/detuning sample tables of one addressed basis."""

    basis: str  # "ground-rydberg" | "digital" | "XY"
    coef: np.ndarray  # complex128 [N, nt]: 0.5*amp*exp(-i phase)
    det: np.ndarray  # float64 [N, nt]: detuning (rad/us), enters H as -det
    uniform: bool  # all rows identical (pure Global addressing)


@dataclasses.dataclass
class HamiltonianSpec:
    """Everything the hot path needs, as plain arrays."""

    n_qudits: int
    dim: int
    eigenbasis: list[str]
    basis_name: str
    interaction_type: str  # "ising" | "XY"
    sampling_times: np.ndarray  # float64 [nt], microseconds
    total_duration_ns: int  # T (before the +1 extension)
    interaction_matrix: np.ndarray  # float64 [K, N, N]
    bad_atoms: np.ndarray  # bool [N]
    drives: list[DriveTable]
    # per-qubit collapse operators: complex128 [n_ops, dim, dim] (coefficient
    # already multiplied in), each applied to every qubit.
    collapse_ops: np.ndarray
    qubit_ids: list[str] = dataclasses.field(default_factory=list)
    # XY + SLM mask: interaction switched off until ``slm_end`` for masked
    # qubits (reference hamiltonian.py:399-424).
    slm_end: int = 0
    slm_targets: list[int] = dataclasses.field(default_factory=list)

    def slm_coefficient(self) -> np.ndarray | None:
        """Samples of the coefficient of the *unmasked* interaction term, or None.

        Restates ``hamiltonian.py:399-411``: ``coeff = ones(duration - 1);
        coeff[:slm_end] = 0`` pushed through ``_adapt_to_sampling_rate``
        (``:87-95``; duration = T + 1 extended samples, so the index grid has
        one more entry than the array it indexes -- kept as is).  Only XY mode
        has a time-dependent interaction; elsewhere the mask acts through the
        drive samples alone.
        """
        if self.interaction_type != "XY" or self.slm_end <= 0 or not len(self.slm_targets):
            return None
        duration = int(self.total_duration_ns) + 1
        coeff = np.ones(duration - 1)
        coeff[0 : self.slm_end] = 0
        nt = len(self.sampling_times)
        idx = np.linspace(0, len(coeff) - 1, nt, dtype=int)
        return coeff[idx]

    # ------------------------------------------------------------------
    @property
    def hilbert_dim(self) -> int:
        return self.dim**self.n_qudits

    @property
    def n_times(self) -> int:
        return len(self.sampling_times)

    def has_interaction(self) -> bool:
        """Whether the static interaction term exists.

        Reference ``hamiltonian.py:393-396``.
        """
        effective_size = self.n_qudits - int(np.sum(self.bad_atoms))
        return "digital" not in self.basis_name and effective_size > 1

    def pair_matrix(self) -> np.ndarray:
        """U_ij with bad atoms removed; [N, N] symmetric, zero diagonal.

        Reference ``hamiltonian.py:260-274, 310-331`` (ising only).
        """
        n = self.n_qudits
        u = np.array(self.interaction_matrix[-1], dtype=np.float64)
        good = ~np.asarray(self.bad_atoms, dtype=bool)
        u = u * good[:, None] * good[None, :]
        u[np.arange(n), np.arange(n)] = 0.0
        if not self.has_interaction():
            u[:] = 0.0
        return u

    def xy_matrix(self) -> np.ndarray:
        """XY exchange couplings U^xy_ij (C3 (1 - 3 cos^2) / r^3) with bad atoms removed.

        Reference ``hamiltonian.py:276-294`` (``interaction_matrix[0]`` in XY mode).
        """
        n = self.n_qudits
        u = np.array(self.interaction_matrix[0], dtype=np.float64)
        good = ~np.asarray(self.bad_atoms, dtype=bool)
        u = u * good[:, None] * good[None, :]
        u[np.arange(n), np.arange(n)] = 0.0
        if not self.has_interaction() or self.interaction_type != "XY":
            u[:] = 0.0
        return u

    # ------------------------------------------------------------------
    def to_npz_dict(self) -> dict[str, Any]:
        out: dict[str, Any] = dict(
            n_qudits=self.n_qudits,
            dim=self.dim,
            eigenbasis=np.array(self.eigenbasis),
            basis_name=self.basis_name,
            interaction_type=self.interaction_type,
            sampling_times=self.sampling_times,
            total_duration_ns=self.total_duration_ns,
            interaction_matrix=self.interaction_matrix,
            bad_atoms=self.bad_atoms,
            collapse_ops=self.collapse_ops,
            qubit_ids=np.array([str(q) for q in self.qubit_ids]),
            slm_end=self.slm_end,
            slm_targets=np.array(self.slm_targets, dtype=np.int64),
            n_drives=len(self.drives),
        )
        for i, d in enumerate(self.drives):
            out[f"drive{i}_basis"] = d.basis
            out[f"drive{i}_uniform"] = d.uniform
            if d.uniform:  # store one row only
                out[f"drive{i}_coef"] = d.coef[:1]
                out[f"drive{i}_det"] = d.det[:1]
            else:
                out[f"drive{i}_coef"] = d.coef
                out[f"drive{i}_det"] = d.det
        return out

    def save(self, path: str, **extra: Any) -> None:
        np.savez_compressed(path, **self.to_npz_dict(), **extra)

    @classmethod
    def from_npz(cls, data: Mapping[str, Any]) -> "HamiltonianSpec":
        n = int(data["n_qudits"])
        drives = []
        for i in range(int(data["n_drives"])):
            uniform = bool(data[f"drive{i}_uniform"])
            coef = np.array(data[f"drive{i}_coef"], dtype=np.complex128)
            det = np.array(data[f"drive{i}_det"], dtype=np.float64)
            if uniform:
                coef = np.repeat(coef, n, axis=0)
                det = np.repeat(det, n, axis=0)
            drives.append(
                DriveTable(str(data[f"drive{i}_basis"]), coef, det, uniform)
            )
        return cls(
            n_qudits=n,
            dim=int(data["dim"]),
            eigenbasis=[str(s) for s in data["eigenbasis"]],
            basis_name=str(data["basis_name"]),
            interaction_type=str(data["interaction_type"]),
            sampling_times=np.array(data["sampling_times"], dtype=np.float64),
            total_duration_ns=int(data["total_duration_ns"]),
            interaction_matrix=np.array(
                data["interaction_matrix"], dtype=np.float64
            ),
            bad_atoms=np.array(data["bad_atoms"], dtype=bool),
            drives=drives,
            collapse_ops=np.array(data["collapse_ops"], dtype=np.complex128),
            qubit_ids=[str(q) for q in data["qubit_ids"]],
            slm_end=int(data["slm_end"]),
            slm_targets=[int(t) for t in data["slm_targets"]],
        )

    @classmethod
    def load(cls, path: str | io.BytesIO) -> "HamiltonianSpec":
        with np.load(path, allow_pickle=False) as data:
            return cls.from_npz(data)


# ----------------------------------------------------------------------
# Host mirror of Hamiltonian.__init__ input handling
# ----------------------------------------------------------------------
def adapt_to_sampling_rate(
    full_array: np.ndarray, sampling_rate: float, duration: int
) -> np.ndarray:
    """Subsample like ``Hamiltonian._adapt_to_sampling_rate``.

    Reference ``hamiltonian.py:87-95``: ``linspace(0, len-1,
    int(rate*duration), dtype=int)``; note the integer truncation makes the
    grid non-uniform for ``rate < 1``.
    """
    indices = np.linspace(
        0, len(full_array) - 1, int(sampling_rate * duration), dtype=int
    )
    return np.asarray(full_array)[..., indices]


def collapse_matrices(
    eigenbasis: Sequence[str],
    local_collapse_ops: Sequence[tuple[Any, Any]],
    depolarizing_pauli_2ds: Mapping[str, Sequence[tuple[Any, str]]],
) -> np.ndarray:
    """Single-qudit collapse matrices (coefficient included).

    Reference ``hamiltonian.py:97-124``: a named projector ``sigma_ab`` is
    ``|a><b|``; a Pauli label is the listed combination of projectors; an
    ndarray is used as is.
    """
    d = len(eigenbasis)
    idx = {s: i for i, s in enumerate(eigenbasis)}

    def proj(name: str) -> np.ndarray:
        assert name.startswith("sigma_") and len(name) == 8, name
        m = np.zeros((d, d), dtype=np.complex128)
        m[idx[name[6]], idx[name[7]]] = 1.0
        return m

    mats = []
    for coeff, op in local_collapse_ops:
        if isinstance(op, str):
            if op in depolarizing_pauli_2ds:
                m = sum(
                    coeff * pc * proj(pn)
                    for pc, pn in depolarizing_pauli_2ds[op]
                )
            else:
                m = coeff * proj(op)
        else:
            m = coeff * np.asarray(op, dtype=np.complex128)
        mats.append(np.asarray(m, dtype=np.complex128))
    if not mats:
        return np.zeros((0, d, d), dtype=np.complex128)
    return np.stack(mats)


def spec_from_pulser(
    samples: Any,
    noise_trajectory: Any,
    basis_data: Any,
    lindblad_data: Any,
    sampling_rate: float = 1.0,
    total_duration_ns: int | None = None,
) -> HamiltonianSpec:
    """Build the spec from the exact arguments of the reference ``Hamiltonian``.

    ``samples`` is the (possibly noisy) ``SequenceSamples`` already extended
    to T+1 (``simulation.py:172-173``); ``noise_trajectory`` a
    ``NoiseTrajectory``.  Mirrors ``hamiltonian.py:45-81, 333-439``.
    """
    register = noise_trajectory.register
    qids = list(register.qubits)
    qindex = {q: i for i, q in enumerate(qids)}
    n = len(qids)
    duration = samples.max_duration
    times = adapt_to_sampling_rate(
        np.arange(duration, dtype=np.double) / 1000, sampling_rate, duration
    )
    nt = len(times)
    nested = samples.to_nested_dict()
    drives: dict[str, DriveTable] = {}
    touched_local: dict[str, bool] = {}

    def table(basis: str) -> DriveTable:
        if basis not in drives:
            drives[basis] = DriveTable(
                basis,
                np.zeros((n, nt), dtype=np.complex128),
                np.zeros((n, nt), dtype=np.float64),
                True,
            )
            touched_local[basis] = False
        return drives[basis]

    for addr in nested:
        for basis in nested[addr]:
            entry = nested[addr][basis]
            if not entry:
                continue
            if addr == "Global":
                c = 0.5 * np.asarray(entry["amp"]) * np.exp(
                    -1j * np.asarray(entry["phase"])
                )
                c = adapt_to_sampling_rate(c, sampling_rate, duration)
                dt = adapt_to_sampling_rate(
                    np.asarray(entry["det"], dtype=np.float64),
                    sampling_rate,
                    duration,
                )
                if np.any(c != 0) or np.any(dt != 0):
                    t = table(basis)
                    t.coef += c[None, :]
                    t.det += dt[None, :]
            else:
                for qid, sq in entry.items():
                    c = 0.5 * np.asarray(sq["amp"]) * np.exp(
                        -1j * np.asarray(sq["phase"])
                    )
                    c = adapt_to_sampling_rate(c, sampling_rate, duration)
                    dt = adapt_to_sampling_rate(
                        np.asarray(sq["det"], dtype=np.float64),
                        sampling_rate,
                        duration,
                    )
                    if np.any(c != 0) or np.any(dt != 0):
                        t = table(basis)
                        k = qindex[qid]
                        t.coef[k] += c
                        t.det[k] += dt
                        touched_local[basis] = True
    for basis, t in drives.items():
        t.uniform = (not touched_local[basis]) or bool(
            np.all(t.coef == t.coef[:1]) and np.all(t.det == t.det[:1])
        )

    bad = np.array(
        [bool(noise_trajectory.bad_atoms[q]) for q in qids], dtype=bool
    )
    imat = np.array(
        noise_trajectory.interaction_matrix.as_array(detach=True),
        dtype=np.float64,
    )
    slm_targets = [qindex[q] for q in samples._slm_mask.targets]
    return HamiltonianSpec(
        n_qudits=n,
        dim=basis_data.dim,
        eigenbasis=list(basis_data.eigenbasis),
        basis_name=basis_data.basis_name,
        interaction_type=basis_data.interaction_type,
        sampling_times=np.asarray(times, dtype=np.float64),
        total_duration_ns=(
            int(total_duration_ns)
            if total_duration_ns is not None
            else int(duration) - 1
        ),
        interaction_matrix=imat,
        bad_atoms=bad,
        drives=list(drives.values()),
        collapse_ops=collapse_matrices(
            basis_data.eigenbasis,
            lindblad_data.local_collapse_ops,
            lindblad_data.depolarizing_pauli_2ds,
        ),
        qubit_ids=[str(q) for q in qids],
        slm_end=int(samples._slm_mask.end),
        slm_targets=slm_targets,
    )

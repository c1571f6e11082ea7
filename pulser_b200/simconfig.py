"""QuTiP-free ``SimConfig``: the deprecated configuration object of ``QutipEmulator``.

Mirror of ``pulser-simulation/pulser_simulation/simconfig.py:43-273`` (same fields, defaults, validation messages,
``from_noise_model`` / ``to_noise_model`` round trip, ``__str__`` layout), so that

    B200Emulator.from_sequence(seq, config=SimConfig(noise=("doppler", "amplitude")))

works where ``pulser_simulation`` cannot be imported (its ``simconfig`` module imports qutip, although it uses it only
as a type tag for ``eff_noise_opers``, ``simconfig.py:23,109,133,257``).  Effective-noise operators are accepted as
anything matrix-like: an object with ``.full()`` (``qutip.Qobj``, ``B200Operator``) or a 2-D array.

All validation is delegated to ``pulser.noise_model.NoiseModel`` exactly where the reference delegates it.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field, fields
from typing import Any, Union

import numpy as np

from . import _compat  # noqa: F401  (locates pulser-core)
from pulser._hamiltonian_data.hamiltonian_data import SUPPORTED_NOISES, doppler_sigma
from pulser.noise_model import _LEGACY_DEFAULTS, NoiseModel

# SimConfig name of the NoiseModel parameters that are called differently (simconfig.py:34-40)
_RENAMED = {"noise_types": "noise", "state_prep_error": "eta", "p_false_pos": "epsilon", "p_false_neg": "epsilon_prime"}


def _operator_matrix(op: Any) -> np.ndarray:
    """Dense matrix of an effective-noise operator; the two TypeErrors of ``simconfig.py:255-262``."""
    if hasattr(op, "full") and callable(op.full):
        mat = np.asarray(op.full())
    elif isinstance(op, (np.ndarray, list, tuple)):
        mat = np.asarray(op)
    else:
        raise TypeError(f"{op} is not a Qobj.")
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1]:
        raise TypeError("Operators are supposed to be of Qutip type 'oper'.")
    return mat


@dataclass(frozen=True)
class SimConfig:
    """Configuration of a simulation (deprecated since pulser v1.6 in favour of ``NoiseModel``).

    Fields and units as in the reference: ``temperature`` is given in uK and stored in K, ``laser_waist`` in um,
    rates in rad/us; ``runs`` noisy runs of ``samples_per_run`` samples each.
    """

    noise: Union[str, tuple[str, ...]] = ()
    runs: int = _LEGACY_DEFAULTS["runs"]
    samples_per_run: int = _LEGACY_DEFAULTS["samples_per_run"]
    temperature: float = _LEGACY_DEFAULTS["temperature"]
    laser_waist: float = _LEGACY_DEFAULTS["laser_waist"]
    amp_sigma: float = _LEGACY_DEFAULTS["amp_sigma"]
    detuning_sigma: float = 0.0
    eta: float = _LEGACY_DEFAULTS["state_prep_error"]
    epsilon: float = _LEGACY_DEFAULTS["p_false_pos"]
    epsilon_prime: float = _LEGACY_DEFAULTS["p_false_neg"]
    relaxation_rate: float = _LEGACY_DEFAULTS["relaxation_rate"]
    dephasing_rate: float = _LEGACY_DEFAULTS["dephasing_rate"]
    hyperfine_dephasing_rate: float = _LEGACY_DEFAULTS["hyperfine_dephasing_rate"]
    depolarizing_rate: float = _LEGACY_DEFAULTS["depolarizing_rate"]
    eff_noise_rates: list = field(default_factory=list, repr=False)
    eff_noise_opers: list = field(default_factory=list, repr=False)
    solver_options: Union[dict, None] = None

    # ------------------------------------------------------------------ construction
    def __post_init__(self) -> None:
        warnings.warn("'SimConfig' has been deprecated, please use `NoiseModel` instead.", DeprecationWarning,
                      stacklevel=2)
        if isinstance(self.noise, str):
            object.__setattr__(self, "noise", (self.noise,))
        if not isinstance(self.temperature, (int, float)):
            raise TypeError(f"'temperature' must be a float, not {type(self.temperature)}.")
        object.__setattr__(self, "temperature", self.temperature / 1e6)  # uK -> K
        NoiseModel._check_noise_types(self.noise)
        for name, value in self.spam_dict.items():
            if value > 1 or value < 0:
                raise ValueError(f"SPAM parameter {name} = {value} must be" + " greater than 0 and less than 1.")
        mats = [_operator_matrix(op) for op in self.eff_noise_opers]
        NoiseModel._check_eff_noise(self.eff_noise_rates, mats, "eff_noise" in self.noise, self.with_leakage)
        NoiseModel._validate_parameters({f.name: getattr(self, f.name) for f in fields(self)})

    @classmethod
    def from_noise_model(cls, noise_model: NoiseModel) -> "SimConfig":
        """The SimConfig equivalent to a NoiseModel (``simconfig.py:113-138``)."""
        relevant = NoiseModel._find_relevant_params(noise_model.noise_types, noise_model.state_prep_error,
                                                    noise_model.amp_sigma, noise_model.laser_waist)
        kwargs: dict[str, Any] = {"noise": noise_model.noise_types}
        for param in relevant:
            kwargs[_RENAMED.get(param, param)] = getattr(noise_model, param)
        if "amplitude" in noise_model.noise_types:
            kwargs.setdefault("laser_waist", float("inf"))  # None there means "no waist", not the legacy default
        kwargs.pop("with_leakage", None)
        if "eff_noise_opers" in kwargs:
            kwargs["eff_noise_opers"] = [np.asarray(op) for op in kwargs["eff_noise_opers"]]
        return cls(**kwargs)

    def to_noise_model(self) -> NoiseModel:
        """The NoiseModel equivalent to this configuration (``simconfig.py:140-160``)."""
        waist = None if math.isinf(self.laser_waist) else self.laser_waist
        relevant = NoiseModel._find_relevant_params(self.noise, self.eta, self.amp_sigma, waist)
        kwargs = {param: getattr(self, _RENAMED.get(param, param)) for param in relevant}
        if "temperature" in kwargs:
            kwargs["temperature"] *= 1e6  # back to uK
        if "eff_noise_opers" in kwargs:
            kwargs["eff_noise_opers"] = [_operator_matrix(op) for op in kwargs["eff_noise_opers"]]
        return NoiseModel(**kwargs)

    # ------------------------------------------------------------------ derived quantities
    @property
    def with_leakage(self) -> bool:
        return "leakage" in self.noise

    @property
    def spam_dict(self) -> dict[str, float]:
        return {"eta": self.eta, "epsilon": self.epsilon, "epsilon_prime": self.epsilon_prime}

    @property
    def doppler_sigma(self) -> float:
        """Standard deviation of the Doppler shift at the configured temperature (rad/us)."""
        return doppler_sigma(self.temperature)

    @property
    def supported_noises(self) -> dict:
        return SUPPORTED_NOISES

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, SimConfig):
            return NotImplemented
        for f in fields(self):
            a, b = getattr(self, f.name), getattr(other, f.name)
            if f.name == "eff_noise_opers":
                if len(a) != len(b) or any(not np.array_equal(_operator_matrix(x), _operator_matrix(y))
                                           for x, y in zip(a, b)):
                    return False
            elif a != b:
                return False
        return True

    __hash__ = None  # type: ignore[assignment]

    def __str__(self, solver_options: bool = False) -> str:
        rows = [("Number of runs:        ", self.runs, True), ("Samples per run:       ", self.samples_per_run, True),
                ("Noise types:           ", ", ".join(self.noise), bool(self.noise)),
                ("SPAM dictionary:       ", self.spam_dict, "SPAM" in self.noise),
                ("Effective noise rates:       ", self.eff_noise_rates, "eff_noise" in self.noise),
                ("Effective noise operators:       ", self.eff_noise_opers, "eff_noise" in self.noise),
                ("Temperature:           ", f"{self.temperature * 1.e6}µK", "doppler" in self.noise),
                ("Laser waist:           ", f"{self.laser_waist}μm", "amplitude" in self.noise),
                ("Amplitude standard dev.:  ", self.amp_sigma, "amplitude" in self.noise),
                ("Relaxation rate: ", self.relaxation_rate, "relaxation" in self.noise),
                ("Dephasing rate: ", f"{self.dephasing_rate} (Rydberg), {self.hyperfine_dephasing_rate} (Hyperfine)",
                 "dephasing" in self.noise),
                ("Depolarizing rate: ", self.depolarizing_rate, "depolarizing" in self.noise)]
        lines = ["Options:", "----------"] + [f"{label}{value}" for label, value, shown in rows if shown]
        if solver_options:
            lines.append("Solver Options: \n" + f"{str(self.solver_options)[10:-1]}")
        return "\n".join(lines).rstrip()

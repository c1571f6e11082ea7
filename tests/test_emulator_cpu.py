"""CPU tests of the ``B200Emulator`` facade (needs pulser-core; the device plan
is replaced by the oracle-backed fake of tests/fake_device.py, so only the host
logic is under test here)."""
import warnings
from collections import Counter

import numpy as np
import pytest

from pulser_b200 import HAVE_PULSER

pytestmark = pytest.mark.skipif(not HAVE_PULSER, reason="pulser-core not importable here")


@pytest.fixture
def emu(monkeypatch):
    from fake_device import FakeDevicePlan, FakeLindbladPlan
    from pulser_b200 import emulator, engine, lindblad

    monkeypatch.setattr(engine, "DevicePlan", FakeDevicePlan)
    monkeypatch.setattr(lindblad, "LindbladPlan", FakeLindbladPlan)
    return emulator


def _seq(n_side=2, duration=400, device=None):
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice

    seq = Sequence(Register.square(n_side, spacing=6.0, prefix="q"), device or MockDevice)
    seq.declare_channel("ch", "rydberg_global")
    seq.add(Pulse.ConstantPulse(duration, 2 * np.pi, np.pi, 0), "ch")
    return seq


def test_constructor_messages(emu):
    """same checks and messages as simulation.py:130-230, 955-1034"""
    from pulser import Register, Sequence
    from pulser.devices import MockDevice

    with pytest.raises(TypeError, match="sequence has to be a valid"):
        emu.B200Emulator.from_sequence({"pulse": "fake"})
    empty = Sequence(Register.square(2, prefix="q"), MockDevice)
    with pytest.raises(ValueError, match="no declared channels"):
        emu.B200Emulator.from_sequence(empty)
    empty.declare_channel("ch", "rydberg_global")
    with pytest.raises(ValueError, match="No instructions given"):
        emu.B200Emulator.from_sequence(empty)
    with pytest.raises(ValueError, match="must be greater than 0"):
        emu.B200Emulator.from_sequence(_seq(), sampling_rate=0.0)
    with pytest.raises(ValueError, match="too small, less than 4 data points"):
        emu.B200Emulator.from_sequence(_seq(duration=16), sampling_rate=0.1)
    sim = emu.B200Emulator.from_sequence(_seq())
    with pytest.raises(ValueError, match="Incompatible shape of initial state"):
        sim.set_initial_state(np.ones(3))
    with pytest.raises(ValueError, match="Wrong evaluation time label"):
        sim.set_evaluation_times("Sometimes")
    with pytest.raises(ValueError, match="extends further than sequence duration"):
        sim.set_evaluation_times([0.1, 9.0])
    with pytest.raises(TypeError, match="Unknown solver options"):
        sim.run(not_an_option=1)


def test_evaluation_times_and_properties(emu):
    sim = emu.B200Emulator.from_sequence(_seq(), evaluation_times="Minimal")
    np.testing.assert_allclose(sim.evaluation_times, [0.0, 0.4])
    assert sim.basis_name == "ground-rydberg" and sim.dim == 2 and sim.total_duration_ns == 400
    assert len(sim.sampling_times) == 401
    sim.set_evaluation_times(0.5)
    assert len(sim.evaluation_times) == 200 and sim.evaluation_times[-1] == 0.4
    sim.set_evaluation_times([0.1, 0.25])
    np.testing.assert_allclose(sim.evaluation_times, [0.0, 0.1, 0.25, 0.4])
    assert sim.initial_state.full()[-1, 0] == 1.0  # all-ground = last basis vector (r, g ordering)


def test_coherent_run_wraps_states(emu, capsys):
    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    sim = emu.B200Emulator.from_sequence(_seq(), evaluation_times=[0.2])
    res = sim.run(print_progress=True, max_step=1e-3, nsteps=5000)  # QuTiP option names are accepted
    assert "Emulating Trajectory 1/1" in capsys.readouterr().out
    assert len(res) == 3 and res._basis_name == "ground-rydberg"
    ref = evolve.sesolve(OracleHamiltonian.from_spec(sim._current_spec), evolve.all_ground_state(sim._current_spec),
                         [0, 0.4], rtol=1e-10, atol=1e-12)[-1]
    np.testing.assert_allclose(res.states[-1].full().ravel(), ref, atol=1e-7)
    fs = res.get_final_state()
    assert abs(fs.full()[np.argmax(np.abs(fs.full()))].imag) < 1e-12  # ignore_global_phase
    np.random.seed(3)
    c = res.sample_final_state(500)
    assert sum(c.values()) == 500 and all(len(k) == 4 for k in c)
    n0 = np.diag(1.0 - ((np.arange(16) >> 3) & 1))  # |r><r| on qubit 0
    occ = res.expect([n0])[0]
    assert occ.shape == (3,) and abs(occ[0]) < 1e-12 and 0 < occ[-1] < 1


def test_noisy_run_trajectories(emu, capsys):
    from fake_device import FakeDevicePlan
    from pulser import NoiseModel

    np.random.seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nm = NoiseModel(temperature=50.0, amp_sigma=0.05, laser_waist=175.0, samples_per_run=7)
        sim = emu.B200Emulator.from_sequence(_seq(duration=200), noise_model=nm, n_trajectories=5,
                                             evaluation_times="Minimal")
        FakeDevicePlan.calls = 0
        res = sim.run(print_progress=True)
    out = capsys.readouterr().out
    assert "Emulating Trajectory 1/5" in out and "Emulating Trajectory 5/5" in out
    assert FakeDevicePlan.calls == 1  # all trajectories evolved as ONE device batch
    assert type(res).__name__ == "NoisyResults" and res.n_measures == 35
    assert all(sum(r.bitstring_counts.values()) == 35 for r in res)
    assert res[0].bitstring_counts == {"0000": 35}  # t = 0: everything in the ground state
    with pytest.raises(ValueError, match="'n_trajectories' must be defined"):
        emu.B200Emulator.from_sequence(_seq(), noise_model=nm)


def test_spam_bad_atoms_are_merged_with_reps(emu):
    from pulser import NoiseModel

    np.random.seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nm = NoiseModel(state_prep_error=0.3, p_false_pos=0.0, p_false_neg=0.0, samples_per_run=1)
        sim = emu.B200Emulator.from_sequence(_seq(n_side=1, duration=100), noise_model=nm, n_trajectories=40,
                                             evaluation_times="Minimal")
        reps = [r for _, r in sim._specs]
        res = sim.run()
    assert sum(reps) == 40 and len(reps) <= 2  # hamiltonian_data.py:795-835 merges identical patterns
    assert sum(res[-1].bitstring_counts.values()) == 40


def test_collapse_operators_use_master_equation(emu):
    from pulser import NoiseModel

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nm = NoiseModel(dephasing_rate=0.4, relaxation_rate=0.2)
        sim = emu.B200Emulator.from_sequence(_seq(n_side=1, duration=200), noise_model=nm, evaluation_times="Minimal")
        res = sim.run()
    rho = res.states[-1]
    assert not rho.isket and rho.shape == (2, 2)
    assert abs(rho.tr() - 1) < 1e-7 and np.trace(rho.full() @ rho.full()).real < 1.0
    w = res[-1]._weights()
    np.testing.assert_allclose(w, np.abs(rho.diag())[::-1] / np.sum(np.abs(rho.diag())))
    r_proj = np.diag([1.0, 0.0])
    assert 0 < res.expect([r_proj])[0][-1] < 1


def test_large_register_collapse_needs_trajectories(emu):
    """Density matrix of 14 qubits does not fit: wave-function Monte Carlo is used and asks for n_trajectories."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import MockDevice

    reg = Register.from_coordinates([(8.0 * (i % 7), 8.0 * (i // 7)) for i in range(14)], prefix="q")
    seq = Sequence(reg, MockDevice)
    seq.declare_channel("ch", "rydberg_global")
    seq.add(Pulse.ConstantPulse(100, 1.0, 0.0, 0), "ch")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(dephasing_rate=0.1), evaluation_times="Minimal")
    assert sim._use_mcwf() and not sim._density_matrix_fits()
    with pytest.raises(ValueError, match="'n_trajectories' must be defined"):
        sim.run()


def test_run_xy(emu):
    """reference tests/pulser_simulation/test_simulation.py:1493-1530 (test_run_xy): XY mode through the facade."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice

    seq = Sequence(Register.from_coordinates([[10, 0], [0, 0]], prefix="atom"), MockDevice)
    seq.declare_channel("ch0", "mw_global")
    seq.add(Pulse.ConstantPulse(1500, 3.0, 1.0, 0.0), "ch0")
    sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01)
    assert sim.basis_name == "XY" and sim.dim == 2
    good = np.r_[1, np.zeros(3)]
    sim.set_initial_state(good)
    np.testing.assert_array_equal(sim.initial_state.full().ravel(), good)  # all-|u> is the default
    res = sim.run()
    final = res.get_final_state().full().ravel()
    assert abs(np.linalg.norm(final) - 1) < 1e-8 and abs(final[0]) < 1.0
    # exchange symmetry of the two-atom register: |ud> and |du> amplitudes agree
    assert abs(final[1] - final[2]) < 1e-8
    assert not sim.samples_obj._measurement
    seq.measure(basis="XY")
    sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01)
    res = sim.run()
    assert sim.samples_obj._measurement == "XY"
    counts = res.sample_final_state(500)
    assert sum(counts.values()) == 500 and set(counts) <= {"00", "01", "10", "11"}


def test_run_xy_slm_mask_equals_removed_qubit(emu):
    """reference tests/pulser_simulation/test_simulation.py:1748-1790 (test_mask_equals_remove_xy), taken to the
    evolved state: while the SLM mask covers the whole sequence the masked atom neither interacts nor is driven,
    so the three-atom run factorises into the two-atom run and the untouched atom.  Exercises the spec extraction
    of the mask (slm_end / slm_targets / coefficient) through the facade."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice

    pulse = Pulse.ConstantPulse(100, 10, 0, 0)
    seq3 = Sequence(Register({"q0": (0, 0), "q1": (10, 10), "q2": (-10, -10)}), MockDevice)
    seq3.set_magnetic_field(0, 1.0, 0.0)
    seq3.declare_channel("ch", "mw_global")
    seq3.config_slm_mask(["q2"])
    seq3.add(pulse, "ch")
    seq2 = Sequence(Register({"q0": (0, 0), "q1": (10, 10)}), MockDevice)
    seq2.set_magnetic_field(0, 1.0, 0.0)
    seq2.declare_channel("ch", "mw_global")
    seq2.add(pulse, "ch")
    sim3 = emu.B200Emulator.from_sequence(seq3, evaluation_times="Minimal")
    sim2 = emu.B200Emulator.from_sequence(seq2, evaluation_times="Minimal")
    f3 = sim3.run().get_final_state().full().ravel()
    f2 = sim2.run().get_final_state().full().ravel()
    # default initial state all-|u> (digit 0, simulation.py:498-505): the masked atom q2 never leaves it
    np.testing.assert_allclose(f3.reshape(4, 2)[:, 0], f2, atol=1e-7)
    assert np.max(np.abs(f3.reshape(4, 2)[:, 1])) < 1e-9


def test_build_operator_mirrors_reference(emu):
    """reference hamiltonian.py:145-229 / test_simulation.py:383-450 (test_building_basis_and_projection_operators,
    test_build_operator_exceptions): tensor products, 'global' sums, ids or labels, error messages."""
    sim = emu.B200Emulator.from_sequence(_seq(), evaluation_times="Minimal")
    ops = sim.op_matrix
    assert set(ops) == {"I", "sigma_rr", "sigma_rg", "sigma_gr", "sigma_gg"}
    r, g = np.array([1.0, 0.0]), np.array([0.0, 1.0])
    np.testing.assert_array_equal(ops["sigma_gr"].toarray(), np.outer(g, r))
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    eye = np.eye(2)
    xiii = sim.build_operator([(x, ["q0"])]).toarray()
    np.testing.assert_array_equal(xiii, np.kron(np.kron(np.kron(x, eye), eye), eye))
    zz = np.diag([1.0, -1.0])
    got = sim.build_operator([(zz, ["q1", "q2"]), ("sigma_rr", ["q3"])]).toarray()
    np.testing.assert_array_equal(got, np.kron(np.kron(np.kron(eye, zz), zz), np.outer(r, r)))
    glob = sim.build_operator([("sigma_rr", "global")]).toarray()
    occ = sum(np.kron(np.kron(np.eye(2**k), np.outer(r, r)), np.eye(2 ** (3 - k))) for k in range(4))
    np.testing.assert_array_equal(glob, occ)
    with pytest.raises(ValueError, match="Duplicate atom"):
        sim.build_operator([("sigma_gg", ["q0", "q0"])])
    with pytest.raises(ValueError, match="not a valid operator"):
        sim.build_operator([("wrong", ["q0"])])
    with pytest.raises(ValueError, match="Invalid qubit names"):
        sim.build_operator([("sigma_gg", ["wrong"])])
    # usable as an observable of the results (qutip.expect replacement)
    res = sim.run()
    val = res.expect([sim.build_operator([("sigma_rr", "global")])])[0]
    final = res.get_final_state().full().ravel()
    assert abs(val[-1] - np.vdot(final, glob @ final).real) < 1e-12


class _DuckConfig:
    """What the deprecated SimConfig entry points need (pulser_simulation.simconfig imports QuTiP)."""

    def __init__(self, noise_model):
        from pulser._hamiltonian_data.hamiltonian_data import SUPPORTED_NOISES

        self._nm = noise_model
        self.noise = tuple(noise_model.noise_types)
        self.supported_noises = SUPPORTED_NOISES

    def to_noise_model(self):
        return self._nm


def test_deprecated_config_entry_points(emu):
    """reference simulation.py:348-477 (set_config / add_config / reset_config) and their tests
    (test_simulation.py:1305-1427): deprecation warnings, messages, the noise model is replaced / merged."""
    from pulser import NoiseModel

    sim = emu.B200Emulator.from_sequence(_seq(), evaluation_times="Minimal")
    with pytest.warns(DeprecationWarning, match="Supplying a 'SimConfig'"):
        with pytest.raises(ValueError, match="is not a valid `SimConfig`"):
            sim.set_config("bad_config")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim.set_config(_DuckConfig(NoiseModel(dephasing_rate=0.3)))
        assert sim.noise_model.noise_types == ("dephasing",) and sim.noise_model.dephasing_rate == 0.3
        assert sim._has_collapse_ops()
        # add_config keeps the parameters already set and adds the new noise types
        sim.add_config(_DuckConfig(NoiseModel(dephasing_rate=0.9, relaxation_rate=0.2)))
        assert set(sim.noise_model.noise_types) == {"dephasing", "relaxation"}
        assert sim.noise_model.dephasing_rate == 0.3 and sim.noise_model.relaxation_rate == 0.2
        sim.reset_config()
        assert sim.noise_model == NoiseModel() and not sim._has_collapse_ops()
        # XY mode does not support relaxation (SUPPORTED_NOISES)
        from pulser import Pulse, Register, Sequence
        from pulser.devices import MockDevice

        seq = Sequence(Register.from_coordinates([[10, 0], [0, 0]], prefix="atom"), MockDevice)
        seq.declare_channel("ch0", "mw_global")
        seq.add(Pulse.ConstantPulse(200, 3.0, 1.0, 0.0), "ch0")
        xy = emu.B200Emulator.from_sequence(seq)
        with pytest.raises(NotImplementedError, match="Interaction mode 'XY' does not support simulation of noise types"):
            xy.set_config(_DuckConfig(NoiseModel(relaxation_rate=0.3)))
    with pytest.raises(ImportError, match="matplotlib"):
        sim.draw()


def test_amp_sigma_noise_reaches_the_spec(emu):
    """reference tests/pulser_simulation/test_simulation.py:2193-2266 (test_amp_sigma_noise): per-channel amplitude
    factors, constant from pulse to pulse, act on the per-qubit tables -- checked on the plain-array spec that
    feeds the CUDA path (coef = 0.5 amp exp(-i phase)), 'all' basis with two addressed bases."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import MockDevice

    seq = Sequence(Register({"q0": (0, 0), "q1": (10, 10)}), MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.declare_channel("ch1", "raman_local", initial_target="q0")
    seq.declare_channel("ch2", "raman_local", initial_target="q1")
    pulse1 = Pulse.ConstantPulse(120, 1, 0, 2.0)
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch1", protocol="no-delay")
    seq.target("q1", "ch1")
    seq.add(pulse1, "ch1", protocol="no-delay")
    seq.add(pulse1, "ch2", protocol="no-delay")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(amp_sigma=0.1), n_trajectories=1)
        clean = emu.B200Emulator.from_sequence(seq)
    noisy, base = sim._current_spec, clean._noiseless_spec()
    assert noisy.eigenbasis == ["r", "g", "h"] and [d.basis for d in noisy.drives] == [d.basis for d in base.drives]
    ryd = {d.basis: d for d in noisy.drives}["ground-rydberg"]
    ryd0 = {d.basis: d for d in base.drives}["ground-rydberg"]
    dig = {d.basis: d for d in noisy.drives}["digital"]
    dig0 = {d.basis: d for d in base.drives}["digital"]
    nz = np.flatnonzero(ryd0.coef[0])
    f0 = (ryd.coef[0, nz[0]] / ryd0.coef[0, nz[0]]).real  # factor of the global channel
    assert f0 > 0 and f0 != 1.0
    for q in range(2):  # same factor on both atoms and on both pulses; the phase survives
        np.testing.assert_allclose(ryd.coef[q], ryd0.coef[q] * f0, rtol=1e-15)
    np.testing.assert_array_equal(ryd.det, ryd0.det)
    # digital basis: q0 driven by ch1 only; q1 first by ch2 then by ch1
    nz0 = np.flatnonzero(dig0.coef[0])
    f1 = (dig.coef[0, nz0[0]] / dig0.coef[0, nz0[0]]).real
    np.testing.assert_allclose(dig.coef[0], dig0.coef[0] * f1, rtol=1e-15)
    nz1 = np.flatnonzero(dig0.coef[1])
    f2 = (dig.coef[1, nz1[0]] / dig0.coef[1, nz1[0]]).real
    assert len({f0, f1, f2}) == 3 and all(f > 0 and f != 1 for f in (f1, f2))
    expected = dig0.coef[1].copy()
    expected[: pulse1.duration] *= f2
    expected[-pulse1.duration - 1:] *= f1
    np.testing.assert_allclose(dig.coef[1], expected, rtol=1e-15)
    # identical rows (one factor for the global channel): the spec keeps the cheaper uniform-drive kernel
    assert ryd.uniform and not dig.uniform


def test_detuning_noise_reaches_the_spec(emu):
    """reference tests/pulser_simulation/test_simulation.py:2269-2317 (test_detuning_noise): the shot-to-shot detuning
    offsets of seed 1337, one per channel, land in the per-qubit detuning tables of the spec."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import MockDevice

    duration = 10
    np.random.seed(1337)
    seq = Sequence(Register({"q0": (0, 0), "q1": (10, 10)}), MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.declare_channel("ch1", "raman_local", initial_target="q0")
    seq.declare_channel("ch2", "raman_local", initial_target="q1")
    pulse1 = Pulse.ConstantPulse(duration, 0, 0, 0)
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch1", protocol="no-delay")
    seq.add(pulse1, "ch2", protocol="no-delay")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(detuning_sigma=0.1), n_trajectories=1)
    tabs = {d.basis: d for d in sim._current_spec.drives}
    ryd, dig = tabs["ground-rydberg"].det, tabs["digital"].det
    np.testing.assert_allclose(ryd[0], np.array([-0.04902824] * (2 * duration) + [0.0]), atol=1e-8)
    np.testing.assert_allclose(ryd[1], ryd[0], atol=0)
    np.testing.assert_allclose(dig[0], np.array([-0.17550787] * duration + [0.0] * (duration + 1)), atol=1e-8)
    np.testing.assert_allclose(dig[1], np.array([-0.20112646] * duration + [0.0] * (duration + 1)), atol=1e-8)


def test_run_from_sequence_and_from_samples_agree(emu):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:615-650 / test_simulation.py (from_sequence vs the
    constructor on sampled sequences): both entry points feed the same spec, hence the same state."""
    from pulser.sampler import sampler

    seq = _seq(duration=300)
    a = emu.B200Emulator.from_sequence(seq, evaluation_times="Minimal")
    b = emu.B200Emulator(sampler.sample(seq), seq.register, seq.device, evaluation_times="Minimal")
    sa, sb = a._current_spec, b._current_spec
    np.testing.assert_array_equal(sa.drives[0].coef, sb.drives[0].coef)
    np.testing.assert_array_equal(sa.drives[0].det, sb.drives[0].det)
    np.testing.assert_array_equal(sa.interaction_matrix, sb.interaction_matrix)
    fa = a.run().get_final_state().full()
    fb = b.run().get_final_state().full()
    np.testing.assert_allclose(fa, fb, rtol=1e-16, atol=0)


def test_eval_times_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:721-816 (test_eval_times), statement by statement."""
    seq = _seq()
    new = lambda: emu.B200Emulator.from_sequence(seq, sampling_rate=1.0)  # noqa: E731
    with pytest.raises(ValueError, match="evaluation_times float must be between 0 and 1."):
        new().set_evaluation_times(3.0)
    with pytest.raises(ValueError, match="Wrong evaluation time label."):
        new().set_evaluation_times(123)
    with pytest.raises(ValueError, match="Wrong evaluation time label."):
        new().set_evaluation_times("Best")
    sim = new()
    with pytest.raises(ValueError, match="Provided evaluation-time list contains negative values."):
        sim.set_evaluation_times([-1, 0, sim.sampling_times[-2]])
    with pytest.raises(ValueError, match="Provided evaluation-time list extends further than sequence duration."):
        sim.set_evaluation_times([0, sim.sampling_times[-1] + 10])
    sim.set_evaluation_times("Full")
    assert sim._eval_times_instruction == "Full"
    np.testing.assert_almost_equal(sim._eval_times_array, sim.sampling_times)
    sim.set_evaluation_times("Minimal")
    np.testing.assert_almost_equal(sim._eval_times_array, np.array([sim.sampling_times[0], sim._tot_duration / 1000]))
    sim.set_evaluation_times([0, sim.sampling_times[-3], sim._tot_duration / 1000])
    np.testing.assert_almost_equal(sim._eval_times_array,
                                   np.array([0, sim.sampling_times[-3], sim._tot_duration / 1000]))
    sim.set_evaluation_times([])
    np.testing.assert_almost_equal(sim._eval_times_array, np.array([0, sim._tot_duration / 1000]))
    sim.set_evaluation_times(0.0001)
    np.testing.assert_almost_equal(sim._eval_times_array, np.array([0, sim._tot_duration / 1000]))
    sim = new()
    sim.set_evaluation_times([sim.sampling_times[-10], sim.sampling_times[-3]])
    np.testing.assert_almost_equal(
        sim._eval_times_array,
        np.array([0, sim.sampling_times[-10], sim.sampling_times[-3], sim._tot_duration / 1000]))
    sim = new()
    sim.set_evaluation_times(0.4)
    np.testing.assert_almost_equal(
        sim.sampling_times[np.linspace(0, len(sim.sampling_times) - 1, int(0.4 * len(sim.sampling_times)), dtype=int)],
        sim._eval_times_array)


def test_empty_sequences_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:434-473 (test_empty_sequences): messages, and a
    sequence of delays only gives all-zero tables whatever the SPAM trajectory."""
    from pulser import NoiseModel, Register, Sequence
    from pulser.devices import MockDevice
    from pulser.sampler import sampler

    reg = Register({"control1": (-4, 0), "target": (0, 4), "control2": (4, 0)})
    seq = Sequence(reg, MockDevice)
    with pytest.raises(ValueError, match="no declared channels"):
        emu.B200Emulator.from_sequence(seq)
    seq.declare_channel("ch0", "mw_global")
    with pytest.raises(ValueError, match="No instructions given"):
        emu.B200Emulator.from_sequence(seq)
    with pytest.raises(ValueError, match="SequenceSamples is empty"):
        emu.B200Emulator(sampler.sample(seq), seq.register, seq.device)
    seq = Sequence(reg, MockDevice)
    seq.declare_channel("test", "raman_local", "target")
    seq.declare_channel("test2", "rydberg_global")
    with pytest.raises(ValueError, match="No instructions given"):
        emu.B200Emulator.from_sequence(seq)
    seq.delay(100, "test")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(
            seq, noise_model=NoiseModel(samples_per_run=1, state_prep_error=0.005, p_false_pos=0.01, p_false_neg=0.05),
            n_trajectories=15)
    for d in sim._current_spec.drives:
        np.testing.assert_equal(d.coef, 0)
        np.testing.assert_equal(d.det, 0)


def test_run_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:636-718 (test_run) without the QuTiP-only pieces: initial
    state validation, progress_bar validation, measurement basis, SPAM with a non-ground initial state."""
    from pulser import NoiseModel

    seq = _seq(duration=400)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01,
                                             noise_model=NoiseModel(p_false_pos=0.01, p_false_neg=0.05))
    n = sim._hamiltonian_data.n_qudits
    bad_initial = np.array([1.0])
    good_initial_array = np.r_[1, np.zeros(sim.dim**n - 1)]
    good_no_dims = np.zeros(sim.dim**n)
    good_no_dims[2] = 1.0
    with pytest.raises(ValueError, match="Incompatible shape of initial state"):
        sim.set_initial_state(bad_initial)
    sim.set_initial_state(good_initial_array)
    sim.run()
    sim.set_initial_state(good_no_dims)
    sim.run()
    seq.measure("ground-rydberg")
    sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01)
    sim.set_initial_state(good_no_dims)
    sim.run()
    assert sim.samples_obj._measurement == "ground-rydberg"
    sim.run(progress_bar=True)
    sim.run(progress_bar=False)
    sim.run(progress_bar=None)
    with pytest.raises(ValueError, match="`progress_bar` must be a bool."):
        sim.run(progress_bar=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01,
                                             noise_model=NoiseModel(samples_per_run=1, state_prep_error=0.1),
                                             n_trajectories=1)
    sim.set_initial_state(good_no_dims)
    with pytest.raises(NotImplementedError,
                       match="Can't combine state preparation errors with an initial state different from the ground."):
        sim.run()


def _ccz_sequence():
    """The fixture sequence of the reference's simulation tests (test_simulation.py:40-95): three atoms, local Raman
    pi_Y pulses then the CCZ on a local Rydberg channel -- 'all' basis, nine 1 us pulses."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import BlackmanWaveform

    reg = Register({"control1": np.array([-4.0, 0.0]), "target": np.array([0.0, 4.0]),
                    "control2": np.array([4.0, 0.0])})
    duration = 1000
    pi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(duration, np.pi), 0.0, 0)
    twopi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(duration, 2 * np.pi), 0.0, 0)
    pi_y = Pulse.ConstantDetuning(BlackmanWaveform(duration, np.pi), 0.0, -np.pi / 2)
    seq = Sequence(reg, DigitalAnalogDevice)
    seq.declare_channel("raman", "raman_local", "control1")
    seq.add(pi_y, "raman")
    seq.target("target", "raman")
    seq.add(pi_y, "raman")
    seq.target("control2", "raman")
    seq.add(pi_y, "raman")
    seq.declare_channel("ryd", "rydberg_local", "control1")
    seq.add(pi_pulse, "ryd", protocol="wait-for-all")
    seq.target("control2", "ryd")
    seq.add(pi_pulse, "ryd")
    seq.target("target", "ryd")
    seq.add(twopi_pulse, "ryd")
    seq.target("control2", "ryd")
    seq.add(pi_pulse, "ryd")
    seq.target("control1", "ryd")
    seq.add(pi_pulse, "ryd")
    seq.add(Pulse.ConstantPulse(duration, 1, 0, 0), "ryd")
    return seq


def test_initialization_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:111-222 (test_initialization_and_construction_of_
    hamiltonian) without the QuTiP type checks."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice, MockDevice
    from pulser.register.register_layout import RegisterLayout
    from pulser.sampler import sampler

    seq = _ccz_sequence()
    fake_sequence = {"pulse1": "fake", "pulse2": "fake"}
    with pytest.raises(TypeError, match="sequence has to be a valid"):
        emu.B200Emulator.from_sequence(fake_sequence)
    with pytest.raises(TypeError, match="sequence has to be a valid"):
        emu.B200Emulator(fake_sequence, Register.square(2, prefix="q"), MockDevice)
    with pytest.raises(ValueError, match="The ids of qubits targeted in Local channels"):
        emu.B200Emulator(sampler.sample(seq),
                         Register({"target": np.array([0.0, 0.0]), "control2": np.array([1.0, 0.0])}), MockDevice)
    with pytest.raises(ValueError, match="'noise_model' and 'config' cannot both be provided"):
        emu.B200Emulator.from_sequence(seq, config=_DuckConfig(NoiseModel()), noise_model=NoiseModel())
    with pytest.raises(ValueError, match="'n_trajectories' must be defined when the NoiseModel contains"
                                         " stochastic noise"):
        emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(amp_sigma=0.1))
    sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.011)
    sampled_seq = sampler.sample(seq)
    ext = sampled_seq.extend_duration(sampled_seq.max_duration + 1)
    for ch in sampled_seq.channels:
        for qty in ("amp", "det", "phase"):
            assert np.all(np.equal(getattr(sim.samples_obj.channel_samples[ch], qty),
                                   getattr(ext.channel_samples[ch], qty)))
    assert sim._current_spec.n_qudits == len(seq.qubit_info)
    assert sim._tot_duration == 9000  # seq has 9 pulses of 1 us
    assert sim._current_spec.qubit_ids == ["control1", "target", "control2"]
    with pytest.raises(ValueError, match="too small, less than"):
        emu.B200Emulator.from_sequence(seq, sampling_rate=0.0001)
    with pytest.raises(ValueError, match="`sampling_rate`"):
        emu.B200Emulator.from_sequence(seq, sampling_rate=5)
    with pytest.raises(ValueError, match="`sampling_rate`"):
        emu.B200Emulator.from_sequence(seq, sampling_rate=-1)
    assert sim._sampling_rate == 0.011
    assert len(sim.sampling_times) == int(sim._sampling_rate * sim._tot_duration)
    with pytest.warns(UserWarning, match="returns a copy of itself"):
        seq_copy = seq.build()
    x = seq_copy.declare_variable("x")
    seq_copy.add(Pulse.ConstantPulse(x, 1, 0, 0), "ryd")
    with pytest.raises(ValueError, match="needs to be built"):
        emu.B200Emulator.from_sequence(seq_copy)
    mapp_reg = RegisterLayout([[0, 0], [10, 10]]).make_mappable_register(1)
    with pytest.raises(ValueError, match="needs to be built"):
        emu.B200Emulator.from_sequence(Sequence(mapp_reg, DigitalAnalogDevice))


def test_extraction_of_sequences_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:225-251 (test_extraction_of_sequences): every pulse of
    every local channel shows up in the per-qubit tables of the spec, as coef = 0.5 amp exp(-i phase) and det."""
    from pulser import Pulse

    seq = _ccz_sequence()
    sim = emu.B200Emulator.from_sequence(seq)
    spec = sim._current_spec
    tabs = {d.basis: d for d in spec.drives}
    qidx = {q: i for i, q in enumerate(spec.qubit_ids)}
    for channel in seq.declared_channels:
        basis = seq.declared_channels[channel].basis
        assert seq.declared_channels[channel].addressing == "Local"
        for slot in seq._schedule[channel]:
            if isinstance(slot.type, Pulse):
                for qubit in slot.targets:
                    k = qidx[qubit]
                    amp = slot.type.amplitude.samples
                    np.testing.assert_allclose(tabs[basis].coef[k, slot.ti:slot.tf],
                                               0.5 * amp * np.exp(-1j * float(slot.type.phase)), rtol=1e-15, atol=0)
                    np.testing.assert_array_equal(tabs[basis].det[k, slot.ti:slot.tf], slot.type.detuning.samples)


def test_noise_port(emu, capfd):
    """reference tests/pulser_simulation/test_simulation.py:891-953 (test_noise): SPAM trajectories of seed 3 are merged
    into the same groups (progress lines identical) and the reference's Counter comes out exactly; the depolarizing
    channel is refused in the 3-level basis with the reference's message; unprepared atoms get all-zero tables."""
    from pulser import NoiseModel

    seq = _ccz_sequence()
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim2 = emu.B200Emulator.from_sequence(
            seq, sampling_rate=0.01,
            noise_model=NoiseModel(samples_per_run=5, p_false_pos=0.01, p_false_neg=0.05, state_prep_error=0.9),
            n_trajectories=15)
        # one trajectory per device batch: the shots are then drawn trajectory by trajectory, the order in which the
        # reference consumes np.random (a batch samples all its trajectories at one evaluation time before it moves
        # on, which is the same distribution drawn in another order)
        counts = sim2.run(print_progress=True, b200_batch=1).sample_final_state()
    out, _ = capfd.readouterr()
    assert out.rstrip("\n").split("\n") == [
        "Emulating Trajectories [1 - 13]/15",
        "Emulating Trajectory 14/15",
        "Emulating Trajectory 15/15",
    ]
    # the reference's hard-coded Counter (real QuTiP sesolve runs of the 3-level CCZ sequence), shot for shot
    assert counts == Counter({"000": 824, "100": 41, "101": 57, "001": 63, "010": 15})
    with pytest.raises(NotImplementedError, match="Cannot include"):
        emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(depolarizing_rate=0.05))
    pending = sim2._pending_trajectories()  # a second use redraws the trajectories (simulation.py:892-902)
    bad = sim2._hamiltonian_data.noise_trajectories[0].trajectory.bad_atoms
    assert any(bad.values())
    spec = pending[0][0]
    for d in spec.drives:
        for q, is_bad in enumerate(bad.values()):
            if is_bad:
                assert np.all(d.coef[q] == 0.0) and np.all(d.det[q] == 0.0)


@pytest.mark.parametrize("noise,result", [
    (("dephasing",), {"0": 572, "1": 428}),
    (("depolarizing",), {"0": 561, "1": 439}),
    (("dephasing", "depolarizing", "relaxation"), {"0": 562, "1": 438}),
    (("eff_noise", "leakage"), {"0": 572, "1": 428}),
])
def test_noises_rydberg_port(emu, noise, result):
    """reference tests/pulser_simulation/test_simulation.py:978-1046 (test_noises_rydberg) through the facade, seed and
    all: np.random.seed(123), build, run, sample_final_state() == the reference's hard-coded Counter.  Exact equality
    means the facade draws from np.random exactly where QutipEmulator does, builds the same collapse operators and
    (with the oracle-backed plan standing in for the device) evolves to the same populations as QuTiP's mesolve."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.noise_model import _LEGACY_DEFAULTS

    np.random.seed(123)
    seq = Sequence(Register.from_coordinates([(0, 0)], prefix="q"), DigitalAnalogDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.add(Pulse.ConstantPulse(2500, np.pi, 0, 0), "ch0")
    params = {
        p: _LEGACY_DEFAULTS[p]
        for p in NoiseModel._find_relevant_params(
            [n for n in noise if n not in ["leakage", "eff_noise"]],
            state_prep_error=_LEGACY_DEFAULTS["state_prep_error"],
            amp_sigma=_LEGACY_DEFAULTS["amp_sigma"],
            laser_waist=_LEGACY_DEFAULTS["laser_waist"],
        )
    }
    with_leakage = "leakage" in noise
    if with_leakage or "eff_noise" in noise:
        params["eff_noise_opers"] = [np.diag([1.0, 0, 0]).astype(complex) if with_leakage
                                     else np.diag([1.0, -1.0]).astype(complex)]
        params["eff_noise_rates"] = [0.1 if with_leakage else 0.025]
    n_trajectories = params.pop("runs", None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01,
                                             noise_model=NoiseModel(with_leakage=with_leakage, **params),
                                             n_trajectories=n_trajectories)
        assert set(sim.noise_model.noise_types) == set(noise)
        res = sim.run()
    assert res.sample_final_state() == Counter(result)
    rho = res.states[-1].full()
    trace_2 = np.trace(rho @ rho).real
    assert trace_2 < 1 and not np.isclose(trace_2, 1)
    if with_leakage:
        state = res.get_final_state().full()
        assert np.allclose(state[2, :], 0) and np.allclose(state[:, 2], 0)


@pytest.mark.parametrize("noisychannel", [True, False])
def test_get_final_state_port(emu, noisychannel):
    """reference tests/pulser_simulation/test_simresults.py:162-241 (test_get_final_state)."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import BlackmanWaveform
    from pulser_b200.results import CoherentResults

    reg = Register({"A": np.array([0.0, 0.0]), "B": np.array([0.0, 10.0])})
    pi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0)
    seq_no_meas = Sequence(reg, DigitalAnalogDevice)
    seq_no_meas.declare_channel("ryd", "rydberg_global")
    seq_no_meas.add(pi_pulse, "ryd")
    seq_no_meas.measure("ground-rydberg")
    np.random.seed(123)
    sim = emu.B200Emulator.from_sequence(seq_no_meas, evaluation_times=0.05)
    results = sim.run()
    if noisychannel:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = emu.B200Emulator(sim.samples_obj, register=seq_no_meas.register,
                                   device=seq_no_meas.device, noise_model=NoiseModel(dephasing_rate=0.01),
                                   evaluation_times=0.05)
    _results = sim.run()
    assert isinstance(_results, CoherentResults)
    final_state = _results.get_final_state()
    assert (not final_state.isket) if noisychannel else final_state.isket
    with pytest.raises(TypeError, match="Can't reduce"):
        _results.get_final_state(reduce_to_basis="digital")
    np.testing.assert_allclose(  # qutip.Qobj.__eq__ compares within qutip.settings.core["atol"] = 1e-12
        _results.get_final_state(reduce_to_basis="ground-rydberg", ignore_global_phase=False).full(),
        _results.states[-1].tidyup().full(), rtol=0, atol=1e-12)
    assert np.all(np.isclose(np.abs(_results.get_final_state(ignore_global_phase=False).full()),
                             np.abs(_results.states[-1].full())))
    assert np.all(np.isclose(np.abs(_results.get_final_state(ignore_global_phase=True).full()),
                             np.abs(_results.states[-1].full())))
    seq_ = Sequence(reg, DigitalAnalogDevice)
    seq_.declare_channel("ryd", "rydberg_global")
    seq_.declare_channel("ram", "raman_local", initial_target="A")
    seq_.add(pi_pulse, "ram")
    seq_.add(pi_pulse, "ram")
    seq_.add(pi_pulse, "ryd")
    results_ = emu.B200Emulator.from_sequence(seq_, evaluation_times="Minimal").run()
    with pytest.raises(ValueError, match="'reduce_to_basis' must be"):
        results_.get_final_state(reduce_to_basis="all")
    with pytest.raises(TypeError, match="Can't reduce to chosen basis"):
        results_.get_final_state(reduce_to_basis="digital")
    h_states = results_.get_final_state(reduce_to_basis="digital", tol=1, normalize=False).full()[1:]
    assert np.linalg.norm(h_states) < 3e-6
    assert np.all(np.isclose(np.abs(results_.get_final_state(reduce_to_basis="ground-rydberg").full()),
                             np.abs(results.states[-1].full()), atol=1e-5))


def test_expect_port(emu):
    """reference tests/pulser_simulation/test_simresults.py:289-380 (test_expect): messages, pi pulse, SPAM pseudo-
    density, and the QuTiP-derived number 0.7804005 (atol 1e-6) of a leakage run with the collapse operator |x><g|."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import BlackmanWaveform
    from pulser_b200.results import CoherentResults

    reg = Register({"A": np.array([0.0, 0.0]), "B": np.array([0.0, 10.0])})
    pi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0)
    seq = Sequence(reg, DigitalAnalogDevice)
    seq.declare_channel("ryd", "rydberg_global")
    seq.add(pi_pulse, "ryd")
    seq.measure("ground-rydberg")
    results = emu.B200Emulator.from_sequence(seq, evaluation_times="Minimal").run()
    with pytest.raises(TypeError, match="must be a list"):
        results.expect("bad_observable")
    with pytest.raises(TypeError, match="Incompatible type"):
        results.expect(["bad_observable"])
    with pytest.raises(ValueError, match="Incompatible shape"):
        results.expect([np.array(3)])
    seq_single = Sequence(Register.from_coordinates([(0, 0)], prefix="q"), DigitalAnalogDevice)
    seq_single.declare_channel("ryd", "rydberg_global")
    seq_single.add(pi_pulse, "ryd")
    proj = lambda d, k: np.diag((np.arange(d) == k).astype(complex))  # noqa: E731
    op = [proj(2, 0)]
    results_single = emu.B200Emulator.from_sequence(seq_single, evaluation_times=0.05).run()
    exp = results_single.expect(op)[0]
    assert np.isclose(exp[-1], 1)
    np.testing.assert_almost_equal(np.diag(results_single._calc_pseudo_density_diag(-1)), np.array([[1, 0], [0, 0]]))
    noise_model = NoiseModel(p_false_pos=0.01, p_false_neg=0.05)
    sim_single = emu.B200Emulator.from_sequence(seq_single, noise_model=noise_model)
    sim_single.set_evaluation_times("Minimal")
    results_single = sim_single.run()
    exp = results_single.expect(op)[0]
    assert len(exp) == 2
    assert isinstance(results_single, CoherentResults)
    assert results_single._meas_errors == {"epsilon": noise_model.p_false_pos,
                                           "epsilon_prime": noise_model.p_false_neg}
    assert np.isclose(exp[0], noise_model.p_false_pos)
    assert np.isclose(exp[-1], 1 - noise_model.p_false_neg)
    np.testing.assert_almost_equal(np.diag(results_single._calc_pseudo_density_diag(-1)),
                                   np.array([[1 - noise_model.p_false_neg, 0], [0, noise_model.p_false_neg]]))
    # with leakage: collapse operator |x><g| (basis(3, 2) @ basis(3, 1).dag()) at rate 0.5
    eff_op = np.zeros((3, 3), dtype=complex)
    eff_op[2, 1] = 1.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim_single = emu.B200Emulator.from_sequence(
            seq_single, noise_model=NoiseModel(eff_noise_rates=[0.5], eff_noise_opers=[eff_op], with_leakage=True),
            sampling_rate=0.1)
    sim_single.set_evaluation_times(0.5)
    res = sim_single.run()
    assert isinstance(res, CoherentResults)
    assert np.isclose(res.expect([proj(3, 0)])[0][-1], 0.7804005, atol=1e-6)  # value hard-coded in the reference
    seq3dim = Sequence(reg, DigitalAnalogDevice)
    seq3dim.declare_channel("ryd", "rydberg_global")
    seq3dim.declare_channel("ram", "raman_local", initial_target="A")
    seq3dim.add(pi_pulse, "ram")
    seq3dim.add(pi_pulse, "ryd")
    exp3dim = emu.B200Emulator.from_sequence(seq3dim, evaluation_times="Minimal").run().expect(
        [np.kron(proj(3, 0), np.eye(3))])
    assert abs(exp3dim[0][-1]) < 1e-9  # reference: 1.9e-14


def test_concurrent_pulses_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:1402-1427 (test_concurrent_pulses): a local and a global
    Rydberg pulse at the same time add up; doppler noise changes detunings, never the drive element H[0, 1]."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice

    seq = Sequence(Register({"q0": (0, 0)}), DigitalAnalogDevice)
    seq.declare_channel("ch_local", "rydberg_local", initial_target="q0")
    seq.declare_channel("ch_global", "rydberg_global")
    pulse = Pulse.ConstantPulse(20, 10, 0, 0)
    seq.add(pulse, "ch_local")
    seq.add(pulse, "ch_global", protocol="no-delay")
    sim_no_noise = emu.B200Emulator.from_sequence(seq)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim_with_noise = emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(samples_per_run=5, temperature=50.0),
                                                        n_trajectories=15)
    for t in sim_no_noise.evaluation_times:
        ham_no_noise = sim_no_noise.get_hamiltonian(t * 1000)
        ham_with_noise = sim_with_noise.get_hamiltonian(t * 1000)
        assert ham_no_noise[0, 1] == ham_with_noise[0, 1]
    assert abs(sim_no_noise.get_hamiltonian(10)[0, 1] - 10.0) < 1e-12  # 0.5 * (10 + 10): both channels drive q0


def test_mask_nopulses_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:1730-1745 (test_mask_nopulses)."""
    from pulser import Register, Sequence
    from pulser.devices import MockDevice
    from pulser.sampler import sampler

    reg = Register({"q0": (0, 0), "q1": (10, 10), "q2": (-10, -10)})
    for channel_type in ["mw_global", "rydberg_global"]:
        seq_empty = Sequence(reg, MockDevice)
        if channel_type == "mw_global":
            seq_empty.set_magnetic_field(0, 1.0, 0.0)
        seq_empty.declare_channel("ch", channel_type)
        seq_empty.delay(duration=100, channel="ch")
        seq_empty.config_slm_mask(["q2"])
        sim_empty = emu.B200Emulator.from_sequence(seq_empty)
        assert seq_empty._slm_mask_time == []
        assert sampler.sample(seq_empty)._slm_mask.end == 0
        assert sim_empty.samples_obj._slm_mask.end == 0
        assert sim_empty._current_spec.slm_coefficient() is None  # no time-dependent interaction without a mask end


@pytest.mark.parametrize("noise", (
    {"detuning_sigma": 1.0}, {"amp_sigma": 1.0}, {"temperature": 10},
    {"temperature": 10, "disable_doppler": True, "trap_depth": 1000, "trap_waist": 0.1},
    {"detuning_hf_psd": [1.0, 2.0], "detuning_hf_omegas": [3.0, 4.0]},
))
def test_noisy_runs_port(emu, noise):
    """reference tests/pulser_simulation/test_simulation.py:2420-2466 (test_noisy_runs): every kind of stochastic noise
    gives one solve per trajectory and a NoisyResults."""
    from fake_device import FakeDevicePlan
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import MockDevice
    from pulser_b200.results import NoisyResults

    np.random.seed(1337)
    duration = 10
    seq = Sequence(Register({"q0": (0, 0), "q1": (10, 10)}), MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.declare_channel("ch1", "raman_local", initial_target="q0")
    seq.declare_channel("ch2", "raman_local", initial_target="q1")
    pulse1 = Pulse.ConstantPulse(duration, 0, 0, 0)
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch1", protocol="no-delay")
    seq.add(pulse1, "ch2", protocol="no-delay")
    nruns = 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, noise_model=NoiseModel(**noise), n_trajectories=nruns)
        result = sim.run()
    assert isinstance(result, NoisyResults)
    assert len(sim._pending_trajectories()) == nruns  # one trajectory per run (they share one device batch here)
    del FakeDevicePlan


@pytest.mark.parametrize("noise_data, expected", [
    (dict(noise_types="doppler"), True),
    (dict(noise_types="amplitude", amp_sigma=1.0), True),
    (dict(noise_types="amplitude", amp_sigma=0.0), False),
    (dict(noise_types="detuning"), True),
    (dict(noise_types="register"), True),
    (dict(noise_types="SPAM", state_prep_error=0.0), False),
    (dict(noise_types="SPAM", state_prep_error=1.0), True),
    (dict(noise_types="other"), False),
    (dict(noise_types={"other", "detuning"}), True),
])
def test_has_stochastic_noise_port(emu, noise_data, expected):
    """reference tests/pulser_simulation/test_simulation.py:2469-2499 (test_has_stochastic_noise)."""
    from types import SimpleNamespace

    assert emu._has_stochastic_noise(SimpleNamespace(**noise_data)) is expected


def test_invalid_solver_error_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:2577-2591 (test_qutip_invalid_solver_error)."""
    from pulser import NoiseModel

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError, match="'fakesolver' is not a valid Solver"):
            emu.B200Emulator.from_sequence(_seq(), noise_model=NoiseModel(detuning_sigma=0.1), solver="fakesolver",
                                           n_trajectories=1)


def test_get_hamiltonian_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:476-588 (test_get_hamiltonian), seeds and numbers as they
    are: the doppler (seed 123, 20000 uK) and register-noise (seed 456) Hamiltonians at t = 144 ns hard-coded in the
    reference come out of the facade (get_hamiltonian = H(t) applied to the basis vectors by the plan)."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import RampWaveform

    simple_reg = Register.from_coordinates([[10, 0], [0, 0]], prefix="atom")
    detun = 1.0
    simple_seq = Sequence(simple_reg, DigitalAnalogDevice)
    simple_seq.declare_channel("ising", "rydberg_global")
    simple_seq.add(Pulse.ConstantDetuning(RampWaveform(1500, 0.0, 2.0), detun, 0.0), "ising")
    simple_sim = emu.B200Emulator.from_sequence(simple_seq, sampling_rate=0.01)
    with pytest.raises(ValueError, match="less than or equal to"):
        simple_sim.get_hamiltonian(1650)
    with pytest.raises(ValueError, match="greater than or equal to"):
        simple_sim.get_hamiltonian(-10)
    simple_ham = simple_sim.get_hamiltonian(143)
    assert np.isclose(simple_ham[0, 0], DigitalAnalogDevice.interaction_coeff / 10**6 - 2 * detun)
    np.random.seed(123)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim_noise = emu.B200Emulator.from_sequence(simple_seq, noise_model=NoiseModel(samples_per_run=1, temperature=20000),
                                                   n_trajectories=15)
    g = 0.09606404
    np.testing.assert_allclose(sim_noise.get_hamiltonian(144), np.array([
        [4.47984523, g, g, 0.0], [g, 12.03082372, 0.0, g], [g, 0.0, -12.97113702, g], [0.0, g, g, 0.0]], dtype=complex),
        rtol=1e-7, atol=1e-8)
    expected_noiseless = emu.B200Emulator.from_sequence(simple_seq).get_hamiltonian(144)
    np.testing.assert_allclose(sim_noise.get_hamiltonian(144, noiseless=True), expected_noiseless)
    np.random.seed(456)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim_noise = emu.B200Emulator.from_sequence(
            simple_seq, noise_model=NoiseModel(samples_per_run=1, temperature=50.0, trap_depth=150.0, trap_waist=1.0),
            n_trajectories=1)
    np.testing.assert_allclose(sim_noise.get_hamiltonian(144), np.array([
        [4.92294305, g, g, 0.0], [g, -0.59902269, 0.0, g], [g, 0.0, -0.70099956, g], [0.0, g, g, 0.0]], dtype=complex),
        rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(sim_noise.get_hamiltonian(144, noiseless=True), expected_noiseless)


def test_results_xy_port(emu):
    """reference tests/pulser_simulation/test_simresults.py:486-527 (test_results_xy)."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice
    from pulser.waveforms import BlackmanWaveform

    reg = Register({"A": np.array([0.0, 0.0]), "B": np.array([0.0, 10.0])})
    seq_ = Sequence(reg, MockDevice)
    seq_.declare_channel("ch0", "mw_global")
    seq_.add(Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0), "ch0")
    seq_.measure("XY")
    results_ = emu.B200Emulator.from_sequence(seq_, evaluation_times="Minimal").run()
    assert results_._dim == 2 and results_._size == 2
    assert results_._basis_name == "XY" and results_._meas_basis == "XY"
    np.testing.assert_array_equal(results_.states[0].full().ravel(), np.r_[1, 0, 0, 0])  # |uu>
    for basis in ("all", "ground-rydberg", "digital"):
        with pytest.raises(TypeError, match="Can't reduce a system in"):
            results_.get_final_state(reduce_to_basis=basis)
    state = results_.get_final_state(reduce_to_basis="XY")
    assert np.all(np.isclose(np.abs(state.full()), np.abs(results_.states[-1].full()), atol=1e-5))


def test_false_positive_port(emu):
    """reference tests/pulser_simulation/test_simresults.py:530-550 (test_false_positive): a pulse padded with long
    zero stretches must still act (the final state differs from the initial one)."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import AnalogDevice
    from pulser.waveforms import BlackmanWaveform, CompositeWaveform, ConstantWaveform

    seq = Sequence(Register.square(2, 5, prefix="q"), AnalogDevice)
    seq.declare_channel("ryd_glob", "rydberg_global")
    seq.add(Pulse.ConstantDetuning(CompositeWaveform(ConstantWaveform(2500, 0.0), BlackmanWaveform(1000, np.pi),
                                                     ConstantWaveform(500, 0.0)), 0, 0), channel="ryd_glob")
    sim = emu.B200Emulator.from_sequence(seq, evaluation_times="Minimal")
    final = sim.run().get_final_state().full()
    assert np.max(np.abs(final - sim.initial_state.full())) > 0.1


def test_sample_final_state_ports(emu):
    """reference tests/pulser_simulation/test_simresults.py:398-446 (test_sim_without_measurement,
    test_sample_final_state, test_sample_final_state_three_level)."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import BlackmanWaveform

    reg = Register({"A": np.array([0.0, 0.0]), "B": np.array([0.0, 10.0])})
    pi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0)

    def seq_no_meas():
        seq = Sequence(reg, DigitalAnalogDevice)
        seq.declare_channel("ryd", "rydberg_global")
        seq.add(pi_pulse, "ryd")
        return seq

    seq = seq_no_meas()
    assert not seq.is_measured()
    results_no_meas = emu.B200Emulator.from_sequence(seq, evaluation_times="Minimal").run()
    np.random.seed(123)
    assert results_no_meas.sample_final_state(1) == Counter({"11": 1})
    seq.measure("ground-rydberg")
    np.random.seed(123)
    results = emu.B200Emulator.from_sequence(seq, evaluation_times="Minimal").run()
    sampling = results.sample_final_state(1234)
    assert len(sampling) == 4  # all states were observed
    results[-1].matching_meas_basis = False
    assert results.sample_final_state(N_samples=911) == {"00": 911}
    seq3 = seq_no_meas()
    seq3.declare_channel("raman", "raman_local", "B")
    seq3.add(pi_pulse, "raman")
    res_3level = emu.B200Emulator.from_sequence(seq3, evaluation_times="Minimal").run()
    assert len(res_3level.sample_final_state()) == 2  # the Raman pi pulse on one atom leaves the other alone
    seq3.measure("ground-rydberg")
    res_3level_gb = emu.B200Emulator.from_sequence(seq3, evaluation_times="Minimal").run()
    assert len(res_3level_gb.sample_final_state()) == 4  # the global Rydberg pulse affects both


def test_mask_local_channel_port(emu):
    """reference tests/pulser_simulation/test_simulation.py:1841-1925 (test_mask_local_channel), on the spec: in Ising
    mode the SLM mask is a detuning of -10 x amplitude on the masked atoms (local table) under the global drive; a
    local Raman channel on q0 with phase pi shows up in the digital table."""
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice

    seq_ = Sequence(Register.square(2, prefix="q"), MockDevice)
    seq_.declare_channel("rydberg_global", "rydberg_global")
    pulse = Pulse.ConstantPulse(1000, 10, 0, 0)
    seq_.config_slm_mask(["q0", "q3"])
    seq_.add(pulse, "rydberg_global")
    seq_.declare_channel("raman_local", "raman_local", initial_target="q0")
    pulse2 = Pulse.ConstantPulse(1000, 10, -5, np.pi)
    seq_.add(pulse2, "raman_local", protocol="no-delay")
    assert seq_._slm_mask_time == [0, 1000] and seq_._slm_mask_targets == {"q0", "q3"}
    sim = emu.B200Emulator.from_sequence(seq_)
    spec = sim._current_spec
    assert spec.slm_coefficient() is None  # only XY mode has a time-dependent interaction
    tabs = {d.basis: d for d in spec.drives}
    ryd, dig = tabs["ground-rydberg"], tabs["digital"]
    amp = np.concatenate((pulse.amplitude.samples, [0]))
    for k, q in enumerate(["q0", "q1", "q2", "q3"]):
        np.testing.assert_array_equal(ryd.coef[k], 0.5 * amp)           # global drive, phase 0, on every atom
        np.testing.assert_array_equal(ryd.det[k], -10 * amp if q in ("q0", "q3") else 0 * amp)
    amp2 = np.concatenate((pulse2.amplitude.samples, [0]))
    det2 = np.concatenate((pulse2.detuning.samples, [0]))
    np.testing.assert_allclose(dig.coef[0], 0.5 * amp2 * np.exp(-1j * np.pi), rtol=1e-15, atol=1e-15)
    np.testing.assert_array_equal(dig.det[0], det2)
    assert np.all(dig.coef[1:] == 0) and np.all(dig.det[1:] == 0)



@pytest.mark.parametrize("three_d", [False, True])
def test_hamiltonian_constructor_arguments_port(three_d):
    """reference tests/pulser_simulation/test_hamiltonian.py:29-80 (test_register_2d / test_register_3d, issue #940):
    ``spec_from_pulser`` takes the exact arguments of the reference's ``Hamiltonian(noisy_samples, trajectory,
    basis_data, lindblad_data, sampling_rate)`` for 2D and 3D registers."""
    from pulser import Pulse, Register, Register3D, Sequence
    from pulser._hamiltonian_data import HamiltonianData
    from pulser.devices import MockDevice
    from pulser_b200.spec import spec_from_pulser

    reg = (Register3D({"q0": np.array([-4.0, 0.0, 0.0]), "q1": np.array([0.0, 4.0, 0.0])}) if three_d
           else Register({"q0": np.array([-4.0, 0.0]), "q1": np.array([0.0, 4.0])}))
    seq = Sequence(reg, MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.declare_channel("ch1", "raman_local", initial_target="q0")
    seq.declare_channel("ch2", "raman_local", initial_target="q1")
    pulse1 = Pulse.ConstantPulse(10, 0, 0, 0)
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch0")
    seq.add(pulse1, "ch1", protocol="no-delay")
    seq.add(pulse1, "ch2", protocol="no-delay")
    data = HamiltonianData.from_sequence(seq)
    for traj, noisy_samples, _ in data.noisy_samples:
        spec = spec_from_pulser(noisy_samples, traj, data.basis_data, data.lindblad_data, 0.5)
        assert spec.n_qudits == 2 and len(spec.sampling_times) == int(0.5 * noisy_samples.max_duration)
        assert spec.dim == data.basis_data.dim and spec.eigenbasis == list(data.basis_data.eigenbasis)
        assert spec.interaction_matrix.shape == (1, 2, 2)
        # (pulser rounds distances to its coordinate precision: 1e-6 relative)
        np.testing.assert_allclose(spec.interaction_matrix[0, 0, 1], MockDevice.interaction_coeff / (4 * np.sqrt(2)) ** 6,
                                   rtol=1e-6)


@pytest.mark.parametrize("leakage", [False, True])
def test_building_basis_and_projection_operators_port(emu, leakage):
    """reference tests/pulser_simulation/test_simulation.py:254-431 (test_building_basis_and_projection_operators):
    basis names, dimensions, basis kets and sigma_ab projectors of every addressing (all / ground-rydberg global and
    local / digital / XY), with and without the leakage level; build_operator messages."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice, MockDevice
    from pulser.sampler import sampler
    from pulser.waveforms import BlackmanWaveform

    def noise_model(dim):
        if not leakage:
            return NoiseModel()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return NoiseModel(with_leakage=True, eff_noise_opers=[np.eye(dim)], eff_noise_rates=[0.0])

    def check(sim, name, letters, pairs):
        dim = len(letters)
        assert sim.basis_name == name + ("_with_error" if leakage else "")
        assert sim.dim == dim
        assert list(sim.basis) == list(letters)
        for k, s in enumerate(letters):
            np.testing.assert_array_equal(sim.basis[s].full().ravel(), np.eye(dim)[k])
        for a, b in pairs:
            want = np.outer(np.eye(dim)[letters.index(a)], np.eye(dim)[letters.index(b)])
            np.testing.assert_array_equal(sim.op_matrix["sigma_" + a + b].toarray(), want)

    seq = _ccz_sequence()
    reg = Register({"control1": np.array([-4.0, 0.0]), "target": np.array([0.0, 4.0]), "control2": np.array([4.0, 0.0])})
    x = ("x",) if leakage else ()
    sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01, noise_model=noise_model(3 + leakage))
    check(sim, "all", ("r", "g", "h") + x, [("r", "r"), ("g", "r"), ("h", "g")] + ([("x", "r")] if leakage else []))
    with pytest.raises(ValueError, match="Duplicate atom"):
        sim.build_operator([("sigma_gg", ["target", "target"])])
    with pytest.raises(ValueError, match="not a valid operator"):
        sim.build_operator([("wrong", ["target"])])
    with pytest.raises(ValueError, match="Invalid qubit names: {'wrong'}"):
        sim.build_operator([("sigma_gg", ["wrong"])])
    op_standard = sim.build_operator([("sigma_gg", ["target"])])
    op_one = sim.build_operator(("sigma_gg", ["target"]))
    assert abs(op_standard - op_one).max() < 1e-10
    pi_pls = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0)
    dim = 2 + leakage

    def one_channel(device, *decl):
        s = Sequence(reg, device)
        s.declare_channel(*decl)
        s.add(pi_pls, decl[0])
        return s

    sim2 = emu.B200Emulator.from_sequence(one_channel(DigitalAnalogDevice, "global", "rydberg_global"),
                                          sampling_rate=0.01, noise_model=noise_model(dim))
    check(sim2, "ground-rydberg", ("r", "g") + x, [("r", "r"), ("g", "r")] + ([("x", "r")] if leakage else []))
    sim2b = emu.B200Emulator.from_sequence(one_channel(DigitalAnalogDevice, "local", "raman_local", "target"),
                                           sampling_rate=0.01, noise_model=noise_model(dim))
    check(sim2b, "digital", ("g", "h") + x, [("g", "g"), ("h", "g")] + ([("x", "h")] if leakage else []))
    sim2c = emu.B200Emulator.from_sequence(one_channel(DigitalAnalogDevice, "local_ryd", "rydberg_local", "target"),
                                           sampling_rate=0.01, noise_model=noise_model(dim))
    check(sim2c, "ground-rydberg", ("r", "g") + x, [("r", "r"), ("g", "r")] + ([("x", "g")] if leakage else []))
    seq_xy = one_channel(MockDevice, "global", "mw_global")
    with pytest.raises(ValueError, match="Bases used in samples should be supported by device."):
        emu.B200Emulator(sampler.sample(seq_xy), seq_xy.register, DigitalAnalogDevice)
    sim_xy = emu.B200Emulator.from_sequence(seq_xy, sampling_rate=0.01, noise_model=noise_model(dim))
    check(sim_xy, "XY", ("u", "d") + x, [("u", "u"), ("d", "u"), ("u", "d")] + ([("u", "x")] if leakage else []))

"""CPU tests of the V2 plugin (``B200Backend``) with the oracle-backed fake device."""
import warnings

import numpy as np
import pytest

from pulser_b200 import HAVE_PULSER

pytestmark = pytest.mark.skipif(not HAVE_PULSER, reason="pulser-core not importable here")


@pytest.fixture
def backend(monkeypatch):
    from fake_device import FakeDevicePlan, FakeLindbladPlan
    from pulser_b200 import backend, engine, lindblad

    monkeypatch.setattr(engine, "DevicePlan", FakeDevicePlan)
    monkeypatch.setattr(lindblad, "LindbladPlan", FakeLindbladPlan)
    return backend


def _seq(n=2, duration=300):
    from pulser import Pulse, Register, Sequence
    from pulser.devices import MockDevice
    from pulser.waveforms import BlackmanWaveform

    reg = Register.from_coordinates([(7.0 * i, 0.0) for i in range(n)], prefix="q")
    seq = Sequence(reg, MockDevice)
    seq.declare_channel("ch", "rydberg_global")
    seq.add(Pulse.ConstantDetuning(BlackmanWaveform(duration, np.pi), 1.0, 0.0), "ch")
    return seq


def test_state_and_operator_protocol(backend):
    S, O = backend.B200State, backend.B200Operator
    eig = ("r", "g")
    st = S.from_state_amplitudes(eigenstates=eig, amplitudes={"rg": 1.0, "gr": 1.0j})
    assert st.n_qudits == 2 and abs(st.overlap(st) - 4.0) < 1e-12
    st = S(st.to_array() / np.sqrt(2), eigenstates=eig)
    assert st.probabilities() == pytest.approx({"rg": 0.5, "gr": 0.5})
    assert st.bitstring_probabilities() == pytest.approx({"10": 0.5, "01": 0.5})
    np.random.seed(0)
    c = st.sample(num_shots=200)
    assert sum(c.values()) == 200 and set(c) <= {"10", "01"}
    n0 = O.from_operator_repr(eigenstates=eig, n_qudits=2, operations=[(1.0, [({"rr": 1.0}, {0})])])
    assert n0.expect(st) == pytest.approx(0.5)
    assert (2.0 * n0 + n0).expect(st) == pytest.approx(1.5)
    assert (n0 @ n0).expect(st) == pytest.approx(0.5)
    rho = S(np.outer(st.to_array(), st.to_array().conj()), eigenstates=eig)
    assert n0.expect(rho) == pytest.approx(0.5) and rho.overlap(st) == pytest.approx(1.0)
    with pytest.raises(ValueError, match="eigenstates"):
        n0.expect(S(np.ones(4) / 2, eigenstates=("g", "h")))


def test_backend_noiseless_observables(backend):
    from pulser.backend.default_observables import BitStrings, Energy, Occupation, StateResult

    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    seq = _seq()
    cfg = backend.B200Config(observables=[
        BitStrings(evaluation_times=[1.0], num_shots=100), StateResult(evaluation_times=[1.0]),
        Occupation(evaluation_times=[0.5, 1.0]), Energy(evaluation_times=[0.0, 0.5, 1.0])])
    np.random.seed(2)
    res = backend.B200Backend(seq, config=cfg).run()
    sim = backend.B200Emulator.from_sequence(seq)
    spec = sim._noiseless_spec()
    H = OracleHamiltonian.from_spec(spec)
    psi0 = evolve.all_ground_state(spec)
    states = evolve.sesolve(H, psi0, [0.0, 0.15, 0.3], rtol=1e-10, atol=1e-12)
    final = res.final_state.to_array()
    np.testing.assert_allclose(final, states[-1] / np.linalg.norm(states[-1]), atol=1e-7)
    occ = res.get_result("occupation", 1.0)
    ref_occ = [np.sum(np.abs(states[-1]) ** 2 * (1 - ((np.arange(4) >> (1 - k)) & 1))) for k in range(2)]
    np.testing.assert_allclose(np.asarray(occ, dtype=float), ref_occ, atol=1e-7)
    # energy self-consistency (reference tests/pulser_simulation/test_qutip_backend_v2.py:111-154)
    for t_rel, st in zip((0.0, 0.5, 1.0), states):
        e_ref = np.vdot(st, H.matrix_at(t_rel * 0.3) @ st).real / np.vdot(st, st).real
        assert float(np.real(res.get_result("energy", t_rel))) == pytest.approx(e_ref, abs=1e-6)
    assert sum(res.final_bitstrings.values()) == 100


def test_backend_stochastic_noise_aggregates(backend):
    import pulser
    from pulser.backend.default_observables import BitStrings, Occupation

    np.random.seed(4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nm = pulser.NoiseModel(temperature=50.0, amp_sigma=0.05, laser_waist=175.0)
        cfg = backend.B200Config(observables=[BitStrings(evaluation_times=[1.0], num_shots=10),
                                              Occupation(evaluation_times=[1.0])],
                                 noise_model=nm, n_trajectories=6)
        res = backend.B200Backend(_seq(duration=200), config=cfg).run()
    assert sum(res.final_bitstrings.values()) == 60  # bag union over 6 trajectories
    occ = np.asarray(res.get_result("occupation", 1.0), dtype=float)
    assert occ.shape == (2,) and np.all((occ > 0) & (occ < 1))


def test_config_validation(backend):
    with pytest.raises(ValueError, match="sampling rate"):
        backend.B200Config(sampling_rate=0.0)
    with pytest.raises(ValueError, match="Invalid solver"):
        backend.B200Config(solver="rk4")
    with pytest.raises(TypeError, match="must be an instance of `B200State`"):
        backend.B200Config(initial_state=np.ones(4))


def test_streaming_observables_match_host_formulas(backend, monkeypatch):
    """Noiseless runs hand the observables a device-resident state (``DeviceStateView``): every default observable
    of pulser/backend/default_observables.py must give what the plain matrix formulas give on the oracle states."""
    from pulser.backend.default_observables import (
        BitStrings, CorrelationMatrix, Energy, EnergySecondMoment, EnergyVariance, Expectation, Fidelity,
        Occupation, StateResult)

    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    seq = _seq(n=3, duration=240)
    eig = ("r", "g")
    S, O = backend.B200State, backend.B200Operator
    target = S.from_state_amplitudes(eigenstates=eig, amplitudes={"rgg": 1.0, "grg": 1.0, "ggr": 1.0})
    target = S(target.to_array() / np.sqrt(3), eigenstates=eig)
    sx0 = O.from_operator_repr(eigenstates=eig, n_qudits=3, operations=[(1.0, [({"rg": 1.0, "gr": 1.0}, {0})])])
    sx0.to_array()  # a general operator: its matrix is legitimately needed (host fallback); build it up front
    times = [0.25, 1.0]
    cfg = backend.B200Config(observables=[
        Occupation(evaluation_times=times), CorrelationMatrix(evaluation_times=times),
        Energy(evaluation_times=times), EnergyVariance(evaluation_times=times),
        EnergySecondMoment(evaluation_times=times), Fidelity(target, evaluation_times=times),
        Expectation(sx0, evaluation_times=times), BitStrings(evaluation_times=[1.0], num_shots=50),
        StateResult(evaluation_times=[1.0])])
    np.random.seed(5)
    built = []
    orig = backend.B200Operator._as_matrix
    monkeypatch.setattr(backend.B200Operator, "_as_matrix", staticmethod(lambda op: (built.append(1), orig(op))[1]))
    res = backend.B200Backend(seq, config=cfg).run()
    monkeypatch.undo()
    # number operators / identity were never turned into matrices during the run
    assert len(built) == 0
    sim = backend.B200Emulator.from_sequence(seq)
    spec = sim._noiseless_spec()
    H = OracleHamiltonian.from_spec(spec)
    T = spec.sampling_times[-1]
    states = evolve.sesolve(H, evolve.all_ground_state(spec), [0.0, 0.25 * T, T], rtol=1e-11, atol=1e-13)[1:]
    idx = np.arange(8)
    is_r = [((idx >> (2 - k)) & 1) == 0 for k in range(3)]
    for t_rel, st in zip(times, states):
        st = st / np.linalg.norm(st)
        p = np.abs(st) ** 2
        corr = np.array([[p[is_r[i] & is_r[j]].sum() for j in range(3)] for i in range(3)])
        np.testing.assert_allclose(np.asarray(res.get_result("occupation", t_rel), dtype=float), np.diag(corr), atol=1e-7)
        np.testing.assert_allclose(np.asarray(res.get_result("correlation_matrix", t_rel), dtype=float), corr, atol=1e-7)
        hm = H.matrix_at(t_rel * T)
        e, e2 = np.vdot(st, hm @ st).real, np.vdot(hm @ st, hm @ st).real
        assert float(np.real(res.get_result("energy", t_rel))) == pytest.approx(e, abs=1e-6)
        assert float(np.real(res.get_result("energy_second_moment", t_rel))) == pytest.approx(e2, abs=1e-5)
        assert float(np.real(res.get_result("energy_variance", t_rel))) == pytest.approx(e2 - e * e, abs=1e-5)
        assert float(res.get_result("fidelity", t_rel)) == pytest.approx(abs(np.vdot(target.to_array(), st)) ** 2, abs=1e-7)
        sx = np.vdot(st, sx0.to_array() @ st).real
        assert float(np.real(res.get_result("expectation", t_rel))) == pytest.approx(sx, abs=1e-7)
    assert sum(res.final_bitstrings.values()) == 50
    np.testing.assert_allclose(res.final_state.to_array(), states[-1] / np.linalg.norm(states[-1]), atol=1e-7)


def test_projector_patterns(backend):
    """Products / multiples of number operators keep their 'projector pattern' (what lets a device-resident state
    answer them with one reduction); everything else drops it."""
    O = backend.B200Operator
    eig = ("r", "g", "h")

    def num(i, letter="r", c=1.0):
        return O.from_operator_repr(eigenstates=eig, n_qudits=3, operations=[(c, [({letter * 2: 1.0}, {i})])])

    assert num(0)._pattern == (1.0, "r", frozenset({0}))
    assert (num(0) @ num(2))._pattern == (1.0, "r", frozenset({0, 2}))
    assert (2.0 * num(1, c=0.5))._pattern == (1.0, "r", frozenset({1}))
    assert (num(0) @ num(1, "h"))._pattern is None          # different eigenstates: not one projector
    assert (num(0) + num(1))._pattern is None
    ident = O.from_operator_repr(eigenstates=eig, n_qudits=3, operations=[(1.0, [])])
    assert ident._pattern == (1.0, None, frozenset())
    assert (ident @ num(1))._pattern == (1.0, "r", frozenset({1}))
    flip = O.from_operator_repr(eigenstates=eig, n_qudits=3, operations=[(1.0, [({"rg": 1.0}, {0})])])
    assert flip._pattern is None
    # lazily built matrices still agree with the eager formulas
    v = np.zeros(27, dtype=complex); v[0 * 9 + 1 * 3 + 0] = 1.0  # |r g r>
    st = backend.B200State(v, eigenstates=eig)
    assert (num(0) @ num(2)).expect(st) == pytest.approx(1.0) and (num(0) @ num(1)).expect(st) == pytest.approx(0.0)


def test_legacy_v1_backend(monkeypatch):
    """reference tests/pulser_simulation/test_qutip_backend.py:43-104 (test_qutip_backend, test_with_default_noise)
    on the V1 mirror: type check, deprecation warning, pi pulse on a local Raman channel, device noise model."""
    import dataclasses

    import pulser
    from fake_device import FakeDevicePlan, FakeLindbladPlan
    from pulser.devices import MockDevice
    from pulser.waveforms import BlackmanWaveform
    from pulser_b200 import backend, engine, lindblad
    from pulser_b200.results import CoherentResults, NoisyResults

    monkeypatch.setattr(engine, "DevicePlan", FakeDevicePlan)
    monkeypatch.setattr(lindblad, "LindbladPlan", FakeLindbladPlan)
    seq = pulser.Sequence(pulser.Register({"q0": (0, 0)}), MockDevice)
    seq.declare_channel("raman_local", "raman_local", initial_target="q0")
    seq.add(pulser.Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0, 0), "raman_local")
    with pytest.raises(TypeError, match="must be of type 'EmulatorConfig'"), pytest.deprecated_call(
            match="'QutipBackend' is deprecated"):
        backend.B200LegacyBackend(seq, pulser.NoiseModel())
    with pytest.deprecated_call(match="'QutipBackend' is deprecated"):
        be = backend.B200LegacyBackend(seq)
    results = be.run()
    assert isinstance(results, CoherentResults)
    np.testing.assert_allclose(results[0].get_state().full().ravel(), [1, 0])
    np.testing.assert_allclose(np.abs(results.get_final_state().full().ravel()), [0, 1], atol=1e-5)
    with pytest.raises(TypeError, match="must be a real device"), pytest.deprecated_call():
        backend.B200LegacyBackend(seq, mimic_qpu=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spam = pulser.NoiseModel(p_false_pos=0.1, p_false_neg=0.05, state_prep_error=0.1, runs=10, samples_per_run=1)
        dev = dataclasses.replace(MockDevice, noise_model=spam)
        be = backend.B200LegacyBackend(seq.with_new_device(dev), config=pulser.EmulatorConfig(prefer_device_noise_model=True))
        noisy = be.run()
    assert isinstance(noisy, NoisyResults)
    assert be._sim_obj.noise_model == spam


def test_noisy_runs_stream_on_the_device_like_the_replay(backend, monkeypatch):
    """Stochastic-noise runs of the V2 backend: the trajectory-streaming path (device-resident state views, noiseless
    Hamiltonian through a device-to-device state copy) gives the observables the replay of stored states gave."""
    import pulser
    from pulser.backend.default_observables import CorrelationMatrix, Energy, EnergySecondMoment, Occupation

    def run(stream):
        np.random.seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nm = pulser.NoiseModel(temperature=50.0, amp_sigma=0.05, laser_waist=175.0)
            times = [0.5, 1.0]
            cfg = backend.B200Config(observables=[Occupation(evaluation_times=times), Energy(evaluation_times=times),
                                                  EnergySecondMoment(evaluation_times=times),
                                                  CorrelationMatrix(evaluation_times=times)],
                                     noise_model=nm, n_trajectories=4)
            be = backend.B200Backend(_seq(n=3, duration=200), config=cfg)
            if not stream:  # force the replay of stored states
                monkeypatch.setattr(be._sim_obj, "_has_collapse_ops", lambda: True)
                monkeypatch.setattr(be._sim_obj, "_use_mcwf", lambda: False)
                monkeypatch.setattr(be._sim_obj, "_check_supported", lambda: None)
            return be.run()

    streamed, replayed = run(True), run(False)
    for name in ("occupation", "energy", "energy_second_moment", "correlation_matrix"):
        for t in (0.5, 1.0):
            a = np.asarray(streamed.get_result(name, t), dtype=float)
            b = np.asarray(replayed.get_result(name, t), dtype=float)
            np.testing.assert_allclose(a, b, atol=1e-9, err_msg=f"{name} at {t}")


def test_rounding_error_eval_time_duplication_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:469-493: an evaluation time that differs from a grid
    point by a rounding error must not be duplicated in the results."""
    import pulser
    from pulser.backend.default_observables import BitStrings

    seq = pulser.Sequence(pulser.Register.square(1, prefix="q"), pulser.AnalogDevice)
    seq.declare_channel("rydberg_global", "rydberg_global")
    seq.add(pulser.Pulse.ConstantPulse(200, 1, 0, 0), "rydberg_global")
    evaluation_times = np.linspace(0.0, 1.0, 201)
    mod = float(evaluation_times[98]) - 1e-16 * 0.5
    config = backend.B200Config(observables=[
        BitStrings(evaluation_times=evaluation_times, num_shots=5),
        BitStrings(evaluation_times=[0.49 - 6e-17], tag_suffix="mod", num_shots=5),
    ])
    res = backend.B200Backend(seq, config=config).run()
    assert len(res.get_result_times("bitstrings")) == 201
    assert len(res.get_result_times("bitstrings_mod")) == 1
    del mod


def test_run_twice_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:565-587 (test_run_twice): a second run() redraws the
    noise trajectories, so the aggregated state differs."""
    import pulser
    from pulser.backend.default_observables import StateResult

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noise_model = pulser.NoiseModel(trap_depth=1.0, trap_waist=1.0, temperature=50.0, disable_doppler=True,
                                        detuning_sigma=5.0)
        cfg = backend.B200Config(default_evaluation_times=[1.0], observables=[StateResult(evaluation_times=[1.0])],
                                 noise_model=noise_model, n_trajectories=4)
        be = backend.B200Backend(_seq(n=2, duration=200), config=cfg)
        np.random.seed(5)
        r1 = be.run()
        r2 = be.run()
    a1, a2 = r1.final_state.to_array(), r2.final_state.to_array()
    ov = abs(np.vdot(a1.ravel(), a2.ravel())) / (np.linalg.norm(a1) * np.linalg.norm(a2))
    assert ov != pytest.approx(1.0)


def _ref_sequence(device=None, scale=1):
    """The ``sequence()`` helper of the reference's V2 backend tests (test_qutip_backend_v2.py:57-88): two atoms at
    the blockade radius, rise / detuning sweep / fall.  ``scale`` shortens the three pulses (oracle-backed CPU runs)."""
    import math

    import pulser

    omega_max = 4 * 2 * math.pi
    u = omega_max / 2
    delta_0, delta_f = -6 * u, 2 * u
    t_rise, t_fall = 500 // scale, 1000 // scale
    t_sweep = int((delta_f - delta_0) / (2 * np.pi * 10) * 1000) // scale
    r = pulser.devices.MockDevice.rydberg_blockade_radius(u)
    reg = pulser.Register.rectangle(1, 2, r, prefix="q")
    seq = pulser.Sequence(reg, device if device is not None else pulser.devices.MockDevice)
    seq.declare_channel("ising_global", "rydberg_global")
    seq.add(pulser.Pulse.ConstantDetuning(pulser.waveforms.RampWaveform(t_rise, 0.0, omega_max), delta_0, 0.0),
            "ising_global")
    seq.add(pulser.Pulse.ConstantAmplitude(omega_max, pulser.waveforms.RampWaveform(t_sweep, delta_0, delta_f), 0.0),
            "ising_global")
    seq.add(pulser.Pulse.ConstantDetuning(pulser.waveforms.RampWaveform(t_fall, omega_max, 0.0), delta_f, 0.0),
            "ising_global")
    return seq


def test_callback_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:91-108 (test_callback): a callback forces the full
    evaluation grid -- one call per sample, noiseless and with a stochastic noise trajectory."""
    import pulser
    from pulser.backend.observable import Callback

    class CountCalls(Callback):
        def __init__(self):
            super().__init__()
            self.counter = 0

        def __call__(self, **kwargs):
            self.counter += 1

    seq = _ref_sequence(scale=10)
    be = backend.B200Backend(seq, config=backend.B200Config(callbacks=[CountCalls()]))
    be.run()
    assert be._config.callbacks[0].counter == seq.get_duration() + 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg = backend.B200Config(callbacks=[CountCalls()], noise_model=pulser.NoiseModel(amp_sigma=0.1), n_trajectories=1)
    be = backend.B200Backend(seq, config=cfg)
    be.run()
    assert be._config.callbacks[0].counter == seq.get_duration() + 1


def test_energy_port(backend, capfd):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:111-154 (test_qutip_backend_v2_energy)."""
    from pulser.backend.default_observables import Energy, StateResult

    seq = _ref_sequence(scale=10)
    with pytest.raises(TypeError, match="'config' must be an instance of 'EmulationConfig'"):
        backend.B200Backend(seq, config="tralala")
    T = seq.get_duration()
    config = backend.B200Config(default_evaluation_times="Full",
                                observables=[StateResult(), Energy(evaluation_times=[n / T for n in range(0, T + 1, 2)])],
                                print_progress=True)
    be = backend.B200Backend(seq, config=config)
    results = be.run()
    assert results.get_result_times("state") != results.get_result_times("energy")
    out, _ = capfd.readouterr()
    assert out == "Emulating Trajectory 1/1\n"
    assert results.get_result("energy", 0.0) == results.energy[0] == pytest.approx(0.0)
    sim = be._sim_obj
    # <psi|H(t)|psi> from the dense Hamiltonian of the facade and the stored state (qutip.expect in the reference)
    t_mid = results.get_result_times("energy")[len(results.energy) // 2]
    psi_mid = results.get_result("state", t_mid).to_array()
    h_mid = sim.get_hamiltonian(t_mid * T)
    assert results.get_result("energy", t_mid) == pytest.approx(np.vdot(psi_mid, h_mid @ psi_mid).real, rel=1e-5)
    psi_end = results.state[-1].to_array()
    h_end = sim.get_hamiltonian(T)
    assert results.get_result("energy", 1.0) == results.energy[-1]
    assert results.energy[-1] == pytest.approx(np.vdot(psi_end, h_end @ psi_end).real, rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("print_progress", [True, False])
def test_default_noise_model_port(backend, capfd, print_progress):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:157-196: prefer_device_noise_model takes the device's
    noise model for the emulation while the config keeps its own; progress lines of the two trajectories."""
    import dataclasses

    import pulser
    from pulser.backend.default_observables import StateResult

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noisy_device = dataclasses.replace(pulser.devices.MockDevice,
                                           noise_model=pulser.NoiseModel(dephasing_rate=0.01, temperature=50))
        config = backend.B200Config(
            observables=[StateResult(evaluation_times=[1.0])],
            noise_model=pulser.NoiseModel(p_false_neg=0.1),
            prefer_device_noise_model=True,
            initial_state=backend.B200State(np.array([1.0, 0, 0, 0], dtype=complex), eigenstates=("r", "g")),
            n_trajectories=2,
            print_progress=print_progress,
        )
        be = backend.B200Backend(_ref_sequence(noisy_device, scale=10), config=config)
        assert be._sim_obj._hamiltonian_data.noise_model.p_false_neg == 0.0
        assert be._sim_obj._hamiltonian_data.noise_model.temperature == 50
        assert be._sim_obj._hamiltonian_data.noise_model.dephasing_rate == 0.01
        assert be._config.noise_model.p_false_neg == 0.1
        be.run()
    out, _ = capfd.readouterr()
    assert out == ("Emulating Trajectory 1/2\nEmulating Trajectory 2/2\n" if print_progress else "")


def test_eval_times_rounding_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:257-284 (test_qutip_backend_v2_eval_times_rounding):
    the legacy evaluation times must never exceed the duration through a rounding error; a plain EmulationConfig
    is accepted and every requested time is stored exactly once."""
    import pulser

    n_points = 25
    for duration in range(400, 600, 28):
        seq = pulser.Sequence(pulser.Register({"q0": (-5, 0), "q1": (5, 0)}), pulser.AnalogDevice)
        seq.declare_channel("rydberg_global", "rydberg_global")
        seq.add(pulser.Pulse(pulser.ConstantWaveform(duration, np.pi), pulser.ConstantWaveform(duration, 0.0), 0),
                "rydberg_global")
        evaluation_times = np.linspace(0, 1, n_points).tolist()
        config = pulser.backend.EmulationConfig(observables=[pulser.backend.StateResult(evaluation_times=evaluation_times)])
        result = backend.B200Backend(seq, config=config).run().state
        assert len(result) == n_points


def test_stochastic_noise_v2_equals_legacy_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:199-254 (test_qutip_backend_v2_stochastic_noise):
    the V2 occupation of 30 noisy trajectories agrees with the legacy facade's within 0.03."""
    import pulser
    from pulser.backend.default_observables import Occupation, StateResult
    from pulser_b200 import emulator

    np.random.seed(123)

    def get_noise_model(samples_per_run):
        return pulser.NoiseModel(temperature=50.0, p_false_neg=0.01, amp_sigma=1e-3, samples_per_run=samples_per_run)

    seq = _ref_sequence(scale=25)
    T = seq.get_duration()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        config = backend.B200Config(default_evaluation_times=(1.0,),
                                    observables=[StateResult(evaluation_times=[1.0]),
                                                 Occupation(evaluation_times=[n / T for n in range(0, T + 1, 10)])],
                                    noise_model=get_noise_model(1), n_trajectories=30)
        be = backend.B200Backend(seq, config=config)
        assert be._sim_obj.n_trajectories == config.n_trajectories
        results = be.run()
        # (the reference keeps the legacy run on the full grid; its states are only read at the occupation times)
        old = emulator.B200Emulator.from_sequence(seq, noise_model=get_noise_model(100), n_trajectories=30,
                                                  evaluation_times=[int(n / T * T) * 1e-3 for n in range(0, T + 1, 10)])
        results_old_api = old.run()
    times = results.get_result_times("occupation")
    occupation = np.array([x[0] for x in results.occupation])
    indices = np.searchsorted(results_old_api._sim_times, np.array([int(t * T) * 1e-3 for t in times]))
    n0 = np.kron(np.diag([1.0, 0.0]), np.eye(2))
    occupation_old_api = results_old_api.expect([n0])[0][indices]
    assert np.max(np.abs(occupation - occupation_old_api)) < 0.03


@pytest.mark.parametrize("amp_sigma", [0.0, 1.0])
def test_leakage_port(backend, amp_sigma):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:287-360 (test_leakage): two far-apart atoms leaking
    from |r> and |g> into |x> at the same rate -- the leaked populations are analytic whatever the drive does."""
    import math

    import pulser
    from pulser.backend.default_observables import StateResult

    reg = pulser.Register.rectangle(1, 2, spacing=1000.0, prefix="q")
    seq = pulser.Sequence(reg, pulser.MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    duration = 500
    seq.add(pulser.Pulse.ConstantPulse(duration, np.pi, 0.0, 0.0), "ch0")
    basisx = np.array([0.0, 0.0, 1.0]).reshape(3, 1)
    basisg = np.array([0.0, 1.0, 0.0]).reshape(3, 1)
    basisr = np.array([1.0, 0.0, 0.0]).reshape(3, 1)
    rate = 0.5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noise_model = pulser.NoiseModel(eff_noise_rates=[rate, rate],
                                        eff_noise_opers=[basisx @ basisr.T, basisx @ basisg.T],
                                        with_leakage=True, amp_sigma=amp_sigma)
        cfg = backend.B200Config(default_evaluation_times=[1.0], observables=[StateResult(evaluation_times=[1.0])],
                                 noise_model=noise_model, solver=backend.Solver.MESOLVER, n_trajectories=1)
        result = backend.B200Backend(seq, config=cfg).run()
    eig = ("r", "g", "x")
    px = basisx @ basisx.T
    p_no = np.diag([1.0, 1.0, 0.0])
    both_leaked = backend.B200Operator(np.kron(px, px), eig)
    one_leaked = backend.B200Operator(np.kron(px, p_no), eig) + backend.B200Operator(np.kron(p_no, px), eig)
    no_leaked = backend.B200Operator(np.kron(p_no, p_no), eig)
    decay = math.exp(-rate * duration / 1000)
    assert one_leaked.expect(result.final_state) == pytest.approx(2 * (1 - decay) * decay, abs=1e-6)
    assert no_leaked.expect(result.final_state) == pytest.approx(decay**2, abs=1e-6)
    assert both_leaked.expect(result.final_state) == pytest.approx((1 - decay) ** 2, abs=1e-6)


def test_register_detuning_detection_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:363-393: register + detuning noise are stochastic,
    the aggregated final state of ten trajectories is a density matrix."""
    import pulser
    from pulser.backend.default_observables import StateResult

    reg = pulser.Register.rectangle(1, 2, spacing=1000.0, prefix="q")
    seq = pulser.Sequence(reg, pulser.MockDevice)
    seq.declare_channel("ch0", "rydberg_global")
    seq.add(pulser.Pulse.ConstantPulse(200, np.pi, 0.0, 0.0), "ch0")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noise_model = pulser.NoiseModel(trap_depth=1.0, trap_waist=1.0, temperature=50.0, disable_doppler=True,
                                        detuning_sigma=5.0)
        assert set(noise_model.noise_types) == {"register", "detuning"}
        cfg = backend.B200Config(default_evaluation_times=[1.0], observables=[StateResult(evaluation_times=[1.0])],
                                 noise_model=noise_model, n_trajectories=10)
        result = backend.B200Backend(seq, config=cfg).run()
    assert result.final_state._state.shape == (4, 4)  # density matrix


def test_aggregation_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:400-466 (test_aggregation): five SPAM trajectories
    with mocked bad atoms (0, 0, 1, 1, 2) on three non-interacting atoms after a pi pulse -- exact aggregated state,
    occupation and bitstrings; EnergyVariance is skipped with a warning; tags keep their observables' UUIDs."""
    from unittest.mock import patch

    import pulser
    from pulser.backend.default_observables import BitStrings, EnergyVariance, Occupation, StateResult

    reg = pulser.Register({"q0": [-1e5, 0], "q1": [1e5, 0], "q2": [0, 1e5]})
    seq = pulser.Sequence(reg, pulser.MockDevice)
    seq.declare_channel("ryd", "rydberg_global")
    seq.add(pulser.Pulse.ConstantDetuning(pulser.BlackmanWaveform(100, np.pi), 0.0, 0.0), "ryd")
    occup = Occupation(evaluation_times=[1.0])
    state = StateResult(evaluation_times=[1.0])
    bitstrings = BitStrings(evaluation_times=[1.0])
    variance = EnergyVariance(evaluation_times=[1.0])
    cfg = backend.B200Config(observables=(occup, state, bitstrings, variance), n_trajectories=5,
                             noise_model=pulser.NoiseModel(state_prep_error=1 / 3))
    with pytest.warns(UserWarning, match="Skipping aggregation of `energy_variance`."):
        with patch("pulser._hamiltonian_data.hamiltonian_data.np.random.uniform") as bad_atoms_mock:
            bad_atoms_mock.side_effect = [
                np.array([0.1, 0.5, 0.6]), np.array([0.1, 0.5, 0.6]), np.array([0.5, 0.1, 0.6]),
                np.array([0.5, 0.1, 0.6]), np.array([0.5, 0.6, 0.1]), np.array([0.1, 0.2, 0.3]),
            ] + [np.array([0.9, 0.9, 0.9])] * 6  # spare draws (noiseless helper data built more than once)
            results = backend.B200Backend(seq, config=cfg).run()
    expected_state = np.zeros((8, 8))
    expected_state[1, 1], expected_state[2, 2], expected_state[4, 4] = 0.2, 0.4, 0.4
    assert np.allclose(results.final_state.to_array(), expected_state, atol=1e-4)
    assert np.allclose(results.occupation[-1], np.array([0.6, 0.6, 0.8]), atol=1e-4)
    assert results.final_bitstrings == {"011": 2000, "101": 2000, "110": 1000}
    assert "energy_variance" not in results.get_result_tags()
    for obs_ in (occup, state, bitstrings):
        assert results.get_result_times(obs_) == [1.0]


def test_config_type_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:396-397."""
    assert backend.B200Backend.config_type is backend.B200Config


@pytest.mark.parametrize("modulation", [True, False])
def test_run_from_sequence_samples_port(backend, modulation):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:615-650 (test_run_from_sequence_samples): the
    Sequence entry point and the sampled-sequence entry point give the same state, to the last bit."""
    import pulser
    from pulser.backend.default_observables import StateResult
    from pulser.sampler import sample

    seq = pulser.Sequence(pulser.Register.square(1, prefix="q"), pulser.AnalogDevice)
    seq.declare_channel("rydberg_global", "rydberg_global")
    seq.add(pulser.Pulse.ConstantPulse(300, 1, 0, 0), "rydberg_global")
    config = None
    if modulation:
        initial_state = backend.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes={"g": 1.0})
        config = backend.B200Config(with_modulation=modulation, observables=[StateResult()],
                                    initial_state=initial_state)
    be = backend.B200Backend(seq, config=config)
    results1 = be.run()
    results2 = be.run_from_sequence_samples(
        sample(seq, modulation=modulation, extended_duration=seq.get_duration(include_fall_time=modulation)),
        seq.register, seq.device, config=config)
    s1 = results1.final_state.to_array()
    s2 = results2.final_state.to_array()
    assert np.allclose(s1, s2, atol=0, rtol=1e-16)  # really the same


def test_dmm_temperature_without_spot_waist_port(backend):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:583-612: register noise with a DMM needs
    `detuning_map_spot_waist` (check inherited from ``EmulatorBackend.__init__``)."""
    import dataclasses

    import pulser
    from pulser.backend.default_observables import StateResult
    from pulser.channels.dmm import DMM
    from pulser.devices import AnalogDevice

    reg = pulser.Register.from_coordinates([(0.0, 0.0), (6.0, 0.0)], center=False, prefix="q")
    det_map = reg.define_detuning_map({"q0": 1.0, "q1": 0.5})
    mock_device = dataclasses.replace(AnalogDevice.to_virtual(), dmm_objects=(DMM(),), reusable_channels=True)
    seq = pulser.Sequence(reg, mock_device)
    seq.declare_channel("ch0", "rydberg_global")
    seq.add(pulser.Pulse.ConstantPulse(100, 1, -1, 0), "ch0")
    seq.config_detuning_map(det_map, "dmm_0")
    seq.add_dmm_detuning(pulser.ConstantWaveform(100, -10), "dmm_0")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        config = backend.B200Config(noise_model=pulser.NoiseModel(trap_waist=1, trap_depth=1, temperature=0.5),
                                    observables=[StateResult(evaluation_times=[1.0])])
    with pytest.raises(ValueError, match="Combining register noise with a DMM requires"):
        backend.B200Backend(seq, config=config)


@pytest.mark.parametrize("amp_sigma", [0.0, 0.5])
def test_output_state_normalization_port(backend, amp_sigma):
    """reference tests/pulser_simulation/test_qutip_backend_v2.py:494-555 (test_output_state_normalization): states
    handed to the observables are normalised -- no fidelity above one."""
    import pulser
    from pulser.backend.default_observables import Fidelity

    factor = 1.2357175818662465 if not amp_sigma else 1.0
    r_interatomic = 5
    register = pulser.Register.hexagon(1, r_interatomic, prefix="q")
    seq = pulser.Sequence(register, pulser.MockDevice)
    seq.declare_channel("rydberg_global", "rydberg_global")
    u = pulser.AnalogDevice.interaction_coeff / r_interatomic**6
    total_duration = 400  # 4000 in the reference (oracle-backed run here)
    interp_pts = np.linspace(0, 1, 4)
    seq.add(pulser.Pulse(
        pulser.InterpolatedWaveform(total_duration, u * np.array([1e-9, 0.22, 0.2181, 1e-9]) * factor, times=interp_pts),
        pulser.InterpolatedWaveform(total_duration, u * np.array([-1, 0.0556, 0.332, 1]), times=interp_pts), 0),
        "rydberg_global")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noise_model = pulser.NoiseModel(amp_sigma=amp_sigma)
        np.random.seed(1234)
        config = backend.B200Backend.default_config.with_changes(noise_model=noise_model)
        results = backend.B200Backend(seq, config=config).run()
        final_state = results.final_state
        assert np.linalg.norm(final_state.to_array()) < 1 + 1e-8
        np.random.seed(1234)
        config = backend.B200Backend.default_config.with_changes(noise_model=noise_model,
                                                                 observables=[Fidelity(final_state)])
        results = backend.B200Backend(seq, config=config).run()
    assert results.fidelity[-1] < 1 + 1e-8


def test_density_matrix_aggregator_port(backend):
    """reference tests/pulser_simulation/test_aggregators.py:6-47."""
    mk = lambda s: backend.B200State.from_state_amplitudes(eigenstates=("r", "g"), amplitudes={s: 1.0})  # noqa: E731
    state1, state2, state3 = mk("rgg"), mk("grg"), mk("ggr")
    acc = backend.density_matrix_aggregator([state1, state2])  # vector and vector
    assert np.isclose(np.trace(acc.to_array()).real, 1.0)
    res1 = np.zeros((8, 8))
    res1[3, 3] = res1[5, 5] = 0.5
    assert np.allclose(acc.to_array(), res1)
    acc = backend.density_matrix_aggregator([acc, state3])  # vector and matrix
    res2 = np.zeros((8, 8))
    res2[3, 3] = res2[5, 5] = 0.25
    res2[6, 6] = 0.5
    assert np.allclose(acc.to_array(), res2)
    acc = backend.density_matrix_aggregator([acc, acc])  # matrix and matrix
    assert np.isclose(np.trace(acc.to_array()).real, 1.0)
    assert np.allclose(acc.to_array(), res2)


def test_config_ports(backend):
    """reference tests/pulser_simulation/test_qutip_config.py:17-146, on B200Config."""
    import json
    import re

    from pulser import NoiseModel
    from pulser.backend.default_observables import BitStrings, StateResult

    obs = [StateResult(evaluation_times=[1.0])]
    with pytest.raises(NotImplementedError, match="does not handle custom interaction matrices."):
        backend.B200Config(observables=obs, interaction_matrix=np.eye(4))
    with pytest.raises(ValueError, match="be greater than 0 and less than or equal to 1"):
        backend.B200Config(observables=obs, sampling_rate=1.2)
    config = backend.B200Config(observables=obs, sampling_rate=0.5)
    assert "sampling_rate" in config._expected_kwargs()
    with pytest.warns(UserWarning, match="The number of samples per run .* is ignored when using"):
        with pytest.warns(DeprecationWarning, match="Setting samples_per_run different to 1 is"):
            backend.B200Config(observables=obs, noise_model=NoiseModel(temperature=45, samples_per_run=5))
    with pytest.raises(TypeError, match=re.escape("If provided, `initial_state` must be an instance of `B200State`")):
        backend.B200Config(observables=obs, initial_state="all-ground")
    assert backend.B200Config.state_type is backend.B200State
    assert backend.B200Config.operator_type is backend.B200Operator
    config = backend.B200Config(observables=obs, progress_bar=True)
    assert config.progress_bar and "progress_bar" in config._expected_kwargs()
    default_times = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    obs_times_1 = np.array([0.2, 0.4, 0.8])
    obs_times_2 = np.array([0.15, 0.35, 0.65, 0.95])
    config = backend.B200Config(
        observables=[StateResult(evaluation_times=obs_times_1), StateResult(evaluation_times=obs_times_2, tag_suffix="second")],
        default_evaluation_times=default_times)
    expected = np.union1d(np.union1d(default_times, obs_times_1), obs_times_2)
    np.testing.assert_almost_equal(config._get_legacy_evaluation_times(1000), expected)
    for solver in backend.Solver:
        for as_str in (True, False):
            config = backend.B200Config(observables=[BitStrings(evaluation_times=[1.0])],
                                        solver=solver if not as_str else str(solver.value))
            ser = config.to_abstract_repr()
            assert json.loads(ser)["solver"] == str(solver.value)
            assert backend.B200Config.from_abstract_repr(ser).solver is solver
    with pytest.raises(ValueError, match="Invalid solver 'fakesolver'"):
        backend.B200Config(observables=[BitStrings(evaluation_times=[1.0])], solver="fakesolver")


def test_package_exports_mirror_pulser_simulation():
    """pulser_simulation/__init__.py:17-40: the emulator, both backends, config / state / operator types, the solver
    enum and the aggregator are importable from the package root."""
    import pulser_b200

    for name in ("B200Emulator", "B200Backend", "B200LegacyBackend", "B200Config", "B200State", "B200Operator",
                 "Solver", "density_matrix_aggregator"):
        assert getattr(pulser_b200, name) is not None and name in pulser_b200.__all__
    with pytest.raises(AttributeError):
        pulser_b200.QutipEmulator

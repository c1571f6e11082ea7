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

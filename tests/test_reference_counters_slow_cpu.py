"""The reference's heavier Counter goldens through the facade (oracle-backed plans): minutes of CPU, so they only run
with PB200_SLOW_TESTS=1 (``PB200_SLOW_TESTS=1 python -m pytest tests/test_reference_counters_slow_cpu.py``).  All of
them were run and reproduce the hard-coded Counters EXACTLY; the lighter ones are in the default
suite (tests/test_oracle_cpu.py, tests/test_emulator_cpu.py).

* test_noisy_xy (reference tests/pulser_simulation/test_simulation.py:1536-1700, MESOLVER variants): XY mode, 15 SPAM
  trajectories x 10 samples with measurement errors, collapse operators, with and without an SLM mask;
* test_noises_all (:1174-1302): three-level basis, 9-15 us sequences, effective noise / dephasing / relaxation;
* the noisy goldens of tests/pulser_simulation/test_simresults.py (results_noisy fixture and the three tests using it).
"""
import os
import warnings
from collections import Counter

import numpy as np
import pytest

from pulser_b200 import HAVE_PULSER

pytestmark = [
    pytest.mark.skipif(not HAVE_PULSER, reason="pulser-core not importable here"),
    pytest.mark.skipif(os.environ.get("PB200_SLOW_TESTS") != "1", reason="minutes of CPU: set PB200_SLOW_TESTS=1"),
]

RES_DEPH_MEQ = {"0000": 830, "0001": 21, "0010": 3, "0100": 80, "1000": 66}
RES_EFF_MEQ = {"0000": 851, "0001": 23, "0010": 8, "0100": 57, "1000": 61}
RES_DEPOL_MEQ = {"0000": 791, "0001": 39, "0010": 10, "0100": 81, "0110": 2, "1000": 67, "1010": 10}
RES_DEPH_ATOM1_MEQ = {"0000": 804, "0001": 105, "0010": 12, "0100": 54, "0101": 8, "1000": 17}
RES_DEPH_ATOM2_MEQ = {"0000": 575, "0001": 334, "0011": 12, "0100": 13, "1000": 56, "1001": 10}


@pytest.fixture
def emu(monkeypatch):
    from fake_device import FakeDevicePlan, FakeLindbladPlan
    from pulser_b200 import emulator, engine, lindblad

    monkeypatch.setattr(engine, "DevicePlan", FakeDevicePlan)
    monkeypatch.setattr(lindblad, "LindbladPlan", FakeLindbladPlan)
    return emulator


@pytest.mark.parametrize("masked_qubit,noise,result", [
    (None, "dephasing", RES_DEPH_MEQ), (None, "eff_noise", RES_EFF_MEQ), (None, "leakage", RES_EFF_MEQ),
    (None, "depolarizing", RES_DEPOL_MEQ), ("atom0", "dephasing", RES_DEPH_ATOM1_MEQ),
    ("atom1", "dephasing", RES_DEPH_ATOM2_MEQ),
])
def test_noisy_xy_port(emu, masked_qubit, noise, result):
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import MockDevice
    from pulser.noise_model import _LEGACY_DEFAULTS

    np.random.seed(15092021)
    seq = Sequence(Register.square(2, prefix="atom"), MockDevice)
    seq.declare_channel("ch0", "mw_global")
    if masked_qubit is not None:
        seq.config_slm_mask([masked_qubit])
    seq.add(Pulse.ConstantPulse(1000, 3.0, 1.0, 0.0), "ch0")
    with_leakage = noise == "leakage"
    if with_leakage or noise == "eff_noise":
        op = np.diag([1.0, -1.0, 0.0]) if with_leakage else np.diag([1.0, -1.0])
        params = dict(eff_noise_opers=[op.astype(complex)], eff_noise_rates=[1.0])
    else:
        params = {f"{noise}_rate": _LEGACY_DEFAULTS[f"{noise}_rate"]}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(
            seq, sampling_rate=0.1,
            noise_model=NoiseModel(samples_per_run=10, with_leakage=with_leakage, state_prep_error=0.4,
                                   p_false_pos=0.01, p_false_neg=0.05, **params),
            n_trajectories=15, solver=emu.Solver.MESOLVER)
        assert set(sim.noise_model.noise_types) == ({"SPAM", noise} if not with_leakage
                                                    else {"SPAM", "leakage", "eff_noise"})
        assert [bool(b) for b in sim._current_spec.bad_atoms] == [True, False, True, False]
        got = sim.run().sample_final_state()
    assert got == Counter(result)


@pytest.mark.parametrize("noise,result", [
    (("dephasing",), {"111": 961, "101": 15, "110": 14, "011": 9, "001": 1}),
    (("eff_noise",), {"111": 961, "101": 15, "110": 14, "011": 9, "001": 1}),
    (("relaxation",), {"000": 459, "010": 202, "001": 168, "100": 167, "101": 4}),
    (("dephasing", "relaxation"), {"000": 451, "010": 205, "001": 170, "100": 168, "101": 6}),
    (("eff_noise", "dephasing"), {"111": 932, "101": 28, "011": 24, "110": 15, "001": 1}),
])
def test_noises_all_port(emu, noise, result):
    from pulser import NoiseModel, Pulse
    from pulser.waveforms import BlackmanWaveform
    from test_emulator_cpu import _ccz_sequence

    seq = _ccz_sequence()
    params = {}
    if "relaxation" in noise:
        pi_y = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, -np.pi / 2)
        for q in ("control1", "target", "control2"):
            seq.target(q, "raman")
            seq.add(pi_y, "raman")
        seq.declare_channel("ryd_glob", "rydberg_global")
        seq.add(Pulse.ConstantDetuning(BlackmanWaveform(1000, 2 * np.pi), 0.0, 0), "ryd_glob")
        seq.measure()
        params["relaxation_rate"] = 1.0
    if "dephasing" in noise:
        params["hyperfine_dephasing_rate"] = 0.1
        params["dephasing_rate"] = 0.1
    if "eff_noise" in noise:
        params["eff_noise_opers"] = [np.diag([1.0, 0, 0]).astype(complex), np.diag([0, 0, 1.0]).astype(complex)]
        params["eff_noise_rates"] = [0.2, 0.2]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = emu.B200Emulator.from_sequence(seq, sampling_rate=0.01, noise_model=NoiseModel(**params))
        assert set(sim.noise_model.noise_types) == set(noise)
        np.random.seed(123)
        got = sim.run().sample_final_state()
    assert got == Counter(result)


def test_simresults_noisy_goldens_port(emu):
    """reference tests/pulser_simulation/test_simresults.py:52-83 (results_noisy fixture), :448-483
    (test_sample_final_state_noisy), :383-390 (test_expect_noisy), :244-276 (test_get_final_state_noisy): fifteen noisy
    sesolve trajectories (doppler, amplitude, SPAM) under np.random.seed(123), drawn trajectory by trajectory like the
    reference (b200_batch=1).  The reference's hard-coded Counter, expectation value and pseudo-density all come out
    exactly -- including the second emulator of the same test, which continues on the same random stream."""
    from pulser import NoiseModel, Pulse, Register, Sequence
    from pulser.devices import DigitalAnalogDevice
    from pulser.waveforms import BlackmanWaveform

    reg = Register({"A": np.array([0.0, 0.0]), "B": np.array([0.0, 10.0])})
    pi_pulse = Pulse.ConstantDetuning(BlackmanWaveform(1000, np.pi), 0.0, 0)
    seq_no_meas = Sequence(reg, DigitalAnalogDevice)
    seq_no_meas.declare_channel("ryd", "rydberg_global")
    seq_no_meas.add(pi_pulse, "ryd")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(123)
        sim = emu.B200Emulator.from_sequence(
            seq_no_meas,
            noise_model=NoiseModel(samples_per_run=5, temperature=50.0, state_prep_error=0.005, p_false_pos=0.01,
                                   p_false_neg=0.05, amp_sigma=1e-3, laser_waist=175.0),
            n_trajectories=15)
        results_noisy = sim.run(b200_batch=1)
        np.random.seed(123)
        assert results_noisy.sample_final_state(N_samples=1234) == Counter({"11": 676, "10": 295, "01": 137, "00": 126})
        res_3level = emu.B200Emulator.from_sequence(
            seq_no_meas,
            noise_model=NoiseModel(samples_per_run=5, temperature=50.0, state_prep_error=0.005, p_false_pos=0.01,
                                   p_false_neg=0.05),
            n_trajectories=10)
        final_state = res_3level.run(b200_batch=1).states[-1]
    assert np.isclose(np.asarray(final_state), np.array([0.38, 0.32, 0.2, 0.1])).all()
    # test_expect_noisy
    bad_op = np.kron(np.eye(2), np.array([[0.0, 1.0], [0.0, 0.0]]))
    with pytest.raises(ValueError, match="non-diagonal"):
        results_noisy.expect([bad_op])
    op = np.kron(np.eye(2), np.diag([1.0, 0.0]))
    assert np.isclose(results_noisy.expect([op])[0][-1], 0.68)
    # test_get_final_state_noisy
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(123)
        seq_ = Sequence(reg, DigitalAnalogDevice)
        seq_.declare_channel("ram", "raman_local", initial_target="A")
        seq_.add(pi_pulse, "ram")
        sim_noisy = emu.B200Emulator.from_sequence(
            seq_,
            noise_model=NoiseModel(samples_per_run=5, temperature=50.0, trap_depth=0.01, trap_waist=0.02,
                                   state_prep_error=0.005, p_false_pos=0.01, p_false_neg=0.05),
            n_trajectories=15)
        res3 = sim_noisy.run(b200_batch=1)
    final = np.asarray(res3.get_final_state())
    assert final[0] == 0.04 and final[2] == 0.96
    assert res3.results[-1] == Counter({"10": 0.96, "00": 0.04})

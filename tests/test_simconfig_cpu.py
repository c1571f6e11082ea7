"""``pulser_b200.SimConfig`` against the reference's own ``tests/pulser_simulation/test_simconfig.py`` (ported
statement by statement, source lines cited; numpy matrices stand where the reference builds ``qutip.Qobj``s because
QuTiP is not installable here) plus the C4 phrasing of BASELINE.json (``SimConfig(noise=("doppler", "amplitude"))``)."""
import numpy as np
import pytest

from pulser_b200 import HAVE_PULSER

pytestmark = [
    pytest.mark.skipif(not HAVE_PULSER, reason="pulser-core not importable"),
    pytest.mark.filterwarnings("ignore:'SimConfig' has been deprecated:DeprecationWarning",
                               "ignore:.*'NoiseModel.runs' is deprecated:DeprecationWarning"),
]


class _FakeQobj:
    """What qutip.Qobj offers to SimConfig: ``full()``."""

    def __init__(self, mat):
        self._m = np.asarray(mat, dtype=complex)

    def full(self):
        return self._m


@pytest.fixture
def matrices():
    return {"I": np.eye(2), "X": np.array([[0.0, 1.0], [1.0, 0.0]]), "Zh": 0.5 * np.diag([1.0, -1.0]),
            "ket": np.array([[1.0], [2.0]]), "I3": np.eye(3), "I4": np.eye(4)}


@pytest.mark.filterwarnings("ignore:Setting samples_per_run different to 1 is")
def test_init_port():  # test_simconfig.py:41-106
    from pulser_b200 import SimConfig
    from pulser._hamiltonian_data.hamiltonian_data import doppler_sigma

    with pytest.deprecated_call(match="'SimConfig' has been deprecated"):
        config = SimConfig(noise=("SPAM", "doppler", "dephasing", "amplitude"), temperature=1000.0, runs=100)
    assert config.temperature == 1000.0 * 1e-6  # in K
    text = config.__str__(True)
    assert "SPAM, doppler, dephasing, amplitude" in text
    assert "1000.0µK" in text and "100" in text and "Solver Options" in text
    assert config.to_noise_model().temperature == 1000.0
    config = SimConfig(noise=("depolarizing", "relaxation", "doppler"))
    assert config.temperature == pytest.approx(50.0e-6)
    assert config.to_noise_model().temperature == 50.0
    text = config.__str__(True)
    assert "depolarizing" in text and "relaxation" in text
    assert f"Depolarizing rate: {config.depolarizing_rate}" in text
    assert f"Relaxation rate: {config.relaxation_rate}" in text
    config = SimConfig(noise="eff_noise", eff_noise_opers=[_FakeQobj(np.eye(2)), np.array([[0, 1], [1, 0]])],
                       eff_noise_rates=[0.3, 0.7])
    text = config.__str__(True)
    assert config.doppler_sigma == doppler_sigma(50.0 * 1e-6)
    assert "Effective noise rates" in text and "Effective noise operators" in text
    with pytest.raises(TypeError, match="'temperature' must be a float"):
        SimConfig(temperature="0.0")
    with pytest.raises(ValueError, match="SPAM parameter"):
        SimConfig(eta=-1.0)
    with pytest.raises(ValueError, match="'amp_sigma' must be greater than or equal to zero"):
        SimConfig(amp_sigma=-0.001)
    with pytest.raises(ValueError, match="'bad_noise' is not a valid noise type."):
        SimConfig(noise=("bad_noise",))


@pytest.mark.filterwarnings("ignore:Setting samples_per_run different to 1 is")
def test_eff_noise_opers_port(matrices):  # test_simconfig.py:109-163
    from pulser_b200 import SimConfig

    with pytest.raises(ValueError, match="The operators list length"):
        SimConfig(noise=("eff_noise"), eff_noise_rates=[1.0])
    with pytest.raises(TypeError, match="eff_noise_rates is a list of floats"):
        SimConfig(noise=("eff_noise"), eff_noise_rates=["0.1"], eff_noise_opers=[matrices["I"]])
    with pytest.raises(ValueError, match="The effective noise parameters have not been filled."):
        SimConfig(noise=("eff_noise"))
    with pytest.raises(TypeError, match="is not a Qobj."):
        SimConfig(noise=("eff_noise"), eff_noise_opers=[2.0], eff_noise_rates=[1.0])
    with pytest.raises(TypeError, match="to be of Qutip type 'oper'."):
        SimConfig(noise=("eff_noise"), eff_noise_opers=[matrices["ket"]], eff_noise_rates=[1.0])
    with pytest.raises(ValueError, match="With leakage, operator's shape"):
        SimConfig(noise=("eff_noise", "leakage"), eff_noise_opers=[matrices["I"]], eff_noise_rates=[1.0])
    with pytest.raises(ValueError, match="With leakage, operator's shape"):
        SimConfig(noise=("eff_noise", "leakage"), eff_noise_opers=[np.eye(5)], eff_noise_rates=[1.0])
    with pytest.raises(ValueError, match="Without leakage, operator's shape"):
        SimConfig(noise=("eff_noise",), eff_noise_opers=[matrices["I4"]], eff_noise_rates=[1.0])
    SimConfig(noise=("eff_noise"), eff_noise_opers=[matrices["X"], matrices["I"]], eff_noise_rates=[0.5, 0.5])


def test_noise_model_conversion_port():  # test_simconfig.py:166-186
    from pulser.noise_model import NoiseModel
    from pulser_b200 import SimConfig

    noise_model = NoiseModel(p_false_neg=0.4, p_false_pos=0.1, amp_sigma=1e-3, runs=10, samples_per_run=1)
    expected = SimConfig(noise=("SPAM", "amplitude"), epsilon=0.1, epsilon_prime=0.4, eta=0.0, amp_sigma=1e-3,
                         laser_waist=float("inf"), runs=10, samples_per_run=1)
    assert SimConfig.from_noise_model(noise_model) == expected
    assert expected.to_noise_model() == noise_model


def test_c4_phrasing_reaches_the_emulator():
    """BASELINE configs[3]: SimConfig(doppler + amplitude noise) handed to from_sequence (the deprecated entry point
    of simulation.py:955-1051) produces the noise model the trajectories are drawn from."""
    import pulser
    from pulser_b200 import B200Emulator, SimConfig

    reg = pulser.Register.square(2, 6.0, prefix="q")
    seq = pulser.Sequence(reg, pulser.MockDevice)
    seq.declare_channel("ryd", "rydberg_global")
    seq.add(pulser.Pulse.ConstantPulse(200, 2 * np.pi, 0.0, 0.0), "ryd")
    cfg = SimConfig(noise=("doppler", "amplitude"), runs=7, temperature=50.0, amp_sigma=0.05)
    with pytest.deprecated_call(match="Supplying a 'SimConfig'"):
        emu = B200Emulator.from_sequence(seq, config=cfg)
    nm = emu.noise_model
    assert set(nm.noise_types) == {"doppler", "amplitude"} and nm.temperature == 50.0 and nm.amp_sigma == 0.05
    assert emu.config == cfg or emu.config.to_noise_model() == nm

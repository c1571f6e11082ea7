"""CPU tests of the result layer (mirrors of QutipResult / simresults)."""
from collections import Counter

import numpy as np
import pytest

from helpers import random_state
from pulser_b200.results import B200Result, CoherentResults, StateVector


def _brute_weights(probs, dim, size, one_idx):
    """The reference's python loop, qutip_result.py:124-151."""
    ex_one = [i for i in range(dim) if i != one_idx]
    probs = probs.reshape([dim] * size)
    w = np.zeros(2**size)
    for dec in range(2**size):
        ind = [ex_one if v == "0" else [one_idx] for v in np.binary_repr(dec, width=size)]
        w[dec] = np.sum(probs[np.ix_(*ind)])
    return w / w.sum()


def test_weights_ground_rydberg_reversed():
    psi = random_state(8, 1)
    r = B200Result(("a", "b", "c"), "ground-rydberg", StateVector(psi), True)
    np.testing.assert_allclose(r._weights(), (np.abs(psi) ** 2)[::-1] / np.sum(np.abs(psi) ** 2))
    assert r._basis_name == "ground-rydberg" and r._eigenbasis == ["r", "g"]


@pytest.mark.parametrize("meas,one", [("ground-rydberg", 0), ("digital", 2)])
def test_weights_all_basis_marginalisation(meas, one):
    psi = random_state(3**4, 2)
    r = B200Result(tuple("abcd"), meas, StateVector(psi), False)
    assert r._basis_name == "all" and r._eigenbasis == ["r", "g", "h"]
    np.testing.assert_allclose(r._weights(), _brute_weights(np.abs(psi) ** 2, 3, 4, one), atol=1e-15)


def test_sampling_uses_reference_rng_recipe():
    psi = random_state(16, 3)
    r = B200Result(tuple("abcd"), "ground-rydberg", StateVector(psi), True)
    np.random.seed(42)
    got = r.get_samples(500)
    np.random.seed(42)
    idx = np.searchsorted(np.cumsum(r._weights()), np.random.rand(500))
    assert got == Counter(np.binary_repr(i, 4) for i in idx)
    assert sum(got.values()) == 500


def test_get_state_global_phase_and_reduce():
    psi = np.zeros(9, dtype=complex)
    psi[4] = np.exp(0.7j) * 0.8  # |g g> in [r,g,h]
    psi[5] = 0.6j                # |g h>
    r = B200Result(("a", "b"), "digital", StateVector(psi), False)
    st = r.get_state()
    assert abs(st.full()[4, 0].imag) < 1e-15 and st.full()[4, 0].real > 0
    red = r.get_state(reduce_to_basis="digital")
    assert red.shape == (4, 1) and abs(red.norm() - 1) < 1e-12
    psi[0] = 0.3
    with pytest.raises(TypeError, match="Can't reduce to chosen basis"):
        B200Result(("a", "b"), "digital", StateVector(psi), False).get_state(reduce_to_basis="digital")


def _coherent(states, times, basis="ground-rydberg", errors=None):
    res = [B200Result(("a", "b"), basis, StateVector(s), True, t) for s, t in zip(states, times)]
    return CoherentResults(res, 2, basis, np.array(times), basis, errors)


def test_expect_and_pseudo_density():
    s0 = np.array([0, 0, 0, 1], dtype=complex)  # gg
    s1 = np.array([1, 0, 0, 0], dtype=complex)  # rr
    cr = _coherent([s0, s1], [0.0, 1.0])
    n0 = np.diag([1.0, 1.0, 0.0, 0.0])  # |r><r| on qubit 0
    np.testing.assert_allclose(cr.expect([n0])[0], [0.0, 1.0])
    with pytest.raises(ValueError, match="Incompatible shape"):
        cr.expect([np.eye(3)])
    eps, epsp = 0.1, 0.2
    ce = _coherent([s0, s1], [0.0, 1.0], errors={"epsilon": eps, "epsilon_prime": epsp})
    # reference tests/pulser_simulation/test_simresults.py:289-380: SPAM limits
    np.testing.assert_allclose(ce.expect([n0])[0], [eps, 1 - epsp])
    with pytest.raises(ValueError, match="non-diagonal"):
        ce.expect([np.ones((4, 4))])


def test_sample_state_with_measurement_errors_statistics():
    s = np.array([0, 0, 0, 1], dtype=complex)
    ce = _coherent([s, s], [0.0, 1.0], errors={"epsilon": 0.3, "epsilon_prime": 0.0})
    np.random.seed(0)
    c = ce.sample_final_state(20000)
    assert abs(c["00"] / 20000 - 0.49) < 0.02 and abs(c["11"] / 20000 - 0.09) < 0.01
    with pytest.raises(IndexError, match="absent from simulation times"):
        ce.sample_state(0.5)


# ---- ports of the reference's tests/pulser_simulation/test_simresults.py (no device involved) ------------------
@pytest.mark.parametrize("basis,exp_basis", [
    ("ground-rydberg_with_error", "ground-rydberg"), ("digital_with_error", "digital"), ("all_with_error", "digital"),
    ("all", "digital"), ("XY_with_error", "XY"),
])
def test_initialization_messages_port(basis, exp_basis):
    """reference test_simresults.py:95-135 (test_initialization), constructor messages."""
    from pulser_b200.results import CoherentResults

    with pytest.raises(ValueError, match="`basis_name` must be"):
        CoherentResults([], 2, "bad_basis", None, [0])
    if "all" in basis:
        with pytest.raises(ValueError, match="`meas_basis` must be 'ground-rydberg' or 'digital'."):
            CoherentResults([], 1, basis, None, "XY")
    else:
        with pytest.raises(ValueError, match=f"`meas_basis` associated to basis_name '{basis}' must be"):
            CoherentResults([], 1, basis, [0], "wrong_measurement_basis")
    with pytest.raises(ValueError, match="only values of 'epsilon' and 'epsilon_prime'"):
        CoherentResults([], 1, basis, [0], exp_basis, {"eta": 0.1, "epsilon": 0.0, "epsilon_prime": 0.4})


@pytest.mark.parametrize("basis,exp_basis", [
    ("ground-rydberg_with_error", "ground-rydberg"), ("digital_with_error", "digital"), ("all_with_error", "digital"),
    ("all", "digital"), ("XY_with_error", "XY"),
])
def test_init_noisy_port(basis, exp_basis):
    """reference test_simresults.py:138-152 (test_init_noisy)."""
    from pulser_b200.results import NoisyResults

    with pytest.raises(ValueError, match="`basis_name` must be"):
        NoisyResults([], 2, "bad_basis", [0], 123)
    assert NoisyResults([], 2, basis, [0], 100)._basis_name == exp_basis

"""GPU tests against the committed golden fixtures (tests/golden/*.npz,
generated from the real pulser-core + oracle by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from pulser_b200.spec import HamiltonianSpec

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
STATE_TOL = 1e-8


def load(name):
    with np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False) as data:
        return HamiltonianSpec.from_npz(data), {k: data[k] for k in data.files}


@pytest.fixture(scope="module")
def engine(lib):
    from pulser_b200 import engine

    assert engine.device_count() > 0
    return engine


def dense_h(plan, D, t_us):
    eye = np.eye(D, dtype=complex)
    return np.stack([plan.apply_h(t_us, eye[j]) for j in range(D)], axis=1)


def test_get_hamiltonian_reference_goldens(engine):
    """reference test_simulation.py:476-588 through pb200_apply_h."""
    spec, extra = load("ref_get_hamiltonian_rate001")
    with engine.DevicePlan(spec) as plan:
        h = dense_h(plan, 4, float(extra["t_ns"]) / 1000)
    assert np.isclose(h[0, 0], float(extra["h00"]))
    for name in ("ref_get_hamiltonian_doppler", "ref_get_hamiltonian_register"):
        spec, extra = load(name)
        with engine.DevicePlan(spec) as plan:
            h = dense_h(plan, 4, float(extra["t_ns"]) / 1000)
        np.testing.assert_allclose(h, extra["h"], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("name", [
    "ref_initial_state_sim", "ref_qutip_backend_pi_pulse", "ref_delays_occupation",
    "orc_c1_square", "orc_all_basis_3atoms", "orc_noisy_traj0", "orc_noisy_traj1", "orc_noisy_traj2",
    "orc_xy_evolution", "ref_mask_two_pulses_xy", "orc_xy_slm_evolution",
])
def test_final_state_goldens(engine, name):
    spec, extra = load(name)
    with engine.DevicePlan(spec) as plan:
        plan.set_state(extra["psi0"])
        plan.propagate(0.0, spec.sampling_times[-1])
        got = plan.get_state()[0]
    assert np.max(np.abs(got - extra["orc_final"])) < STATE_TOL
    if "ref_final" in extra:  # reference's own loose pin (rtol 1e-2)
        g = got * np.exp(-1j * np.angle(got[np.argmax(np.abs(got))]))
        assert np.max(np.abs(g - extra["ref_final"])) < 1e-2
    if "ref_final_abs" in extra:
        np.testing.assert_allclose(np.abs(got), extra["ref_final_abs"], atol=1e-5)
    if "ref_r_occupation" in extra:
        assert abs(abs(got[0]) ** 2 - float(extra["ref_r_occupation"])) < 1e-4


def test_noisy_trajectories_as_one_batch(engine):
    specs, finals, psi0 = [], [], None
    for i in range(3):
        s, e = load(f"orc_noisy_traj{i}")
        specs.append(s); finals.append(e["orc_final"]); psi0 = e["psi0"]
    with engine.DevicePlan(specs) as plan:
        plan.set_state(psi0)
        plan.propagate(0.0, specs[0].sampling_times[-1])
        got = plan.get_state()
    for g, f in zip(got, finals):
        assert np.max(np.abs(g - f)) < STATE_TOL
    probs_sum = np.sum(np.abs(got) ** 2, axis=1)
    np.testing.assert_allclose(probs_sum, 1.0, atol=1e-9)


def test_all_basis_apply_h(engine):
    from oracle.matfree import MatFreeHamiltonian

    spec, _ = load("orc_all_basis_3atoms")
    mf = MatFreeHamiltonian(spec)
    rng = np.random.default_rng(0)
    v = rng.normal(size=27) + 1j * rng.normal(size=27)
    with engine.DevicePlan(spec) as plan:
        for t in (0.2, 1.0, 1.9):
            assert np.max(np.abs(plan.apply_h(t, v) - mf.apply(t, v))) < 1e-12


def test_nonuniform_sampling_grid(engine):
    """sampling_rate = 0.3: integer-truncated, non-uniform sampling times (hamiltonian.py:87-95)."""
    spec, extra = load("orc_sampling_rate_03")
    assert len(np.unique(np.round(np.diff(spec.sampling_times), 9))) > 1
    with engine.DevicePlan(spec) as plan:
        plan.set_state(extra["psi0"])
        plan.propagate(0.0, spec.sampling_times[-1])
        got = plan.get_state()[0]
    assert np.max(np.abs(got - extra["orc_final"])) < STATE_TOL


def test_leakage_effective_noise_lindblad(engine):
    """3-level (r, g, x) register with effective-noise collapse operators: master equation, tol 1e-4."""
    from pulser_b200.lindblad import LindbladPlan

    spec, extra = load("orc_leakage_lindblad")
    with LindbladPlan(spec) as lp:
        lp.set_state(extra["psi0"])
        lp.propagate(0.0, spec.sampling_times[-1])
        rho = lp.get_rho()[0]
    assert np.max(np.abs(rho - extra["orc_rho"])) < 1e-4
    assert abs(np.trace(rho).real - 1.0) < 1e-6
    x_pop = sum(rho[i, i].real for i in range(9) if 2 in divmod(i, 3))
    assert x_pop > 1e-3  # the leakage state got populated


@pytest.mark.parametrize("order", [0, 1, 3])
def test_interpolation_orders(engine, order):
    """QobjEvo coefficient interpolation is switchable (step / linear / cubic); each matches the oracle run
    with the same rule."""
    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    spec, extra = load("orc_noisy_traj1")
    ref = evolve.sesolve(OracleHamiltonian.from_spec(spec), extra["psi0"], [0.0, spec.sampling_times[-1]],
                         order=order, rtol=1e-12, atol=1e-14, max_step=2.5e-4)[-1]
    with engine.DevicePlan(spec, interp_order=order) as plan:
        plan.set_state(extra["psi0"])
        plan.propagate(0.0, spec.sampling_times[-1])
        got = plan.get_state()[0]
    # piecewise-constant / piecewise-linear H(t) limits the accuracy of the adaptive ODE oracle, not ours
    assert np.max(np.abs(got - ref)) < (STATE_TOL if order == 3 else 5e-7)


def test_get_xy_hamiltonian_reference_golden(engine):
    """reference tests/pulser_simulation/test_simulation.py:1430-1491 on the device: H(143 ns) column by column."""
    spec, extra = load("ref_get_xy_hamiltonian")
    c3, c6 = float(extra["c3"]), float(extra["c6"])
    with engine.DevicePlan(spec) as plan:
        h = dense_h(plan, spec.hilbert_dim, float(extra["t_ns"]) / 1000)
    assert h[1, 2] == c3 / 10**3
    assert abs(h[1, 4] - (-2 * c3 / 10**3)) < 1e-10
    assert h[0, 1] == 0.5 * 3.0
    n_d = np.array([0, 1, 1, 2, 1, 2, 2, 3])
    vdw = np.array([2 + 1 / 8, 1 / 8, 1, 0, 1, 0, 0, 0]) * c6 / 1e6
    np.testing.assert_array_almost_equal(np.diag(h).real, -1.0 * n_d + vdw)
    assert np.allclose(h, h.conj().T)


def test_xy_slm_mask_reference_property(engine):
    """reference test_simulation.py:1792-1838 (test_mask_two_pulses_xy) through pb200_apply_h: while the SLM mask
    is on the Hamiltonian is (two-qubit H) x 1, afterwards the unmasked three-qubit one (matrices produced from the
    real pulser objects by tests/golden/make_golden.py --slm), and in between the spline-weighted mixture the oracle
    builds from the two QobjEvo terms of hamiltonian.py:399-424."""
    from oracle.ref_hamiltonian import OracleHamiltonian

    spec, extra = load("ref_mask_two_pulses_xy")
    orc = OracleHamiltonian.from_spec(spec)
    with engine.DevicePlan(spec) as plan:
        for t, ref in zip((0.01, 0.05, 0.09), extra["h_two_kron"]):
            np.testing.assert_allclose(dense_h(plan, 8, t), ref, atol=1e-11)
        for t, ref in zip((0.15, 0.2, 0.29), extra["h_three"]):
            np.testing.assert_allclose(dense_h(plan, 8, t), ref, atol=1e-11)
        for t in (0.0985, 0.0995, 0.1003, 0.1012, 0.1049):  # around the switch: interpolated 0/1 coefficient
            np.testing.assert_allclose(dense_h(plan, 8, t), orc.matrix_at(t).toarray(), atol=1e-11)


@pytest.mark.parametrize("integrator", [1, 2])
def test_xy_slm_mask_evolution_integrators(engine, integrator):
    spec, extra = load("orc_xy_slm_evolution")
    with engine.DevicePlan(spec) as plan:
        plan.set_state(extra["psi0"])
        st = plan.propagate(0.0, spec.sampling_times[-1], integrator=integrator)
        got = plan.get_state()[0]
    assert st["integrator"] == integrator
    assert np.max(np.abs(got - extra["orc_final"])) < STATE_TOL


def test_effective_size_disjoint_xy_reference_golden(engine):
    """reference test_simulation.py:1960-1998 through pb200_apply_h (bad atoms + SLM mask in XY mode)."""
    spec, extra = load("ref_effective_size_disjoint_xy")
    with engine.DevicePlan(spec) as plan:
        np.testing.assert_allclose(dense_h(plan, 16, 0.0), extra["h0"], atol=1e-13)
        plan.set_state("all-ground")
        plan.propagate(0.0, spec.sampling_times[-1])
        assert abs(plan.norm2()[0] - 1.0) < 1e-10

"""GPU parity tests of the time-dependent Taylor propagator (``integrator=3``; what ``integrator=0`` picks for one
global drive of constant phase on a d = 2 register).

It replaces ``qutip.sesolve`` (reference ``pulser_simulation/simulation.py:729-735``) on the QobjEvo the reference
builds at ``hamiltonian.py:246-439``.  Bound: max |psi_gpu - psi_oracle| <= 1e-8 (north star, Schroedinger).
"""
import numpy as np
import pytest

from helpers import random_local_spec, random_state
from pulser_b200 import workloads as W

pytestmark = pytest.mark.gpu

STATE_TOL = 1e-8


@pytest.fixture(scope="module")
def engine(lib):
    from pulser_b200 import engine

    assert engine.device_count() > 0, "GPU tests need a CUDA device"
    return engine


def _oracle(spec, psi0, times):
    from oracle import evolve
    from oracle.ref_hamiltonian import OracleHamiltonian

    return evolve.sesolve(OracleHamiltonian.from_spec(spec), psi0, times, rtol=1e-13, atol=1e-15)


def _blackman_spec(n, phase=0.0, T=600):
    """smooth (curved) waveforms: the polynomial degrees above 1 and the stored-gather history are exercised"""
    amp = W.blackman(T, 2.2 * np.pi)
    det = -8.0 + 20.0 * np.sin(np.linspace(0.0, 1.3, T)) ** 2
    coords = W.disc_register(n, 16.0, 5.0, 3)
    return W.ising_global_spec(coords, W.C6_LEVEL_60, amp, det, phase=phase)


@pytest.mark.parametrize("n", [1, 2, 6, 10, 11, 12])
def test_blockade_sweep_vs_oracle(engine, n):
    """C2-shaped sequence: small registers run the one-thread-per-amplitude stage, N >= 11 the tiled one."""
    from oracle import evolve

    spec = W.config_c2(n=n, seed=20)
    psi0 = evolve.all_ground_state(spec)
    tf = spec.sampling_times[-1]
    ref = _oracle(spec, psi0, [0.0, tf])[-1]
    with engine.DevicePlan(spec) as plan:
        plan.set_state("all-ground")
        st = plan.propagate(0.0, tf, integrator=3)
        got = plan.get_state()[0]
    assert st["integrator"] == 3 and st["n_checks"] == 0 and st["n_applies"] > 0
    assert st["n_applies"] < 2.5 * spec.total_duration_ns
    assert np.max(np.abs(got - ref)) < STATE_TOL
    assert abs(np.linalg.norm(got) - 1.0) < 1e-9
    assert st["err_estimate"] < 1e-8   # a-priori bound: fit residuals + Taylor remainders


@pytest.mark.parametrize("n,phase", [(7, 0.0), (7, 0.83), (12, -2.1)])
def test_smooth_waveforms_and_constant_phase(engine, n, phase):
    """Blackman amplitude + curved detuning, drive phase 0 and != 0 (complex unit, the Q sums of the kernel)."""
    spec = _blackman_spec(n, phase)
    psi0 = random_state(spec.hilbert_dim, 5)
    tf = spec.sampling_times[-1]
    ref = _oracle(spec, psi0, [0.0, tf])[-1]
    with engine.DevicePlan(spec) as plan:
        plan.set_state(psi0)
        st = plan.propagate(0.0, tf, integrator=3)
        got = plan.get_state()[0]
    assert st["integrator"] == 3
    assert np.max(np.abs(got - ref)) < STATE_TOL


def test_auto_rule_and_evaluation_times(engine):
    """integrator 0 takes the Taylor propagator here; a run cut at arbitrary (off-grid) evaluation times gives the
    oracle's states at each of them, and so does a run cut at EVERY sampling time ("Full" evaluation times: one call,
    i.e. one exact cubic step, per interval)."""
    from oracle import evolve

    spec = W.config_c2(n=9, seed=4)
    psi0 = evolve.all_ground_state(spec)
    tf = spec.sampling_times[-1]
    cuts = [0.0, 0.31337, 0.5, 1.70004, 3.0001, tf]
    refs = _oracle(spec, psi0, cuts)
    with engine.DevicePlan(spec) as plan:
        plan.set_state("all-ground")
        for a, b, ref in zip(cuts[:-1], cuts[1:], refs[1:]):
            st = plan.propagate(a, b)
            assert st["integrator"] == 3
            assert np.max(np.abs(plan.get_state()[0] - ref)) < STATE_TOL
        plan.set_state("all-ground")
        t_full = spec.sampling_times[480:561]          # across the kink of the waveform at 500 ns
        plan.set_state(_oracle(spec, psi0, [0.0, t_full[0]])[-1])
        applies = 0
        for a, b in zip(t_full[:-1], t_full[1:]):
            st = plan.propagate(float(a), float(b))
            assert st["integrator"] == 3
            applies += st["n_applies"]
        ref_end = _oracle(spec, psi0, [0.0, float(t_full[-1])])[-1]
        assert np.max(np.abs(plan.get_state()[0] - ref_end)) < STATE_TOL
        assert applies < 16 * (len(t_full) - 1)


@pytest.mark.parametrize("n", [5, 11, 12])
def test_noise_trajectory_batch_vs_oracle(engine, n):
    """C4-shaped batch: doppler offsets (zero on the padded last sample, i.e. det_k = theta + c_k M(t)) and per-atom
    amplitude factors -- the separable form the batched Taylor stage handles (blockIdx.y = trajectory)."""
    from oracle import evolve

    amp, det = W.blockade_sweep_waveforms(t_rise=80, t_sweep=200, t_fall=80)
    coords = W.disc_register(n, 14.0, 5.0, 5)
    base = W.ising_global_spec(coords, W.C6_LEVEL_60, amp, det, phase=0.4)
    rng = np.random.default_rng(n)
    specs = [W.noisy_trajectory_spec(base, coords, rng.normal(0, 1.5, n), max(0.0, rng.normal(1.0, 0.05)), 60.0)
             for _ in range(3)]
    tf = base.sampling_times[-1]
    psi0 = evolve.all_ground_state(base)
    with engine.DevicePlan(specs) as plan:
        plan.set_state("all-ground")
        st = plan.propagate(0.0, tf)
        got = plan.get_state().copy()
        plan.set_state("all-ground")
        st2 = plan.propagate(0.0, tf, integrator=2)
        lan = plan.get_state().copy()
    assert st["integrator"] == 3 and st2["integrator"] == 2
    assert np.max(np.abs(got - lan)) < STATE_TOL
    for b, spec in enumerate(specs):
        ref = _oracle(spec, psi0, [0.0, tf])[-1]
        assert np.max(np.abs(got[b] - ref)) < STATE_TOL, b


def test_not_applicable_falls_back_or_raises(engine):
    """per-qubit drives with moving phases: auto keeps the Magnus path, integrator 3 is refused loudly."""
    spec = random_local_spec(6, T=120, seed=3)
    with engine.DevicePlan(spec) as plan:
        plan.set_state("all-ground")
        st = plan.propagate(0.0, spec.sampling_times[-1])
        assert st["integrator"] in (1, 2)
        plan.set_state("all-ground")
        with pytest.raises(Exception, match="Taylor"):
            plan.propagate(0.0, spec.sampling_times[-1], integrator=3)


def test_tolerance_scaling(engine):
    """the a-priori budget follows ``tol``: a looser run costs fewer H-applies and both stay inside their bound."""
    from oracle import evolve

    spec = W.config_c2(n=10, seed=7)
    psi0 = evolve.all_ground_state(spec)
    tf = spec.sampling_times[-1]
    ref = _oracle(spec, psi0, [0.0, tf])[-1]
    out = {}
    with engine.DevicePlan(spec) as plan:
        for tol in (1e-5, 1e-8, 1e-10):
            plan.set_state("all-ground")
            st = plan.propagate(0.0, tf, integrator=3, tol=tol)
            out[tol] = (np.max(np.abs(plan.get_state()[0] - ref)), st["n_applies"])
    for tol, (err, _) in out.items():
        assert err < max(tol, 3e-10), (tol, err)     # the oracle itself is good to ~1e-10
    assert out[1e-5][1] < out[1e-8][1] < out[1e-10][1]


@pytest.mark.parametrize("name", ["c2_n20"])
def test_full_size_against_the_magnus_paths(engine, name):
    """C2 (N = 20, L2-resident) at its configured size: the Taylor run agrees with the Richardson-CF4 run at a 100x
    tighter tolerance to 1e-8, at a fraction of the H-applies.  (C5, N = 24, HBM-resident: the same comparison against
    three Magnus runs is tests/test_gpu_full_size.py::test_c5_whole_sequence.)"""
    spec = {"c2_n20": lambda: W.config_c2(n=20), "c5_n24": lambda: W.config_c5(n=24)}[name]()
    tf = spec.sampling_times[-1]
    with engine.DevicePlan(spec) as plan:
        plan.set_state("all-ground")
        st3 = plan.propagate(0.0, tf)
        got = plan.get_state()[0]
        assert st3["integrator"] == 3
        assert abs(plan.norm2()[0] - 1.0) < 1e-9
        plan.set_state("all-ground")
        st1 = plan.propagate(0.0, tf, integrator=1 if name == "c2_n20" else 2, tol=1e-10)
        ref = plan.get_state()[0]
    assert np.max(np.abs(got - ref)) < STATE_TOL
    assert st3["n_applies"] < 0.3 * st1["n_applies"]
    assert st3["n_applies"] < 2.0 * spec.total_duration_ns

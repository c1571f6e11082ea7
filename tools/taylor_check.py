"""Time-dependent Taylor propagator (integrator 3): parity against the DOP853 oracle at small N, against the
Chebyshev path at a tight tolerance at medium N, and the C2 / C5 timings."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W


def run(spec, **kw):
    with engine.DevicePlan(spec) as plan:
        plan.set_state("all-ground")
        t0 = time.time()
        st = plan.propagate(0.0, spec.sampling_times[-1], **kw)
        wall = time.time() - t0
        return plan.get_state()[0], st, wall


def main():
    from oracle import evolve
    sizes = [int(a) for a in sys.argv[1:]] or [6, 8, 10, 12, 14, 20]
    for n in sizes:
        spec = W.config_c2(n=n, seed=20 if n <= 12 else None) if n <= 22 else W.config_c5(n=n)
        T = spec.total_duration_ns
        got, st, wall = run(spec, integrator=3)
        row = {"n": n, "taylor": {"applies_per_ns": round(st["n_applies"] / T, 3), "steps": st["n_steps"], "gpu_ms": round(st["gpu_ms"], 2),
                                  "wall_ms": round(wall * 1e3, 2), "us_per_apply": round(st["gpu_ms"] * 1e3 / st["n_applies"], 2),
                                  "max_rho": round(st["max_rho"], 2), "err_bound": st["err_estimate"],
                                  "mean_step": round(st["mean_step_samples"], 1), "norm-1": abs(np.linalg.norm(got) - 1.0)}}
        if n <= 12:
            psi0 = evolve.all_ground_state(spec)
            from oracle.ref_hamiltonian import OracleHamiltonian
            ref = evolve.sesolve(OracleHamiltonian.from_spec(spec), psi0, [0.0, spec.sampling_times[-1]], rtol=1e-13, atol=1e-15)[-1]
            row["err_vs_oracle"] = float(np.linalg.norm(got - ref))
        if n <= 22:
            ref2, st2, wall2 = run(spec, integrator=1, tol=1e-10)
            row["err_vs_cheb_tol1e-10"] = float(np.linalg.norm(got - ref2))
            ref3, st3, wall3 = run(spec, integrator=1)
            row["cheb_default"] = {"applies_per_ns": round(st3["n_applies"] / T, 3), "gpu_ms": round(st3["gpu_ms"], 2),
                                   "err_vs_tight": float(np.linalg.norm(ref3 - ref2))}
        print(json.dumps(row), flush=True)


main()

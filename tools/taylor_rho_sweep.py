"""rho target sweep of the Taylor propagator on C2 (N=20): cost and error against a tight Chebyshev run."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
spec = W.config_c2(n=n)
T = spec.total_duration_ns
tf = spec.sampling_times[-1]
with engine.DevicePlan(spec) as plan:
    plan.set_state("all-ground")
    plan.propagate(0.0, tf, integrator=1, tol=1e-11)
    ref = plan.get_state()[0]
for rho in (8000, 10000, 12000, 14000, 16000, 18000):
    os.environ["PB200_TAYLOR_RHO_MILLI"] = str(rho)
    with engine.DevicePlan(spec) as plan:
        best = None
        for _ in range(2):
            plan.set_state("all-ground")
            st = plan.propagate(0.0, tf, integrator=3)
            best = st if best is None or st["gpu_ms"] < best["gpu_ms"] else best
        psi = plan.get_state()[0]
    print(json.dumps({"n": n, "rho_target": rho / 1e3, "applies_per_ns": round(best["n_applies"] / T, 3), "steps": best["n_steps"],
                      "gpu_ms": round(best["gpu_ms"], 2), "steps_per_s": round(T / best["gpu_ms"] * 1e3, 1),
                      "err_vs_cheb_tol1e-11": float(np.linalg.norm(psi - ref)), "norm2-1": float(np.vdot(psi, psi).real - 1.0)}), flush=True)

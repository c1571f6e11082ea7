import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W
from oracle import evolve
from oracle.ref_hamiltonian import OracleHamiltonian
for n in (8, 11):
    spec = W.config_c2(n=n, seed=20)
    psi0 = evolve.all_ground_state(spec); tf = spec.sampling_times[-1]
    ref = evolve.sesolve(OracleHamiltonian.from_spec(spec), psi0, [0.0, tf], rtol=1e-13, atol=1e-15)[-1]
    with engine.DevicePlan(spec) as plan:
        for K in (4, 8, 16, 32):
            for ex in (0, 1):
                plan.set_state("all-ground")
                st = plan.propagate(0.0, tf, tol=-1.0, max_step=K, extrapolate=ex)
                got = plan.get_state()[0]
                print(json.dumps({"n": n, "K": K, "extrap": ex, "err2": float(np.linalg.norm(got-ref)),
                                  "applies_per_ns": round(st["n_applies"]/4000,2), "norm-1": float(abs(np.linalg.norm(got)-1))}))

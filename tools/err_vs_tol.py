import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W
from oracle import evolve
from oracle.ref_hamiltonian import OracleHamiltonian
for n in (8,):
    spec = W.config_c2(n=n, seed=20)
    psi0 = evolve.all_ground_state(spec); tf = spec.sampling_times[-1]
    ref = evolve.sesolve(OracleHamiltonian.from_spec(spec), psi0, [0.0, tf], rtol=1e-13, atol=1e-15)[-1]
    with engine.DevicePlan(spec) as plan:
        for tol in (1e-9, 1e-8):
          for Wd in (8,):
            for kmax in (0,):
                plan.set_state("all-ground")
                st = plan.propagate(0.0, tf, tol=tol, refine_window=Wd, max_step=kmax)
                got = plan.get_state()[0]
                print(json.dumps({"n": n, "tol": tol, "W": Wd, "kmax": kmax, "err2": float(np.linalg.norm(got-ref)), "mean_step": round(st["mean_step_samples"],2),
                                  "applies_per_ns": round(st["n_applies"]/4000,2), "checks": st["n_checks"], "est": st["err_estimate"]}))
        # segment-wise error: where does it come from?
        plan.set_state("all-ground")
        H = OracleHamiltonian.from_spec(spec)
        cur = psi0
        for (a,b) in [(0,0.45),(0.45,0.55),(0.55,2.95),(2.95,3.05),(3.05,3.95),(3.95,4.0)]:
            refseg = evolve.sesolve(H, cur, [a, b], rtol=1e-13, atol=1e-15)[-1]
            plan.set_state(cur); plan.propagate(a, b, tol=1e-9)
            got = plan.get_state()[0]
            print("segment", a, b, float(np.linalg.norm(got-refseg)))
            cur = refseg

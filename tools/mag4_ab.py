"""A/B of the single-exponential Magnus step (PB200_MAG4) on C2 / C5-shaped runs + accuracy at N = 10 / 12."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W
from oracle import evolve
from oracle.ref_hamiltonian import OracleHamiltonian

refs = {}
for n in (10, 12):
    spec = W.config_c2(n=n)
    refs[n] = evolve.sesolve(OracleHamiltonian.from_spec(spec), evolve.all_ground_state(spec), [0.0, spec.sampling_times[-1]], rtol=1e-13, atol=1e-15)[-1]
for mag4 in (0, 1):
    os.environ["PB200_MAG4"] = str(mag4)
    row = {"mag4": mag4}
    for n in (10, 12):
        spec = W.config_c2(n=n)
        with engine.DevicePlan(spec) as p:
            p.set_state("all-ground"); st = p.propagate(0.0, spec.sampling_times[-1])
            row[f"err_n{n}"] = float(np.linalg.norm(p.get_state()[0] - refs[n]))
    for n, integ in ((20, 1), (20, 2), (22, 0)):
        spec = W.config_c2(n=n)
        with engine.DevicePlan(spec) as p:
            for rep in range(2):
                p.set_state("all-ground"); st = p.propagate(0.0, spec.sampling_times[-1], integrator=integ)
            row[f"n{n}_i{integ}"] = {"steps_per_s": round(4000 / (st["gpu_ms"] * 1e-3), 1), "applies_per_ns": round(st["n_applies"] / 4000, 3),
                                     "us_per_apply": round(st["gpu_ms"] * 1e3 / st["n_applies"], 2), "exps": st["n_exponentials"],
                                     "launches": st["n_launches"], "used": st["integrator"], "rejected": st["n_rejected"]}
    print(json.dumps(row), flush=True)

import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
spec = W.config_c3(n=n)
t0 = time.time()
with engine.DevicePlan(spec) as plan:
    plan.set_state("all-ground")
    ms, _ = plan.bench_apply(0.7, 10)
    for integ in (1, 2, 0):
        plan.set_state("all-ground")
        t1 = time.time()
        st = plan.propagate(0.0, spec.sampling_times[-1], integrator=integ)
        dt = time.time() - t1
        print(json.dumps({"n": n, "D": spec.hilbert_dim, "apply_us": round(ms / 10 * 1e3, 1), "integrator": integ, "used": int(st["integrator"]), "wall_s": round(dt, 2),
                          "steps_per_s": round(spec.total_duration_ns / dt, 1), "applies_per_ns": round(st["n_applies"] / spec.total_duration_ns, 2),
                          "norm2": float(plan.norm2()[0]), "occ_h": [round(float(x), 4) for x in plan.occupation(2)[0][:4]]}))

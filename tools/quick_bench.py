"""Ad-hoc GPU measurements (not the contract bench): apply kernel + C2 run."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W

def main():
    ns = [int(a) for a in sys.argv[1:]] or [20]
    for n in ns:
        spec = W.config_c2(n=n) if n <= 22 else W.config_c5(n=n)
        t0 = time.time()
        with engine.DevicePlan(spec) as plan:
            t1 = time.time()
            plan.set_state("all-ground")
            ms, launches = plan.bench_apply(1.0, 50)
            per = ms / 50
            D = spec.hilbert_dim
            print(json.dumps({"n": n, "plan_s": round(t1 - t0, 3), "apply_us": round(per * 1e3, 2),
                              "launches_per_apply": launches // 50,
                              "alg_GBs": round(40 * D / (per * 1e-3) / 1e9, 1)}))
            if True:
                for K, tol, integ in ((0, 0.0, 1), (0, 0.0, 2)):
                    plan.set_state("all-ground")
                    t2 = time.time()
                    st = plan.propagate(0.0, spec.sampling_times[-1], max_step=K, tol=tol, integrator=integ)
                    wall = time.time() - t2
                    T = spec.total_duration_ns
                    st.update({"K": K, "tol": tol, "wall_s": round(wall, 3), "steps_per_s": round(T / wall, 1),
                               "applies_per_ns": round(st["n_applies"] / T, 2),
                               "us_per_apply": round(st["gpu_ms"] * 1e3 / st["n_applies"], 2),
                               "norm2": float(plan.norm2()[0])})
                    print(json.dumps(st))
main()

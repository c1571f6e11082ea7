"""One C2 / C5 run of the Taylor propagator (integrator 3) for profiling: python tools/taylor_run.py N [reps]."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
integ = int(os.environ.get("INTEG", "3"))
spec = W.config_c2(n=n) if n <= 22 else W.config_c5(n=n)
T = spec.total_duration_ns
with engine.DevicePlan(spec) as plan:
    for r in range(reps):
        plan.set_state("all-ground")
        t0 = time.time()
        st = plan.propagate(0.0, spec.sampling_times[-1], integrator=integ)
        wall = time.time() - t0
        print(json.dumps({"n": n, "integrator": st["integrator"], "gpu_ms": round(st["gpu_ms"], 2), "wall_ms": round(wall * 1e3, 2),
                          "applies": st["n_applies"], "applies_per_ns": round(st["n_applies"] / T, 3), "steps": st["n_steps"],
                          "us_per_apply": round(st["gpu_ms"] * 1e3 / st["n_applies"], 2), "steps_per_s": round(T / wall, 1),
                          "norm2": float(plan.norm2()[0])}), flush=True)

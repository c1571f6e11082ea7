"""A/B of partner-sum forwarding (PB200_FWD) on C2-shaped Chebyshev runs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W
for n in [int(a) for a in sys.argv[1:]] or [18, 20, 22]:
    spec = W.config_c2(n=n)
    os.environ["PB200_FWD_MAX_N"] = "30"   # A/B beyond the automatic 17 <= N <= 19 window
    for fwd in (0, 1):
        os.environ["PB200_FWD"] = str(fwd)
        with engine.DevicePlan(spec) as plan:
            best = None
            for rep in range(3):
                plan.set_state("all-ground")
                st = plan.propagate(0.0, spec.sampling_times[-1], integrator=1)
                if best is None or st["gpu_ms"] < best["gpu_ms"]: best = st
            print(json.dumps({"n": n, "fwd": fwd, "gpu_ms": round(best["gpu_ms"], 2), "steps_per_s": round(4000 / best["gpu_ms"] * 1e3, 1),
                              "us_per_apply": round(best["gpu_ms"] * 1e3 / best["n_applies"], 2),
                              "us_per_launch": round(best["gpu_ms"] * 1e3 / best["n_launches"], 2),
                              "applies": best["n_applies"], "norm2": float(plan.norm2()[0])}), flush=True)

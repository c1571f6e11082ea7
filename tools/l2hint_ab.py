"""A/B of the L2 eviction-priority hints (PB200_L2HINT) on the stage kernel: bare H-apply, Clenshaw stages, Lanczos."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W

ns = [int(a) for a in sys.argv[1:]] or [20, 22, 24]
for n in ns:
    spec = W.config_c2(n=n) if n <= 22 else W.config_c5(n=n)
    D = spec.hilbert_dim
    for hint in (0, 1, 2, 3):
        os.environ["PB200_L2HINT"] = str(hint)
        with engine.DevicePlan(spec) as plan:
            plan.set_state("all-ground")
            ms, _ = plan.bench_apply(1.0, 30)
            row = {"n": n, "l2hint": hint, "bare_apply_us": round(ms / 30 * 1e3, 2),
                   "bare_GBs": round(40 * D / (ms / 30 * 1e-3) / 1e9, 1)}
            for integ, name in ((1, "cheb"), (2, "lanczos")):
                plan.set_state("all-ground")
                plan.propagate(0.0, 0.6)
                st = plan.propagate(0.6, 0.9, integrator=integ)
                row[f"{name}_us_per_apply"] = round(st["gpu_ms"] * 1e3 / max(st["n_applies"], 1), 2)
                row[f"{name}_applies"] = st["n_applies"]
            print(json.dumps(row), flush=True)

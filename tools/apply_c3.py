"""Bare H-apply timing of the 3-level (C3-shaped) register: python tools/apply_c3.py [n] [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
spec = W.config_c3(n=n)
with engine.DevicePlan(spec) as plan:
    plan.set_state("all-ground")
    ms, launches = plan.bench_apply(0.7, reps)
    D = spec.hilbert_dim
    print(json.dumps({"n": n, "D": D, "apply_us": round(ms / reps * 1e3, 1), "alg_GBs": round(40.0 * D * reps / ms / 1e6, 1),
                      "tiled": os.environ.get("PB200_TILED", "1"), "k": os.environ.get("PB200_TILED_K", "0")}))

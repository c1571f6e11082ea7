import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
spec = W.config_c2(n=n) if n <= 22 else W.config_c5(n=n)
with engine.DevicePlan(spec) as plan:
    plan.set_state("all-ground")
    print(plan.bench_apply(1.0, reps))

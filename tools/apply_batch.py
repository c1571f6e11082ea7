import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W
n = int(sys.argv[1]); B = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
spec = W.config_c2(n=n)
with engine.DevicePlan([spec]*B) as plan:
    plan.set_state("all-ground")
    print(plan.bench_apply(1.0, reps))

"""Target of compute-sanitizer: a C2-shaped N=12 propagate (Chebyshev and Lanczos) + sampling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W
spec = W.config_c2(n=12, seed=20, t_rise=40, t_sweep=80, t_fall=40)
with engine.DevicePlan(spec) as plan:
    for integ in (1, 2):
        plan.set_state("all-ground")
        st = plan.propagate(0.0, spec.sampling_times[-1], integrator=integ)
        print("integrator", integ, "launches", st["n_launches"], "norm2", float(plan.norm2()[0]))
    print(plan.sample(100, "r"))

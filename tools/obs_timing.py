"""Wall time of the on-device observable reductions at N = 20 (C2 state), per call, including the result copy."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W

spec = W.config_c2(n=20)
with engine.DevicePlan(spec) as plan:
    plan.set_state("all-ground")
    plan.propagate(0.0, 0.3)
    psi = plan.get_state()[0]
    out = {}
    for name, fn in (("occupation", lambda: plan.occupation(0)), ("correlation", lambda: plan.correlation(0)),
                     ("energy", lambda: plan.energy(0.3)), ("overlap", lambda: plan.overlap(psi)),
                     ("sample_1000", lambda: plan.sample(1000, "r")), ("get_state", lambda: plan.get_state())):
        fn()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        out[name + "_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    # host formulas on the downloaded state, for scale
    t0 = time.perf_counter()
    p = np.abs(psi) ** 2
    idx = np.arange(len(p))
    bits = [((idx >> (19 - k)) & 1) == 0 for k in range(20)]
    corr = np.array([[p[bits[i] & bits[j]].sum() for j in range(i, 20)] + [0.0] * i for i in range(20)])
    out["host_numpy_correlation_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    print(json.dumps(out))

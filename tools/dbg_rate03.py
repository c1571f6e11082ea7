import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np
from test_gpu_golden import load
from pulser_b200 import engine
spec, extra = load("orc_sampling_rate_03")
with engine.DevicePlan(spec) as plan:
    for ce, tol, cap in [(12, 0, None), (4, 0, None), (2, 0, None), (1, 0, None), (12, 1e-10, None), (12, -1.0, None)]:
        plan.set_state(extra["psi0"])
        st = plan.propagate(0.0, spec.sampling_times[-1], check_every=ce, tol=tol, max_step=(1 if tol < 0 else 0))
        got = plan.get_state()[0]
        print(json.dumps({"check_every": ce, "tol": tol, "err": float(np.max(np.abs(got-extra["orc_final"]))), "steps": st["n_steps"], "checks": st["n_checks"],
                          "mean_step": round(st["mean_step_samples"],2), "applies": st["n_applies"], "est": st["err_estimate"], "max_rho": round(st["max_rho"],2)}))

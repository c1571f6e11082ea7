"""Where the end-to-end time of one C2 Sequence goes (bench.py's e2e leg, phase by phase)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pulser_b200 import engine, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
spec = W.config_c2(n=n)
tf = spec.sampling_times[-1]; D = spec.hilbert_dim
psi0 = np.zeros(D, dtype=np.complex128); psi0[D - 1] = 1.0
psi0 = torch.from_numpy(psi0).pin_memory().numpy()
out_pinned = torch.empty(D, dtype=torch.complex128).pin_memory().numpy()
for it in range(4):
    t = [time.perf_counter()]
    plan = engine.DevicePlan(spec); t.append(time.perf_counter())
    plan.set_state(psi0); t.append(time.perf_counter())
    st = plan.propagate(0.0, tf); t.append(time.perf_counter())
    final = plan.get_state()[0]; t.append(time.perf_counter())
    plan.close(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(json.dumps({"iter": it, "plan_ms": round(d[0], 2), "set_state_ms": round(d[1], 2), "propagate_ms": round(d[2], 2),
                      "gpu_ms": round(st["gpu_ms"], 2), "get_state_ms": round(d[3], 2), "close_ms": round(d[4], 2),
                      "total_ms": round(float(np.sum(d)), 2), "applies": st["n_applies"], "launches": st["n_launches"]}))

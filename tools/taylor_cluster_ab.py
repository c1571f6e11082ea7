"""A/B of the Taylor stage kernel: single-CTA (partners through L2) against thread-block clusters of 4 / 8 / 16 tiles
(partners across the cluster bits through DSMEM).  python tools/taylor_cluster_ab.py N [N ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pulser_b200 import engine, workloads as W

for n in [int(a) for a in sys.argv[1:]] or [20]:
    spec = W.config_c2(n=n) if n <= 22 else W.config_c5(n=n)
    T = spec.total_duration_ns
    ref = None
    for cb in (0, 2, 3, 4):
        os.environ["PB200_TAYLOR_CLUSTER"] = str(cb)
        try:
            with engine.DevicePlan(spec) as plan:
                best = None
                for rep in range(3):
                    plan.set_state("all-ground")
                    st = plan.propagate(0.0, spec.sampling_times[-1], integrator=3)
                    best = st if best is None or st["gpu_ms"] < best["gpu_ms"] else best
                psi = plan.get_state()[0]
            if ref is None:
                ref = psi
            print(json.dumps({"n": n, "cluster_bits": cb, "gpu_ms": round(best["gpu_ms"], 2),
                              "us_per_apply": round(best["gpu_ms"] * 1e3 / best["n_applies"], 2),
                              "steps_per_s": round(T / (best["gpu_ms"] * 1e-3), 1),
                              "max_diff_vs_single_cta": float(np.max(np.abs(psi - ref)))}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"n": n, "cluster_bits": cb, "error": str(e)[:300]}), flush=True)

"""Partner-sum forwarding (PB200_FWD) against the single-pass stage kernel: device time per H-apply on
C2-shaped sequences.   python tools/fwd_sweep.py [n ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W

def main():
    ns = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 21, 22]
    for n in ns:
        kw = {} if n <= 20 else dict(t_rise=100, t_sweep=500, t_fall=200)
        spec = W.config_c2(n=n, **kw)
        T = spec.total_duration_ns
        for tb in ((11, 12) if n >= 21 else (11,)):
            for fwd in (0, 1):
                os.environ["PB200_FWD"] = str(fwd); os.environ["PB200_TILE_BITS"] = str(tb)
                with engine.DevicePlan(spec) as plan:
                    best = None
                    for rep in range(3):
                        plan.set_state("all-ground")
                        st = plan.propagate(0.0, spec.sampling_times[-1], integrator=1)
                        if best is None or st["gpu_ms"] < best["gpu_ms"]: best = st
                    print(json.dumps({"n": n, "tile_bits": tb, "fwd": fwd, "gpu_ms": round(best["gpu_ms"], 2),
                                      "steps_per_s": round(T / best["gpu_ms"] * 1e3, 1),
                                      "applies_per_ns": round(best["n_applies"] / T, 2),
                                      "us_per_apply": round(best["gpu_ms"] * 1e3 / best["n_applies"], 2),
                                      "us_per_launch": round(best["gpu_ms"] * 1e3 / best["n_launches"], 2),
                                      "alg_GBs": round(40.0 * spec.hilbert_dim * best["n_applies"] / best["gpu_ms"] / 1e6, 1),
                                      "norm2": float(plan.norm2()[0])}), flush=True)
main()

"""Partner-sum forwarding (PB200_FWD) against the single-pass stage kernel: device time per H-apply on
C2-shaped sequences.   python tools/fwd_sweep.py n [cfg ...]   cfg = fwd:flags:reg_bits:tile_bits"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_b200 import engine, workloads as W

def main():
    n = int(sys.argv[1])
    cfgs = sys.argv[2:] or ["0:0:3:11", "1:0:3:11"]
    kw = {} if n <= 20 else dict(t_rise=100, t_sweep=500, t_fall=200)
    spec = W.config_c2(n=n, **kw)
    T = spec.total_duration_ns
    for cfg in cfgs:
        fwd, flags, rb, tb = (int(x) for x in cfg.split(":"))
        os.environ.update({"PB200_FWD": str(fwd), "PB200_FWD_FLAGS": str(flags), "PB200_REG_BITS": str(rb),
                           "PB200_TILE_BITS": str(tb)})
        try:
            with engine.DevicePlan(spec) as plan:
                best = None
                for rep in range(2):
                    plan.set_state("all-ground")
                    st = plan.propagate(0.0, spec.sampling_times[-1], integrator=1)
                    if best is None or st["gpu_ms"] < best["gpu_ms"]: best = st
                print(json.dumps({"n": n, "cfg": cfg, "gpu_ms": round(best["gpu_ms"], 2),
                                  "steps_per_s": round(T / best["gpu_ms"] * 1e3, 1),
                                  "us_per_apply": round(best["gpu_ms"] * 1e3 / best["n_applies"], 2),
                                  "us_per_launch": round(best["gpu_ms"] * 1e3 / best["n_launches"], 2),
                                  "alg_GBs": round(40.0 * spec.hilbert_dim * best["n_applies"] / best["gpu_ms"] / 1e6, 1),
                                  "norm2": float(plan.norm2()[0])}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"n": n, "cfg": cfg, "error": str(e)[:200]}), flush=True)
main()

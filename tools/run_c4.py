"""C4: 16-atom noisy trajectories striped over the GPUs of one node, one all-reduce at the end.

    python tools/run_c4.py [n_traj] [batch]                      # 1 GPU
    torchrun --nproc-per-node 2 tools/run_c4.py [n_traj] [batch] # striped
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pulser_b200 import engine, parallel, workloads as W

def main():
    n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    specs = W.config_c4(n_traj)                      # same list on every rank (same seed)
    mine = [specs[j] for j in parallel.stripe(n_traj, rank, world)]
    n = specs[0].n_qudits; tf = specs[0].sampling_times[-1]
    rng = np.random.default_rng(1000 + rank)
    hist = np.zeros((1, 1 << n), dtype=np.int64)
    dens = np.zeros(n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stats = {"n_applies": 0, "gpu_ms": 0.0, "n_launches": 0}
    for b0 in range(0, len(mine), batch):
        chunk = mine[b0:b0 + batch]
        with engine.DevicePlan(chunk, device=local) as plan:
            plan.set_state("all-ground")
            st = plan.propagate(0.0, tf)
            for k in stats: stats[k] += st[k]
            probs = plan.probabilities()
        idx = np.arange(1 << n)
        for p in probs:                            # one shot per trajectory (samples_per_run = 1), r <-> '1'
            s = rng.choice(len(p), p=p / p.sum())
            hist[0, (1 << n) - 1 - s] += 1
            dens += np.array([p[((idx >> (n - 1 - k)) & 1) == 0].sum() for k in range(n)])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    hist = parallel.all_reduce_sum(hist)            # THE collective: bitstring histogram
    dens = parallel.all_reduce_sum(dens) / n_traj   # and the averaged Rydberg densities
    tmax = parallel.all_reduce_sum(np.array([dt])) if world == 1 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    if rank == 0:
        print(json.dumps({"config": "C4", "n_traj": n_traj, "world": world, "batch": batch, "wall_s": round(dt, 3),
                          "traj_per_s": round(n_traj / dt, 2), "steps_per_s_total": round(n_traj * 4000 / dt, 1),
                          "shots": int(hist.sum()), "mean_rydberg_density": round(float(dens.mean()), 6),
                          "applies_per_traj_step": round(stats["n_applies"] / 4000 / max(1, (len(mine) + batch - 1) // batch), 2),
                          "gpu_ms_rank0": round(stats["gpu_ms"], 1)}))
    if world > 1:
        dist.destroy_process_group()
main()

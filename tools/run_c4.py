"""C4 alone: the striped noise-trajectory leg of bench.py (device sampling, one all-reduce).

    PB200_BENCH_C4_TRAJ=256 python tools/run_c4.py                       # 1 GPU
    PB200_BENCH_C4_TRAJ=1024 torchrun --nproc-per-node 8 tools/run_c4.py # striped
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    from pulser_b200 import build
    build.build()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    out = bench.c4_leg(local, rank, world, dist, barrier, torch.cuda.current_stream())
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


main()

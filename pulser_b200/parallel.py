"""Multi-GPU plumbing: one process per GPU, trajectories striped over ranks,
exactly one all-reduce at the end.

The reference's noisy run is a serial Python loop over independent trajectory
solves (``pulser-simulation/pulser_simulation/simulation.py:885-915``) whose
per-time bitstring Counters are summed (``:848-861``); counts and observable
sums are additive, so each rank evolves its stripe and a single
``all_reduce(SUM)`` merges them (NCCL on GPUs, gloo in the CPU tests).  There
is no collective inside the time loop.
"""
from __future__ import annotations

from collections import Counter
from typing import Sequence

import numpy as np


def world_size() -> int:
    try:
        import torch.distributed as dist

        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    except Exception:  # pragma: no cover
        return 1


def rank() -> int:
    try:
        import torch.distributed as dist

        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:  # pragma: no cover
        return 0


def sync_numpy_random() -> None:
    """Put the global ``np.random`` stream of every rank in the same state: rank 0 draws a seed from its own stream
    and broadcasts it.  The striping below relies on every rank drawing the SAME list of noise trajectories
    (``HamiltonianData`` samples them from ``np.random`` in its constructor), while the ranks consume different
    numbers of sampling uniforms afterwards; called before every (re)draw and after every striped run."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    seed = torch.tensor([int(np.random.randint(0, 2**31 - 1)) if dist.get_rank() == 0 else 0], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        seed = seed.to(f"cuda:{torch.cuda.current_device()}")
    dist.broadcast(seed, src=0)
    np.random.seed(int(seed.item()))


def stripe(n_items: int, rank: int, world: int) -> list[int]:
    """Trajectory j -> rank j mod world."""
    return list(range(rank, n_items, world))


def counters_to_histogram(counters: Sequence[Counter], n_bits: int) -> np.ndarray:
    """[n_eval] Counters of bitstrings -> int64 [n_eval, 2^n_bits]."""
    hist = np.zeros((len(counters), 1 << n_bits), dtype=np.int64)
    for i, c in enumerate(counters):
        for bitstr, cnt in c.items():
            hist[i, int(bitstr, 2)] += cnt
    return hist


def histogram_to_counters(hist: np.ndarray, n_bits: int) -> list[Counter]:
    out = []
    for row in hist:
        nz = np.nonzero(row)[0]
        out.append(Counter({np.binary_repr(int(i), n_bits): int(row[i]) for i in nz}))
    return out


def all_reduce_sum(arr: np.ndarray, device: str | None = None) -> np.ndarray:
    """Sum an array over all ranks of the default process group (no-op if
    torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.to(device or f"cuda:{torch.cuda.current_device()}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def merge_trajectory_counts(local: Sequence[Counter], n_bits: int) -> list[Counter]:
    """The single collective of a striped noisy run."""
    hist = counters_to_histogram(local, n_bits)
    return histogram_to_counters(all_reduce_sum(hist), n_bits)

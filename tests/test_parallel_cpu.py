"""world_size-2 gloo test of the trajectory striping + final all-reduce."""
import os
import socket
from collections import Counter

import numpy as np
import pytest

from pulser_b200 import parallel


def test_stripe_partitions_everything():
    for n, w in [(10, 3), (1024, 8), (5, 8)]:
        parts = [parallel.stripe(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_histogram_roundtrip():
    cs = [Counter({"010": 3, "111": 1}), Counter({"000": 7})]
    h = parallel.counters_to_histogram(cs, 3)
    assert h.sum() == 11 and parallel.histogram_to_counters(h, 3) == cs


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_traj, n_bits, n_eval = 11, 3, 2
        rng_all = [np.random.default_rng(100 + j) for j in range(n_traj)]
        local = [Counter() for _ in range(n_eval)]
        for j in parallel.stripe(n_traj, rank, world):  # "evolve" my stripe
            for e in range(n_eval):
                idx = rng_all[j].integers(0, 1 << n_bits, size=5)
                local[e] += Counter(np.binary_repr(int(i), n_bits) for i in idx)
        merged = parallel.merge_trajectory_counts(local, n_bits)
        obs = parallel.all_reduce_sum(np.array([float(rank + 1), 2.0]))
        np.random.seed(1000 + rank)          # ranks start from different global streams ...
        parallel.sync_numpy_random()         # ... and leave with the same one
        draws = np.random.rand(3).tolist()
        q.put((rank, [dict(c) for c in merged], obs.tolist(), draws))
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_equals_serial():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # serial reference
    serial = [Counter() for _ in range(2)]
    for j in range(11):
        rng = np.random.default_rng(100 + j)
        for e in range(2):
            idx = rng.integers(0, 8, size=5)
            serial[e] += Counter(np.binary_repr(int(i), 3) for i in idx)
    for rank, merged, obs, draws in results:
        assert [Counter(m) for m in merged] == serial
        assert obs == [3.0, 4.0]
    assert results[0][3] == results[1][3]  # sync_numpy_random: identical streams afterwards

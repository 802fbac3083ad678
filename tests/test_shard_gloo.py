"""world_size-2 `gloo` test of the segment-parallel path (CPU): round-robin assignment, gather in segment order, every
rank ends with the same session digest.  Each rank "proves" its segments with the CPU oracle on a tiny circuit -- the
host-side sharding logic is what is under test (the GPU parity tests cover the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from zktls_b200 import shard

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)


def test_round_robin_assignment_is_a_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            seen = sorted(s for r in range(world) for s in shard.segments_for_rank(n, r, world))
            assert seen == list(range(n))
            for r in range(world):
                assert all(shard.owner_of(s, world) == r for s in shard.segments_for_rank(n, r, world))
    with pytest.raises(ValueError):
        shard.segments_for_rank(4, 2, 2)


def _prove(seg):
    from oracle import oracle as O
    from zktls_b200 import circuit, synth
    pr = O.Prover(circuit.syn_circuit(**SMALL).blob())
    io, code, data, accum = synth.trace_a(SMALL, 6, 100 + seg)
    pr.begin(6, io, code, data)
    return pr.finish(accum)


def _worker(rank, world, port, n_segments, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = {s: _prove(s) for s in shard.segments_for_rank(n_segments, rank, world)}
    results = shard.gather_results(local, n_segments, rank, world, dist)
    with open(os.path.join(out_dir, f"digest{rank}.txt"), "w") as f:
        f.write(shard.session_digest(results))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_the_same_session(tmp_path):
    n_segments, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n_segments, str(tmp_path)), nprocs=world, join=True)
    digests = [open(tmp_path / f"digest{r}.txt").read() for r in range(world)]
    assert digests[0] == digests[1]
    single = shard.session_digest(shard.gather_results({s: _prove(s) for s in range(n_segments)}, n_segments, 0, 1))
    assert digests[0] == single, "sharded result differs from the single-process result"


def _queue_worker(rank, world, port, n_segments, out_dir):
    import threading
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = shard.SharedSegmentQueue(n_segments, "t0")
    dist.barrier()
    taken = []
    def work():
        while True:
            i = q.take()
            if i is None:
                return
            taken.append(i)
            if rank == 0:          # rank 0 is the "slow" rank: it must end up with fewer segments, not hold the others back
                import time; time.sleep(0.02)
    ths = [threading.Thread(target=work) for _ in range(3)]          # three workers per rank, as bench.py runs them
    for t in ths: t.start()
    for t in ths: t.join()
    local = {s: np.full(4, s, dtype=np.uint32) for s in taken}
    results = shard.gather_results(local, n_segments, rank, world, dist)      # raises on a missing or doubly taken segment
    assert [int(r[0]) for r in results] == list(range(n_segments))
    with open(os.path.join(out_dir, f"taken{rank}.txt"), "w") as f:
        f.write(" ".join(str(s) for s in sorted(taken)))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_queue_hands_every_segment_to_exactly_one_rank(tmp_path):
    n_segments, world = 40, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_queue_worker, args=(world, port, n_segments, str(tmp_path)), nprocs=world, join=True)
    taken = [[int(x) for x in open(tmp_path / f"taken{r}.txt").read().split()] for r in range(world)]
    assert sorted(taken[0] + taken[1]) == list(range(n_segments))
    assert len(taken[0]) < len(taken[1]), "the slow rank should have taken fewer segments"


def test_gather_detects_missing_and_duplicate_segments():
    class FakeDist:
        def __init__(self, parts): self.parts = parts
        def all_gather_object(self, out, obj):
            for i, p in enumerate(self.parts): out[i] = p
    a = np.arange(4, dtype=np.uint32)
    with pytest.raises(RuntimeError, match="not proven"):
        shard.gather_results({0: a}, 3, 0, 2, FakeDist([{0: a}, {1: a}]))
    with pytest.raises(RuntimeError, match="twice"):
        shard.gather_results({0: a}, 2, 0, 2, FakeDist([{0: a}, {0: a, 1: a}]))

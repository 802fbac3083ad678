"""Segment-parallel sharding: which rank proves which continuation segment, and how results are gathered.

The reference proves a session's continuation segments one after another inside `prove_with_opts`
(/root/reference/crates/guest-prover-r0/src/prover.rs:90); they are independent until the lift/join recursion, so here
segment i goes to rank i mod G (one process per GPU) and there is NO collective on the data path.  The only exchange is
the gather of the results (seal + roots, ~0.27 MB per segment) for the recursion stage / the caller.
"""
import numpy as np


def segments_for_rank(n_segments, rank, world):
    """Round-robin assignment: rank r proves segments r, r + world, r + 2 world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_segments, world))


def owner_of(segment, world):
    return segment % world


def gather_results(local, n_segments, rank, world, dist=None):
    """local: {segment index: uint32 array (seal or digest)} proven by this rank.  Returns the full list in segment order
    on every rank (all_gather_object over the process group; plain dict merge when world == 1)."""
    if world == 1:
        return [np.asarray(local[i]) for i in range(n_segments)]
    parts = [None] * world
    dist.all_gather_object(parts, {int(k): np.asarray(v) for k, v in local.items()})
    merged = {}
    for part in parts:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"segment {k} proven twice")
            merged[k] = v
    missing = [i for i in range(n_segments) if i not in merged]
    if missing:
        raise RuntimeError(f"segments {missing} were not proven by any rank")
    return [merged[i] for i in range(n_segments)]


def session_digest(results):
    """Order-sensitive checksum of a session's per-segment results (what ranks compare to agree on the gathered set)."""
    import hashlib
    h = hashlib.sha256()
    for r in results:
        h.update(np.ascontiguousarray(r, dtype=np.uint32).tobytes())
    return h.hexdigest()

"""Segment-parallel sharding: which rank proves which continuation segment, and how results are gathered.

The reference proves a session's continuation segments one after another inside `prove_with_opts`
(/root/reference/crates/guest-prover-r0/src/prover.rs:90); they are independent until the lift/join recursion, so here
segment i goes to rank i mod G (one process per GPU) and there is NO collective on the data path.  The only exchange is
the gather of the results (seal + roots, ~0.27 MB per segment) for the recursion stage / the caller.
"""
import os
import numpy as np


def segments_for_rank(n_segments, rank, world):
    """Round-robin assignment: rank r proves segments r, r + world, r + 2 world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_segments, world))


def owner_of(segment, world):
    return segment % world


def gpu_numa_node(index):
    """NUMA node of GPU `index` (NVML's PCI bus id -> /sys/bus/pci/devices/<bdf>/numa_node), or None when the platform does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        for name in (bus.lower(), bus.lower()[4:] if len(bus) > 12 else bus.lower()):      # NVML pads the domain to 8 hex digits
            path = f"/sys/bus/pci/devices/{name}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read().strip())
                return node if node >= 0 else None
    except Exception:       # noqa: BLE001
        pass
    return None


def bind_rank_to_gpu_numa_node(index):
    """Best effort: run this rank's host threads on the CPUs of its GPU's NUMA node, so that the pinned trace buffers it allocates
    afterwards are first-touched there and the 1.17 GB upload of every segment does not cross the socket interconnect.  Tries the
    node's cpulist from sysfs first, NVML's affinity mask second.  Returns {"node", "cpus"} describing what was applied (cpus = 0:
    nothing was changed)."""
    cpus, node = set(), gpu_numa_node(index)
    try:
        if node is not None:
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus |= set(range(int(lo), int(hi or lo) + 1))
    except Exception:       # noqa: BLE001
        cpus = set()
    if not cpus:
        try:
            import pynvml
            pynvml.nvmlInit()
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        except Exception:   # noqa: BLE001
            cpus = set()
    try:
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:       # noqa: BLE001
        cpus = set()
    return {"node": node, "cpus": len(cpus)}


class SharedSegmentQueue:
    """The segments of one session handed out across ALL ranks in arrival order: `take()` returns the next unclaimed segment index, or
    None when the session is exhausted.  An atomic counter in torch.distributed's key-value store (the rendezvous store every process
    group already has) -- no collective, no data between ranks.  For boxes where the ranks are not equally fast end to end (on this
    pool GPUs 0-3 get 23 GB/s of host->device bandwidth and GPUs 4-7 35 GB/s when all eight upload): a static i mod N split waits for
    the slowest rank, a queue lets it take fewer segments.  `take()` may be called from several threads of a rank."""

    def __init__(self, n_segments, name, store=None):
        import threading
        if store is None:
            import torch.distributed as dist
            store = dist.distributed_c10d._get_default_store()
        self.n, self.key, self.store, self.lock = int(n_segments), f"zkb200/queue/{name}", store, threading.Lock()

    def take(self):
        with self.lock:
            i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n else None


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

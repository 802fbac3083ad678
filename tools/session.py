"""BASELINE.json configs[3]/[4] stand-in: a multi-segment session -- S independent SYN-280 segments of 2^po2
cycles -- proven segment-parallel over the ranks of one box (torchrun, one process per GPU, `--inflight` segments per GPU),
results gathered in segment order.  Prints one JSON line: wall time for the session, segments/s, and the session digest.

    python tools/session.py --segments 64
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/session.py --segments 64
"""
import argparse, hashlib, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from zktls_b200 import circuit, shard
from zktls_b200.hal import B200Hal, Buffer
from zktls_b200.prover import SegmentProver

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=64); ap.add_argument("--po2", type=int, default=20); ap.add_argument("--inflight", type=int, default=2)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
P = 2013265921; n = 1 << a.po2; shp = circuit.SYN280
blob = circuit.syn_circuit(**shp).blob()
hals = [B200Hal(local) for _ in range(a.inflight)]
provers = [SegmentProver(h, blob) for h in hals]
g = torch.Generator(device=dev); g.manual_seed(0xB2000000)      # same trace on every rank; segments differ by their io words
ts = [torch.randint(0, P, (shp[k] * n,), device=dev, dtype=torch.int64, generator=g).to(torch.int32) for k in ("code_cols", "data_cols", "accum_cols")]
bufs = [[Buffer(h, t.data_ptr(), t.numel(), 1, owner=t) for t in ts] for h in hals]
mine = shard.segments_for_rank(a.segments, rank, world)
io_of = lambda seg: np.random.default_rng(1000 + seg).integers(0, P, size=shp["out_size"], dtype=np.uint32)
for w, pr in enumerate(provers):
    pr.prove(a.po2, io_of(0), *bufs[w])                           # warm-up (tables, JIT module, memory pool)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
local_res = {}
def work(w):
    for seg in mine[w::a.inflight]:
        seal = provers[w].prove(a.po2, io_of(seg), *bufs[w])
        local_res[seg] = np.frombuffer(hashlib.sha256(seal.tobytes()).digest(), dtype=np.uint32).copy()
t0 = time.time()
ths = [threading.Thread(target=work, args=(w,)) for w in range(a.inflight)]
for th in ths: th.start()
for th in ths: th.join()
torch.cuda.synchronize()
results = shard.gather_results(local_res, a.segments, rank, world, dist if world > 1 else None)    # the only exchange: 32 bytes per segment
dt = time.time() - t0
t = torch.tensor([dt], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"session of {a.segments} SYN-280 segments, 2^{a.po2} cycles each", "n_gpus": world, "inflight_per_gpu": a.inflight,
                      "wall_s": float(t.item()), "segments_per_s": a.segments / float(t.item()), "session_digest": shard.session_digest(results)}), flush=True)
for pr in provers: pr.close()
if world > 1:
    dist.destroy_process_group()

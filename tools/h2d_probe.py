"""Aggregate pinned-host -> device copy bandwidth with one process per GPU (torchrun): the ceiling of bench.py's end-to-end arm,
whose every segment uploads 1.17 GB.  Each rank copies its own pinned buffer to its GPU `reps` times between two barriers;
rank 0 prints per-rank and aggregate GB/s.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/h2d_probe.py"""
import os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
bound = None
if "--bind" in sys.argv:        # run (and first-touch the pinned buffer) on the CPUs local to this rank's GPU, as bench.py does
    from zktls_b200.shard import bind_rank_to_gpu_numa_node
    bound = bind_rank_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 1174405120          # one SYN-280 2^20 segment's three trace groups
h = torch.empty(nbytes // 4, dtype=torch.int32).pin_memory(); h.fill_(rank + 1)
d = torch.empty_like(h, device=dev)
streams = [torch.cuda.Stream() for _ in range(3)]
def run(reps, concurrent):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.time()
    for r in range(reps):
        if concurrent:          # three chunks on three streams, like three workers' uploads
            third = h.numel() // 3
            for k, s in enumerate(streams):
                with torch.cuda.stream(s):
                    d[k * third:(k + 1) * third].copy_(h[k * third:(k + 1) * third], non_blocking=True)
        else:
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.time()
    if world > 1: dist.barrier()
    return reps * nbytes / (t1 - t0) / 1e9
for concurrent in (False, True):
    run(2, concurrent)
    g = run(10, concurrent)
    t = torch.tensor([g], device=dev, dtype=torch.float64)
    allg = [torch.zeros_like(t) for _ in range(world)]
    if world > 1: dist.all_gather(allg, t)
    else: allg = [t]
    if rank == 0:
        per = [float(x.item()) for x in allg]
        print(json.dumps({"ranks": world, "bound": bound, "three_streams": concurrent, "per_rank_GBps": [round(x, 1) for x in per], "aggregate_GBps": round(sum(per), 1)}), flush=True)
if world > 1: dist.destroy_process_group()

"""Latency of a small device->host read (tiny kernel + 32-byte async copy + stream synchronize: one Fiat-Shamir round trip of the
prover) while ANOTHER stream of the same process uploads 940 MB traces back to back, as a function of the upload's chunk size.
    python tools/ubench/d2h_under_h2d.py"""
import time, json, torch
dev = torch.device("cuda", 0)
words = 224 << 20
h = torch.empty(words, dtype=torch.int32).pin_memory(); h.fill_(7)
d = torch.empty(words, dtype=torch.int32, device=dev)
small_d = torch.zeros(8, dtype=torch.int32, device=dev); small_h = torch.empty(8, dtype=torch.int32).pin_memory()
up, work = torch.cuda.Stream(), torch.cuda.Stream()

def round_trips(n):
    lat = []
    with torch.cuda.stream(work):
        for _ in range(n):
            t0 = time.perf_counter()
            small_d.add_(1)
            small_h.copy_(small_d, non_blocking=True)
            work.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
    return lat

def upload(reps, chunk_mb):
    with torch.cuda.stream(up):
        for _ in range(reps):
            if chunk_mb == 0:
                d.copy_(h, non_blocking=True)
            else:
                cw = chunk_mb << 18
                for off in range(0, words, cw):
                    d[off:off + cw].copy_(h[off:off + cw], non_blocking=True)

round_trips(50); upload(1, 0); torch.cuda.synchronize()
base = round_trips(200)
print(json.dumps({"uploads": "none", "round_trip_ms_median": round(sorted(base)[100], 4), "max": round(max(base), 3)}))
for chunk in (0, 256, 64, 16, 4, 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    upload(6, chunk)
    t_enq = (time.perf_counter() - t0) * 1e3
    lat = []
    while not up.query():
        lat += round_trips(1)
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) * 1e3
    lat.sort()
    print(json.dumps({"upload_chunk_mb": chunk or "whole (940 MB)", "enqueue_ms": round(t_enq, 2), "six_uploads_ms": round(t_all, 1), "GBps": round(6 * words * 4 / t_all / 1e6, 1),
                      "round_trips": len(lat), "round_trip_ms_median": round(lat[len(lat) // 2], 4) if lat else None, "p99": round(lat[int(len(lat) * 0.99)], 3) if lat else None,
                      "max": round(lat[-1], 3) if lat else None}), flush=True)

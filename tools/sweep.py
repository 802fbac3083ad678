"""Kernel sweep (BASELINE.json configs[2]): BabyBear iNTT / x4 LDE over 2^16..2^24 rows x 64..512 columns and Poseidon2
hash_rows / Merkle build over 2^16..2^24 leaves, timed with CUDA events on the ctx stream, against the HBM roofline
(algorithmic bytes) and the measured INT32 modmul ceiling.  Prints one JSON object per shape.

    python tools/sweep.py [--max-po2 24] [--reps 3] > profiles/rN_sweep.jsonl
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zktls_b200.hal import B200Hal

HBM = 6650.0          # GB/s, fallback peak (B200_PROFILING.md) unless MEASURED_PEAKS.json exists
MODMUL = 18.61e12 / 5     # Montgomery products per second at the hardware multiplier-pipe slot peak (64 lanes x 148 SMs x 1.965 GHz, 5 slots each)
ap = argparse.ArgumentParser(); ap.add_argument("--max-po2", type=int, default=24); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-bytes", type=float, default=60e9)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = os.path.join(root, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    HBM = json.load(open(pk)).get("hbm_gbs", HBM)
hal = B200Hal(0)


def timed(fn, reps):
    fn(); fn(); hal.sync()
    hal.timer_start()
    for _ in range(reps): fn()
    return hal.timer_stop() / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


rng = np.random.default_rng(0xB2001000)
for po2 in range(16, a.max_po2 + 1, 2):
    n = 1 << po2
    for cols in (64, 128, 256, 512):
        if 4 * n * cols * 5 > a.max_bytes: continue
        buf = hal.alloc_elem(cols * n); big = hal.alloc_elem(cols * 4 * n)
        # seeded non-trivial contents: one random column replicated (values do not affect timing; parity is tested elsewhere)
        col = rng.integers(0, 2013265921, size=n, dtype=np.uint32)
        for c in range(min(cols, 4)): buf.slice(c * n, n).copy_from(col)
        ms = timed(lambda: hal.batch_interpolate_ntt_zk_shift(buf, cols), a.reps)
        by = 8 * n * cols; bf = cols * (n // 2) * po2 + cols * n
        emit(op="intt_zk_shift", po2=po2, cols=cols, ms=ms, alg_GBps=by / ms / 1e6, hbm_frac=by / ms / 1e6 / HBM, modmul_per_s=bf / ms * 1e3, int32_frac=bf / ms * 1e3 / MODMUL)
        ms = timed(lambda: hal.batch_expand_into_evaluate_ntt(big, buf, cols, 2), a.reps)
        by = 20 * n * cols; bf = cols * 2 * n * po2 + cols * 4 * n
        emit(op="lde_x4", po2=po2, cols=cols, ms=ms, alg_GBps=by / ms / 1e6, hbm_frac=by / ms / 1e6 / HBM, modmul_per_s=bf / ms * 1e3, int32_frac=bf / ms * 1e3 / MODMUL)
        del buf, big
for po2 in range(16, a.max_po2 + 1, 2):
    rows = 1 << po2
    for cols in (16, 64, 256):
        if 4 * rows * cols > a.max_bytes: continue
        m = hal.alloc_elem(rows * cols); d = hal.alloc_digest(rows)
        ms = timed(lambda: hal.hash_rows(d, m), a.reps)
        by = 4 * rows * cols + 32 * rows; perms = rows * ((cols + 15) // 16)
        emit(op="hash_rows", po2=po2, cols=cols, ms=ms, alg_GBps=by / ms / 1e6, hbm_frac=by / ms / 1e6 / HBM, perms_per_s=perms / ms * 1e3, modmul_per_s=1356 * perms / ms * 1e3, int32_frac=1356 * perms / ms * 1e3 / MODMUL)
        del m, d
    nodes = hal.alloc_digest(2 * rows)
    ms = timed(lambda: hal.merkle_build(nodes, rows), a.reps)
    by = 96 * (rows - 1); perms = rows - 1
    emit(op="merkle_build", po2=po2, ms=ms, alg_GBps=by / ms / 1e6, hbm_frac=by / ms / 1e6 / HBM, perms_per_s=perms / ms * 1e3, modmul_per_s=1356 * perms / ms * 1e3, int32_frac=1356 * perms / ms * 1e3 / MODMUL)
    del nodes
# poseidon_254 suite (first slice of SURVEY 8f-4): ~830 256-bit Montgomery products (8 x 8 limb products each) per permutation
for po2 in (16, 18, 20):
    rows = 1 << po2
    nodes = hal.alloc_digest(2 * rows)
    ms = timed(lambda: hal.p254_merkle_build(nodes, rows), a.reps)
    emit(op="poseidon254_merkle_build", po2=po2, ms=ms, perms_per_s=(rows - 1) / ms * 1e3)
    del nodes
    m = hal.alloc_elem(rows * 16); d = hal.alloc_digest(rows)
    ms = timed(lambda: hal.p254_hash_rows(d, m), a.reps)
    emit(op="poseidon254_hash_rows", po2=po2, cols=16, ms=ms, perms_per_s=rows / ms * 1e3, note="provisional row packing: 16 columns = 2 words = 1 permutation")
    del m, d
hal.close()

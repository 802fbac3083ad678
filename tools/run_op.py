"""Runs one operator at benchmark shape on cuda:0 (the short command wrapped by `ncu --set full`)."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zktls_b200.hal import B200Hal
ap = argparse.ArgumentParser(); ap.add_argument("op"); ap.add_argument("--po2", type=int, default=20); ap.add_argument("--cols", type=int, default=224); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
hal = B200Hal(0); n = 1 << a.po2
if a.op == "hash_rows":
    rows = 4 * n; m = hal.alloc_elem(rows * a.cols); d = hal.alloc_digest(rows)
    hal.hash_rows(d, m); hal.sync(); hal.timer_start()
    for _ in range(a.reps): hal.hash_rows(d, m)
    print(f"hash_rows {a.cols} x 2^{a.po2 + 2}: {hal.timer_stop() / a.reps:.3f} ms")
elif a.op == "merkle":
    rows = 4 * n; d = hal.alloc_digest(2 * rows)
    hal.merkle_build(d, rows); hal.sync(); hal.timer_start()
    for _ in range(a.reps): hal.merkle_build(d, rows)
    print(f"merkle_build 2^{a.po2 + 2} leaves: {hal.timer_stop() / a.reps:.3f} ms  env", {k: v for k, v in os.environ.items() if k.startswith("ZKB_")})
elif a.op == "intt":
    b = hal.alloc_elem(a.cols * n)
    hal.batch_interpolate_ntt_zk_shift(b, a.cols); hal.sync(); hal.timer_start()
    for _ in range(a.reps): hal.batch_interpolate_ntt_zk_shift(b, a.cols)
    ms = hal.timer_stop() / a.reps; print(f"intt+shift {a.cols} x 2^{a.po2}: {ms:.3f} ms  {8 * n * a.cols / ms / 1e6:.0f} GB/s  env", {k: v for k, v in os.environ.items() if k.startswith("ZKB_")})
elif a.op == "lde":
    b = hal.alloc_elem(a.cols * n); o = hal.alloc_elem(a.cols * 4 * n)
    hal.batch_expand_into_evaluate_ntt(o, b, a.cols, 2); hal.sync(); hal.timer_start()
    for _ in range(a.reps): hal.batch_expand_into_evaluate_ntt(o, b, a.cols, 2)
    ms = hal.timer_stop() / a.reps; print(f"lde x4 {a.cols} x 2^{a.po2}: {ms:.3f} ms  {20 * n * a.cols / ms / 1e6:.0f} GB/s  env", {k: v for k, v in os.environ.items() if k.startswith("ZKB_")})
elif a.op == "eval_check":
    import os
    from zktls_b200 import circuit
    shape = circuit.SYN280; blob = circuit.syn_circuit(**shape).blob(); dom = 4 * n
    acc = hal.alloc_elem(shape["accum_cols"] * dom); code = hal.alloc_elem(shape["code_cols"] * dom); data = hal.alloc_elem(shape["data_cols"] * dom)
    chk = hal.alloc_elem(4 * dom); mixg = np.arange(shape["mix_size"], dtype=np.uint32); outg = np.arange(shape["out_size"], dtype=np.uint32); pm = np.array([5, 6, 7, 8], np.uint32)
    hal.eval_check(chk, blob, acc, code, data, mixg, outg, pm, a.po2); hal.sync()
    hal.timer_start()
    for _ in range(a.reps): hal.eval_check(chk, blob, acc, code, data, mixg, outg, pm, a.po2)
    print("eval_check mode", os.environ.get("ZKB_EVAL_CHECK", "jit"), "minblocks", os.environ.get("ZKB_EC_MINBLOCKS", "4"), "ms/launch", hal.timer_stop() / a.reps)
hal.timer_start(); hal.sync(); print(a.op, "done, launches", hal.kernel_launches())

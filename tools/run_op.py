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
    for _ in range(a.reps): hal.hash_rows(d, m)
elif a.op == "merkle":
    rows = 4 * n; d = hal.alloc_digest(2 * rows)
    for _ in range(a.reps): hal.merkle_build(d, rows)
elif a.op == "intt":
    b = hal.alloc_elem(a.cols * n)
    for _ in range(a.reps): hal.batch_interpolate_ntt_zk_shift(b, a.cols)
elif a.op == "lde":
    b = hal.alloc_elem(a.cols * n); o = hal.alloc_elem(a.cols * 4 * n)
    for _ in range(a.reps): hal.batch_expand_into_evaluate_ntt(o, b, a.cols, 2)
hal.timer_start(); hal.sync(); print(a.op, "done, launches", hal.kernel_launches())

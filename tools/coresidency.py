"""Do a multiplier-bound kernel (hash_rows) and an ALU/LSU/HBM-bound one (LDE, eval_check) from two streams share the SMs, and does
that pay?  Times, on cuda:0: each op alone, the two back to back on one stream, and the two concurrently on two ctxs (streams).
Run once per ZKB_HASH_CTAS_PER_SM setting (0 = default grid, 4 / 5 = persistent CTAs per SM that leave registers free)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zktls_b200 import circuit
from zktls_b200.hal import B200Hal

po2, cols = 20, 224
n = 1 << po2
A, B = B200Hal(0), B200Hal(0)
mat = A.alloc_elem(4 * n * cols); dig = A.alloc_digest(4 * n)
src = B.alloc_elem(n * cols); big = B.alloc_elem(4 * n * cols)
shape = circuit.SYN280; blob = circuit.syn_circuit(**shape).blob(); dom = 4 * n
acc = B.alloc_elem(shape["accum_cols"] * dom); code = B.alloc_elem(shape["code_cols"] * dom)
chk = B.alloc_elem(4 * dom); mixg = np.arange(shape["mix_size"], dtype=np.uint32); outg = np.arange(shape["out_size"], dtype=np.uint32); pm = np.array([5, 6, 7, 8], np.uint32)

def hash_(k):
    for _ in range(k): A.hash_rows(dig, mat)
def lde(k):
    for _ in range(k): B.batch_expand_into_evaluate_ntt(big, src, cols, 2)
def ec(k):
    for _ in range(k): B.eval_check(chk, blob, acc, code, big, mixg, outg, pm, po2)

def wall(fns):
    A.sync(); B.sync(); t0 = time.time()
    ths = [threading.Thread(target=f) for f in fns]
    for t in ths: t.start()
    for t in ths: t.join()
    A.sync(); B.sync()
    return (time.time() - t0) * 1e3

hash_(2); lde(2); ec(2); A.sync(); B.sync()
H = 4
t_hash = wall([lambda: hash_(H)]) / H
for name, op, reps in (("lde", lde, 6 * H), ("eval_check", ec, 14 * H)):
    t_op = wall([lambda: op(reps)]) / reps
    t_both = wall([lambda: hash_(H), lambda: op(reps)])
    serial = H * t_hash + reps * t_op
    print(f"ZKB_HASH_CTAS_PER_SM={os.environ.get('ZKB_HASH_CTAS_PER_SM', '0')}: hash_rows {t_hash:.2f} ms, {name} {t_op:.3f} ms; {H} x hash + {reps} x {name}: "
          f"serial {serial:.1f} ms, concurrent {t_both:.1f} ms ({100 * (1 - t_both / serial):+.1f} % saved)")

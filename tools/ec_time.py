"""Times eval_check alone (CUDA events) for the benchmark circuit or SYN-HEAVY over the 2^22-point domain of a 2^20-cycle segment.
usage: python tools/ec_time.py [syn280|heavy] [po2]   (environment: ZKB_EC_* knobs of csrc/k_eval_jit.cu)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zktls_b200 import circuit
from zktls_b200.hal import B200Hal
which = sys.argv[1] if len(sys.argv) > 1 else "syn280"
po2 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
b = circuit.syn_heavy_circuit() if which == "heavy" else circuit.syn_circuit(**circuit.SYN280)
blob = b.blob()
hal = B200Hal(0)
dom = 4 << po2
bufs = [hal.alloc_elem(n * dom) for n in b.group_size]
chk = hal.alloc_elem(4 * dom)
mg, og, pm = np.arange(1, 1 + b.mix_size, dtype=np.uint32), np.arange(1, 1 + b.out_size, dtype=np.uint32), np.arange(3, 7, dtype=np.uint32)
t0 = time.time()
hal.eval_check(chk, blob, *bufs, mg, og, pm, po2); hal.sync()
first = time.time() - t0
hal.eval_check(chk, blob, *bufs, mg, og, pm, po2)
reps = int(os.environ.get("EC_TIME_REPS", "3"))          # many repetitions: the sustained (power-capped) rate instead of the burst rate
hal.timer_start()
for _ in range(reps):
    hal.eval_check(chk, blob, *bufs, mg, og, pm, po2)
ms = hal.timer_stop() / reps
knobs = {k: v for k, v in os.environ.items() if k.startswith("ZKB_EC_")}
print(f"eval_check {which} po2 {po2}: {ms:.3f} ms  (first call incl. JIT / cache load {first:.1f} s)  {knobs}")

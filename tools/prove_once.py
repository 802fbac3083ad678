"""Proves SYN-280 segments on cuda:0 (device-resident Trace A) -- the short command wrapped by ncu for launch lists."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zktls_b200 import circuit
from zktls_b200.hal import B200Hal, Buffer
from zktls_b200.prover import SegmentProver
ap = argparse.ArgumentParser(); ap.add_argument("--po2", type=int, default=20); ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
P = 2013265921; n = 1 << a.po2; shape = circuit.SYN280
dev = torch.device("cuda", 0); hal = B200Hal(0); pr = SegmentProver(hal, circuit.syn_circuit(**shape).blob())
g = torch.Generator(device=dev); g.manual_seed(1)
ts = [torch.randint(0, P, (shape[k] * n,), device=dev, dtype=torch.int64, generator=g).to(torch.int32) for k in ("code_cols", "data_cols", "accum_cols")]
bufs = [Buffer(hal, t.data_ptr(), t.numel(), 1, owner=t) for t in ts]
io = np.arange(shape["out_size"], dtype=np.uint32)
torch.cuda.synchronize()
for _ in range(a.steps):
    seal = pr.prove(a.po2, io, *bufs)
print("seal words", seal.size, "launches", hal.kernel_launches())

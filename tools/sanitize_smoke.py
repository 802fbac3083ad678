"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every operator once at small sizes + one
segment proof through the staged pipeline + verification.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zktls_b200 import circuit, synth
from zktls_b200.hal import B200Hal
from zktls_b200.prover import SegmentProver, verify_segment, control_id
P = 2013265921
rng = np.random.default_rng(1)
fp = lambda n: rng.integers(0, P, size=n, dtype=np.uint32)
hal = B200Hal(0)
for po2, count in ((4, 3), (10, 2), (13, 2), (16, 1)):
    b = hal.copy_from_elem(fp(count << po2)); hal.batch_interpolate_ntt_zk_shift(b, count); hal.batch_bit_reverse(b, count)
    o = hal.alloc_elem(count << (po2 + 2)); hal.batch_expand_into_evaluate_ntt(o, b, count, 2)
    d = hal.alloc_digest(1 << (po2 + 2)); hal.hash_rows(d, o)
    nodes = hal.alloc_digest(2 << (po2 + 2)); hal.merkle_build(nodes, 1 << (po2 + 2))
pd = hal.copy_from_extelem(fp(4 * 5000)); hal.poly_divide(pd, fp(4)); hal.prefix_products(pd)
xs = hal.copy_from_extelem(fp(4 * 5)); out = hal.alloc_extelem(5)
hal.batch_evaluate_any(hal.copy_from_elem(fp(3 << 12)), 3, hal.copy_from_u32(np.array([0, 1, 2, 2, 0], np.uint32)), xs, out)
src = hal.copy_from_elem(fp(4096)); gr = hal.alloc_elem(3 * 16); hal.gather_rows(gr, src, np.array([0, 7, 255], dtype=np.uint32), 16, 256)
shape = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
blob = circuit.syn_circuit(**shape).blob()
pr = SegmentProver(hal, blob)
po2 = 10
io, code, data = synth.trace_b_code_data(shape, po2, 3)
code_m, data_m = synth.to_mont(code), synth.to_mont(data)
mix = pr.begin(po2, io, code_m, data_m)
accum_m = synth.to_mont(synth.trace_b_accum(shape, po2, 3, code, data, io, mix))
seal = pr.finish(accum_m); verify_segment(blob, seal, control_id(po2, pr.roots()[0]))
tr = synth.trace_a(shape, po2, 5)
pr.stage(po2, *tr[1:]); pr.stage(po2, *tr[1:]); s1 = pr.prove_staged(tr[0]); s2 = pr.prove_staged(tr[0])
assert np.array_equal(s1, s2)
pr.close()
# round 2: device-side accumulate (witness program as data), the heavy-circuit forms of eval_check, the poseidon_254 kernels
d_acc = hal.copy_from_elem(accum_m)
hal.accumulate(blob, d_acc, hal.copy_from_elem(code_m), hal.copy_from_elem(data_m), mix, io, po2)
red = dict(accum_cols=6, code_cols=6, data_cols=12, mix_size=5, out_size=4, majors=2, fanout=(2, 2, 2), leaf_constraints=6)
hb = circuit.syn_heavy_circuit(**red); hblob = hb.blob()
for form in ("compact", "flat"):
    os.environ["ZKB_EC_FORM"] = form
    dom = 4 << 7
    chk = hal.alloc_elem(4 * dom)
    hal.eval_check(chk, hblob, *[hal.copy_from_elem(fp(n * dom)) for n in hb.group_size], fp(5), fp(4), fp(4), 7)
os.environ.pop("ZKB_EC_FORM")
nodes = hal.copy_from_digest(fp(2 * 64 * 8) & 0x0fffffff); hal.p254_merkle_build(nodes, 64)
pd = hal.alloc_digest(33); hal.p254_hash_rows(pd, hal.copy_from_elem(fp(33 * 9)))
hal.sync(); hal.close()
print("sanitize smoke ok")

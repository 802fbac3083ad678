"""GPU parity AT THE BENCHMARK SIZES (BASELINE.json configs[1] and configs[2]).

* configs[1] literally: the SYN-280 segment of 2^20 cycles that bench.py times (Trace A, the seed of bench.py's
  cpu_baseline leg) is proven on the GPU through the C-ABI and by the CPU oracle; the 7 Merkle / FRI roots and the
  whole seal must be identical word for word.
* configs[2] above 2^20 ("larger: sampled rows/columns + root vs the oracle"): hash_rows over the 224 x 2^22 data-group
  matrix (sampled rows), the 224-column iNTT + x4 LDE 2^20 -> 2^22 (sampled columns, each compared in full), and the
  three-pass NTT plans (po2 23..26), all against the oracle -- not against the repo's own level kernels.

32-bit index arithmetic changes character at these sizes (column offsets beyond 2^26 elements, matrices beyond 2^32 bytes),
which is why the small-shape tests in test_hal_parity.py are not enough.
"""
import numpy as np
import pytest

from zktls_b200 import circuit, synth

pytestmark = pytest.mark.gpu

P = 2013265921
BENCH_SEED = 0xB200          # bench.py cpu_baseline(): synth.trace_a(SYN280, po2, 0xB200)


@pytest.fixture(scope="module")
def hal():
    from zktls_b200.hal import B200Hal
    h = B200Hal(0)
    yield h
    h.close()


@pytest.fixture(scope="module")
def torch_dev():
    import torch
    return torch, torch.device("cuda", 0)


def device_random_fp(torch_dev, count, seed):
    torch, dev = torch_dev
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randint(0, P, (count,), device=dev, dtype=torch.int64, generator=g).to(torch.int32)     # < 2^31: the same bits as u32
    torch.cuda.synchronize()          # the ctx has its own stream
    return t


def as_buffer(hal, t):
    from zktls_b200.hal import Buffer
    return Buffer(hal, t.data_ptr(), t.numel(), 1, owner=t)


def to_u32(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.slow
def test_benchmark_segment_seal_and_roots_match_the_oracle(hal, oracle):
    """BASELINE configs[1]: 'roots bit-exact vs CpuHal' at 2^20 cycles, SYN-280, Trace A."""
    from zktls_b200.prover import SegmentProver
    shape, po2 = circuit.SYN280, 20
    blob = circuit.syn_circuit(**shape).blob()
    io, code_m, data_m, accum_m = synth.trace_a(shape, po2, BENCH_SEED)
    gp = SegmentProver(hal, blob)
    seal_g = gp.prove(po2, io, code_m, data_m, accum_m)
    roots_g = gp.roots()
    gp.close()
    op = oracle.Prover(blob)
    op.begin(po2, io, code_m, data_m)
    seal_o = op.finish(accum_m)
    roots_o = op.roots()
    assert roots_g.shape == (7, 8)
    assert np.array_equal(roots_g, roots_o), "Merkle / FRI roots differ at the benchmark size"
    assert seal_g.size == seal_o.size == 67489
    diff = np.nonzero(seal_g != seal_o)[0]
    assert diff.size == 0, f"first differing seal word at {diff[:5]} of {seal_g.size}"


def test_hash_rows_data_group_matrix_sampled_rows(hal, oracle, torch_dev):
    """hash_rows over 224 columns x 2^22 rows (the dominant kernel's benchmark shape, 3.76 GB): 600 sampled rows -- the first and
    last CTAs, both sides of the 2^32-byte offset, and random rows -- against the oracle's sponge over the same rows."""
    torch, dev = torch_dev
    rows, cols = 1 << 22, 224
    m = device_random_fp(torch_dev, rows * cols, 0xB2001000)
    dig = hal.alloc_digest(rows)
    hal.hash_rows(dig, as_buffer(hal, m))
    hal.sync()
    rng = np.random.default_rng(22)
    sample = np.unique(np.concatenate([np.arange(0, 130), np.arange(rows - 130, rows), np.arange((1 << 21) - 20, (1 << 21) + 20), rng.integers(0, rows, 300)]))
    idx = torch.from_numpy(sample).to(dev)
    sub = to_u32(m.view(cols, rows)[:, idx].contiguous().view(-1))          # cols x len(sample), column-major like the big one
    want = oracle.hash_rows(sub, sample.size, cols).reshape(-1, 8)
    got = dig.to_numpy().reshape(rows, 8)[sample]
    assert np.array_equal(got, want)


def test_intt_and_lde_224_columns_sampled_columns(hal, oracle, torch_dev):
    """The data group's NTTs at the benchmark shape: iNTT + zk-shift of 224 x 2^20 and the x4 LDE into 224 x 2^22.  Sampled columns
    (first, last, the ones straddling the 2^26-element and 2^32-byte offsets, a few random) are compared in FULL with the oracle."""
    torch, dev = torch_dev
    po2, cols = 20, 224
    n = 1 << po2
    x = device_random_fp(torch_dev, cols * n, 0xB2001001)
    x0 = x.clone()
    buf = as_buffer(hal, x)
    hal.batch_interpolate_ntt_zk_shift(buf, cols)
    big = hal.alloc_elem(cols * 4 * n)
    hal.batch_expand_into_evaluate_ntt(big, buf, cols, 2)
    hal.sync()
    picks = sorted({0, 1, 63, 64, 127, 128, 223, 222, 17, 150, 201})      # 64 * 2^20 = 2^26 elements; 256 * 4 * 2^22 ... = 2^32 bytes at column 256 (LDE: column 64)
    coeffs = to_u32(x.view(cols, n)[picks].contiguous().view(-1)).reshape(len(picks), n)
    evals = big.to_numpy().reshape(cols, 4 * n)[picks]
    src = to_u32(x0.view(cols, n)[picks].contiguous().view(-1)).reshape(len(picks), n)
    for k, c in enumerate(picks):
        want_c = oracle.zk_shift(oracle.batch_interpolate_ntt(src[k], 1, po2), 1, po2)
        assert np.array_equal(coeffs[k], want_c), f"iNTT + zk_shift differs in column {c}"
        want_e = oracle.batch_expand_into_evaluate_ntt(want_c, 1, po2, 2)
        assert np.array_equal(evals[k], want_e), f"LDE differs in column {c}"


@pytest.mark.parametrize("po2,count", [(23, 3), (24, 2), (25, 1)])
def test_three_pass_interpolate_against_the_oracle(hal, oracle, po2, count):
    """po2 > 22 runs the three-pass tiled plan; round 1 compared it only with the repo's own level kernels."""
    x = np.random.default_rng(2300 + po2).integers(0, P, size=count << po2, dtype=np.uint32)
    b = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt_zk_shift(b, count)
    got = b.to_numpy().reshape(count, -1)
    for c in (0, count - 1):
        col = x.reshape(count, -1)[c]
        want = oracle.zk_shift(oracle.batch_interpolate_ntt(col, 1, po2), 1, po2)
        assert np.array_equal(got[c], want), f"column {c}"


@pytest.mark.parametrize("in_po2,count,eb", [(21, 3, 2), (22, 2, 2), (23, 1, 0), (24, 1, 2)])
def test_three_pass_lde_against_the_oracle(hal, oracle, in_po2, count, eb):
    """forward NTTs whose output exceeds 2^22 points (up to MAX_CYCLES_PO2 24 + 2 expand bits = 2^26), against the oracle."""
    x = np.random.default_rng(2400 + in_po2).integers(0, P, size=count << in_po2, dtype=np.uint32)
    out = hal.alloc_elem(count << (in_po2 + eb))
    hal.batch_expand_into_evaluate_ntt(out, hal.copy_from_elem(x), count, eb)
    got = out.to_numpy().reshape(count, -1)
    for c in sorted({0, count - 1}):
        want = oracle.batch_expand_into_evaluate_ntt(x.reshape(count, -1)[c], 1, in_po2, eb)
        assert np.array_equal(got[c], want), f"column {c}"


def test_merkle_root_over_2p22_leaves_matches_the_oracle(hal, oracle, torch_dev):
    """merkle_build over 2^22 leaf digests (22 levels: 11 grid launches + the single-CTA tail): root and sampled inner nodes."""
    rows = 1 << 22
    leaves = device_random_fp(torch_dev, rows * 8, 0xB2001002)
    torch, dev = torch_dev
    from zktls_b200.hal import Buffer
    nodes_t = torch.zeros(2 * rows * 8, device=dev, dtype=torch.int32)
    nodes_t[rows * 8:] = leaves
    torch.cuda.synchronize()
    nodes = Buffer(hal, nodes_t.data_ptr(), 2 * rows, 8, owner=nodes_t)
    hal.merkle_build(nodes, rows)
    got = nodes.to_numpy().reshape(2 * rows, 8)
    want = oracle.merkle_build(np.concatenate([np.zeros(rows * 8, np.uint32), to_u32(leaves)]), rows).reshape(2 * rows, 8)
    assert np.array_equal(got[1], want[1]), "root"
    assert np.array_equal(got[1:4096], want[1:4096]), "top of the tree"
    pick = np.random.default_rng(5).integers(4096, rows, 2000)
    assert np.array_equal(got[pick], want[pick])

"""SECOND, INDEPENDENT CPU oracle for the proving path -- pure Python / numpy, written from SURVEY.md App. A-D only.

Test infrastructure ("parity unpinned" against the Rust crates, like oracle/*.hpp: their source is not in /root/reference).
Purpose: the C++ oracle and the CUDA product were two restatements by one author sharing a reading
of the spec; this file is a third derivation that shares NOTHING with either -- no header, no Montgomery arithmetic, no
recursive butterflies, no replicate-expand trick:

  * the field is computed on CANONICAL values (plain `a * b % P` in uint64); Montgomery words exist only at the I/O boundary
    (digests, seal words, trace inputs), so a slip in a Montgomery constant or reduction cannot cancel out;
  * the 213 round constants come from a Grain LFSR written here from App. B.1 (not from poseidon2_consts.h);
  * NTTs are an iterative decimation-in-time DFT over all columns at once; `batch_interpolate_ntt` = DFT with w^-1, scale,
    permute; the LDE = un-permute, ZERO-PAD (not replicate), full DFT of size 4n -- the textbook definition of C.1 / C.3;
  * the sponge is absorb-then-permute over all rows at once (B.3), the rng follows B.4, the transcript D.0-D.5.

tests/golden/make_golden.py refuses to write a golden file unless this oracle and the C++ oracle agree on every array,
both seals included.  Only tests/ and tests/golden/ may import this module.
"""
import hashlib

import numpy as np

P = 2013265921
_P = np.uint64(P)
R = (1 << 32) % P
RINV = pow(1 << 32, -1, P)
INV_RATE, QUERIES, FRI_FOLD, FRI_MIN_DEGREE, CHECK_SIZE, EXT = 4, 50, 16, 256, 16, 4
NBETA = P - 11
U64 = np.uint64


# ------------------------------------------------------------------------------------------------ field (canonical values)
def enc(a):
    """canonical value(s) -> stored Montgomery word(s)"""
    return ((np.asarray(a, dtype=U64) % _P) * U64(R) % _P).astype(np.uint32)


def dec(w):
    """stored Montgomery word(s) -> canonical value(s) (uint64)"""
    return np.asarray(w, dtype=U64) * U64(RINV) % _P


def fmul(a, b):
    return np.asarray(a, dtype=U64) * np.asarray(b, dtype=U64) % _P


def fadd(a, b):
    return (np.asarray(a, dtype=U64) + np.asarray(b, dtype=U64)) % _P


def fsub(a, b):
    return (np.asarray(a, dtype=U64) + _P - np.asarray(b, dtype=U64)) % _P


def rou(po2):
    """primitive 2^po2-th root of unity: 137^(2^(27 - po2))"""
    return pow(137, 1 << (27 - po2), P)


def bitrev_perm(k):
    idx = np.arange(1 << k, dtype=np.int64)
    r = np.zeros_like(idx)
    for b in range(k):
        r |= ((idx >> b) & 1) << (k - 1 - b)
    return r


# Fp4 = Fp[x] / (x^4 + 11); arrays of shape (..., 4), canonical
def f4(a0=0, a1=0, a2=0, a3=0):
    return np.array([a0 % P, a1 % P, a2 % P, a3 % P], dtype=U64)


def f4mul(a, b):
    a = np.asarray(a, dtype=U64); b = np.asarray(b, dtype=U64)
    a0, a1, a2, a3 = (a[..., i] for i in range(4)); b0, b1, b2, b3 = (b[..., i] for i in range(4))
    nb = U64(NBETA)
    m = lambda x, y: x * y % _P
    c0 = (m(a0, b0) + m(nb, (m(a1, b3) + m(a2, b2) + m(a3, b1)) % _P)) % _P
    c1 = (m(a0, b1) + m(a1, b0) + m(nb, (m(a2, b3) + m(a3, b2)) % _P)) % _P
    c2 = (m(a0, b2) + m(a1, b1) + m(a2, b0) + m(nb, m(a3, b3))) % _P
    c3 = (m(a0, b3) + m(a1, b2) + m(a2, b1) + m(a3, b0)) % _P
    return np.stack([c0, c1, c2, c3], axis=-1)


def f4scale(a, s):
    return np.asarray(a, dtype=U64) * np.asarray(s, dtype=U64)[..., None] % _P if np.ndim(s) else np.asarray(a, dtype=U64) * U64(int(s) % P) % _P


def f4pow(a, e):
    r = f4(1); a = np.asarray(a, dtype=U64)
    while e:
        if e & 1:
            r = f4mul(r, a)
        a = f4mul(a, a); e >>= 1
    return r


def f4inv(a):
    # a^(P^4 - 2) would do; use the norm: a^-1 = conj-product / norm.  Derived here by brute linear algebra instead of the survey's
    # closed form (independence): solve the 4x4 system (multiplication-by-a matrix) * x = 1 over Fp.
    a = [int(v) for v in a]
    M = [[0] * 4 for _ in range(4)]
    for j in range(4):            # column j = a * x^j
        e = [0] * 4; e[j] = 1
        prod = [int(v) for v in f4mul(f4(*a), f4(*e))]
        for i in range(4):
            M[i][j] = prod[i]
    rhs = [1, 0, 0, 0]
    for c in range(4):            # Gauss-Jordan mod P
        piv = next(r for r in range(c, 4) if M[r][c] % P)
        M[c], M[piv] = M[piv], M[c]; rhs[c], rhs[piv] = rhs[piv], rhs[c]
        iv = pow(M[c][c], -1, P)
        M[c] = [v * iv % P for v in M[c]]; rhs[c] = rhs[c] * iv % P
        for r in range(4):
            if r != c and M[r][c]:
                f = M[r][c]
                M[r] = [(x - f * y) % P for x, y in zip(M[r], M[c])]; rhs[r] = (rhs[r] - f * rhs[c]) % P
    return f4(*rhs)


def f4powers(x, n):
    """[x^0, x^1, ..., x^(n-1)] by doubling, shape (n, 4)"""
    out = np.zeros((n, 4), dtype=U64); out[0] = f4(1)
    have, step = 1, np.asarray(x, dtype=U64)
    while have < n:
        take = min(have, n - have)
        out[have: have + take] = f4mul(out[:take], step)
        have += take; step = f4mul(step, step)
    return out


# ------------------------------------------------------------------------------------------------ Poseidon2 (App. B)
def grain_round_constants():
    """App. B.1: Poseidon Grain LFSR, field=1 sbox=0 n=31 t=24 R_F=8 R_P=21."""
    bits = []
    for value, width in ((1, 2), (0, 4), (31, 12), (24, 12), (8, 10), (21, 10)):
        bits += [(value >> (width - 1 - i)) & 1 for i in range(width)]
    bits += [1] * 30
    assert len(bits) == 80
    state = bits

    def clock():
        nonlocal state
        new = state[62] ^ state[51] ^ state[38] ^ state[23] ^ state[13] ^ state[0]
        state = state[1:] + [new]
        return new
    for _ in range(160):
        clock()

    def emit():
        while True:
            first = clock(); second = clock()
            if first:
                return second
    out = []
    while len(out) < 24 * 8 + 21:
        v = 0
        for _ in range(31):
            v = (v << 1) | emit()
        if v < P:
            out.append(v)
    return out


_RC = grain_round_constants()
assert _RC[:4] == [0x0FA20C37, 0x0795BB97, 0x12C60B9C, 0x0EABD88E]
assert hashlib.sha256(np.array(_RC, dtype="<u4").tobytes()).hexdigest() == "9f7aa102258e5f0e2fbfcb1a50b141bcb1914a08385a4c5640e1ae4dde1da983"
RC_FIRST = np.array(_RC[:96], dtype=U64).reshape(4, 24)
RC_INT = np.array(_RC[96:117], dtype=U64)
RC_LAST = np.array(_RC[117:], dtype=U64).reshape(4, 24)
DIAG = np.array([int(x, 16) for x in """
409133f0 1667a8a1 06a6c7b6 6f53160e 273b11d1 03176c5d 72f9bbf9 73ceba91
5cdef81d 01393285 46daee06 065d7ba6 52d72d6f 05dd05e0 3bab4b63 6ada3842
2fc5fbec 770d61b0 5715aae9 03ef0e90 75b6c770 242adf5f 00d0ca4c 36c0e388""".split()], dtype=U64)
# M4 as a matrix (B.2 rows), applied by a small integer matmul instead of the add chain
_M4 = np.array([[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]], dtype=U64)


def _m_ext(s):
    """s: (N, 24) canonical.  o = M4 on each chunk of 4; s[i] = o[i] + sum over chunks of o[4c + i%4]"""
    N = s.shape[0]
    o = (s.reshape(N, 6, 4) @ _M4.T) % _P          # entries < 16 * 2^31: no overflow
    sums = o.sum(axis=1) % _P
    return ((o + sums[:, None, :]) % _P).reshape(N, 24)


def _pow7(x):
    x2 = x * x % _P; x3 = x2 * x % _P; x4 = x2 * x2 % _P
    return x3 * x4 % _P


def poseidon2_mix(s):
    """s: (N, 24) canonical values -> permuted (App. B.2)"""
    s = _m_ext(np.asarray(s, dtype=U64) % _P)
    for r in range(4):
        s = _m_ext(_pow7((s + RC_FIRST[r]) % _P))
    for r in range(21):
        s = s.copy()
        s[:, 0] = _pow7((s[:, 0] + RC_INT[r]) % _P)
        tot = s.sum(axis=1) % _P
        s = (tot[:, None] + DIAG * s % _P) % _P
    for r in range(4):
        s = _m_ext(_pow7((s + RC_LAST[r]) % _P))
    return s


def unpadded_hash_many(m):
    """m: (N, L) canonical values; B.3 sponge over every row at once -> (N, 8) canonical"""
    m = np.asarray(m, dtype=U64)
    N, L = m.shape
    state = np.zeros((N, 24), dtype=U64)
    if L == 0:
        return poseidon2_mix(state)[:, :8]
    for start in range(0, L, 16):
        blk = m[:, start: start + 16]
        state[:, : blk.shape[1]] = blk          # overwrite
        state[:, blk.shape[1]: 16] = 0          # zero pad (only ever happens for the last, partial block)
        state = poseidon2_mix(state)
    return state[:, :8]


def hash_words(words):
    """hash_elem_slice over stored Montgomery words -> digest (8 stored Montgomery words)"""
    return enc(unpadded_hash_many(dec(np.asarray(words, dtype=np.uint32))[None, :])[0])


def hash_pair_many(a, b):
    """digests (N, 8) Montgomery words each -> (N, 8) Montgomery words"""
    return enc(unpadded_hash_many(np.concatenate([dec(a), dec(b)], axis=1)))


class Rng:
    """App. B.4"""

    def __init__(self):
        self.cells = np.zeros(24, dtype=U64); self.pool_used = 0

    def _mix(self):
        self.cells = poseidon2_mix(self.cells[None, :])[0]

    def mix(self, digest_words):
        if self.pool_used != 0:
            self._mix(); self.pool_used = 0
        self.cells[:8] = (self.cells[:8] + dec(digest_words)) % _P
        self._mix()

    def random_elem(self):
        if self.pool_used == 16:
            self._mix(); self.pool_used = 0
        v = int(self.cells[self.pool_used]); self.pool_used += 1
        return v

    def random_ext_elem(self):
        return f4(*[self.random_elem() for _ in range(4)])

    def random_bits(self, bits):
        val = self.random_elem()
        for _ in range(3):
            nv = self.random_elem()
            if val == 0:
                val = nv
        return val & ((1 << bits) - 1)


# ------------------------------------------------------------------------------------------------ NTT (App. C.1 - C.4)
def _powers(base, n):
    """[base^0 .. base^(n-1)] (n a power of two) by doubling"""
    pw = np.empty(n, dtype=U64); pw[0] = 1
    have, s = 1, base % P
    while have < n:
        pw[have: 2 * have] = pw[:have] * U64(s) % _P
        have *= 2; s = s * s % P
    return pw


def dft(x, w):
    """x: (count, n) canonical, natural order; returns X[j] = sum_i x[i] w^(ij), natural order.  Iterative radix-2 DIT:
    bit-reverse the input, then log2 n passes of butterflies with growing span."""
    x = np.asarray(x, dtype=U64)
    count, n = x.shape
    a = x[:, bitrev_perm(n.bit_length() - 1)].copy()
    half = 1
    while half < n:
        tw = _powers(pow(w, n // (2 * half), P), half)
        a = a.reshape(count, n // (2 * half), 2, half)
        u = a[:, :, 0, :]; v = a[:, :, 1, :] * tw % _P
        a = np.stack([(u + v) % _P, (u + _P - v) % _P], axis=2).reshape(count, n)
        half *= 2
    return a


def batch_interpolate_ntt(words, count, po2):
    """C.1: natural-order evaluations -> coefficient j at index rev(j), scaled 1/n.  Montgomery words in and out."""
    n = 1 << po2
    ev = dec(words).reshape(count, n)
    coeffs = dft(ev, pow(rou(po2), -1, P)) * U64(pow(n, -1, P)) % _P
    out = np.empty_like(coeffs); out[:, bitrev_perm(po2)] = coeffs
    return enc(out).ravel()


def zk_shift(words, count, po2):
    n = 1 << po2
    x = dec(words).reshape(count, n)
    return enc(x * _powers(3, n)[bitrev_perm(po2)] % _P).ravel()


def batch_expand_into_evaluate_ntt(words, count, in_po2, eb):
    """C.3 by its definition: bit-reversed coefficients -> natural, zero-pad to n * 2^eb, full DFT -> natural-order evaluations."""
    n = 1 << in_po2
    br = dec(words).reshape(count, n)
    nat = br[:, bitrev_perm(in_po2)]                 # natural coefficient j sits at index rev(j)
    padded = np.zeros((count, n << eb), dtype=U64); padded[:, :n] = nat
    return enc(dft(padded, rou(in_po2 + eb))).ravel()


def batch_bit_reverse(words, count, po2):
    return np.asarray(words, dtype=np.uint32).reshape(count, 1 << po2)[:, bitrev_perm(po2)].ravel().copy()


# ------------------------------------------------------------------------------------------------ hashing operators (C.5, C.6)
def hash_rows(matrix_words, rows, cols):
    m = dec(np.asarray(matrix_words, dtype=np.uint32)).reshape(cols, rows).T if cols else np.zeros((rows, 0), dtype=U64)
    return enc(unpadded_hash_many(m)).ravel()


def merkle_build(leaf_digests, rows):
    """heap-indexed nodes (2*rows, 8) Montgomery words; leaves at [rows, 2 rows), root at 1"""
    nodes = np.zeros((2 * rows, 8), dtype=np.uint32)
    nodes[rows:] = np.asarray(leaf_digests, dtype=np.uint32).reshape(rows, 8)
    size = rows
    while size > 1:
        half = size // 2
        nodes[half: size] = hash_pair_many(nodes[size: 2 * size: 2], nodes[size + 1: 2 * size: 2])
        size = half
    return nodes


def merkle_params(rows, queries=QUERIES):
    layers = rows.bit_length() - 1
    top_layer = 0
    for i in range(1, layers):          # Rust `1..layers`
        if (1 << i) <= queries:
            top_layer = i
    return layers, top_layer, 1 << top_layer


# ------------------------------------------------------------------------------------------------ small operators (C.7 - C.11, C.13)
def batch_evaluate_any(coeffs_words, poly_count, po2, which, xs_words):
    n = 1 << po2
    c = dec(coeffs_words).reshape(poly_count, n)
    xs = dec(xs_words).reshape(-1, 4)
    out = []
    for j, w in enumerate(which):
        pw = f4powers(xs[j], n)
        out.append((c[int(w)][:, None] * pw % _P).sum(axis=0) % _P)          # n < 2^33 terms of < 2^31 each: no overflow
    return enc(np.array(out, dtype=U64)).ravel()


def fri_fold(in_words, mix_words, out_count):
    m = out_count
    x = dec(in_words).reshape(4, 16 * m)
    mix = dec(mix_words)
    tot = np.zeros((m, 4), dtype=U64); cur = f4(1)
    idx = np.arange(m)
    for i in range(16):
        r = int(f"{i:04b}"[::-1], 2) * m + idx
        tot = (tot + f4mul(x[:, r].T, cur)) % _P
        cur = f4mul(cur, mix)
    return enc(tot.T).ravel()


def poly_divide(p, z):
    """p: (n, 4) canonical Fp4 coefficients (low first); synthetic division by (x - z) in place; returns remainder"""
    z = [int(v) for v in z]
    cur = [0, 0, 0, 0]
    zz = f4(*z)
    for i in range(p.shape[0] - 1, -1, -1):
        nxt = (f4mul(zz, np.array(cur, dtype=U64)) + p[i]) % _P
        p[i] = cur
        cur = nxt
    return np.array(cur, dtype=U64)


def poly_interpolate(xs, fx):
    """coefficients (low first) of the polynomial of degree < len(xs) through (xs[i], fx[i]); Fp4.  Solved as a Vandermonde system
    (not the survey's Lagrange product form) -- any correct interpolation gives the same coefficients."""
    k = len(xs)
    rows = [[f4powers(xs[i], k)[j] for j in range(k)] + [np.asarray(fx[i], dtype=U64)] for i in range(k)]
    for c in range(k):
        piv = next(r for r in range(c, k) if rows[r][c].any())
        rows[c], rows[piv] = rows[piv], rows[c]
        iv = f4inv(rows[c][c])
        rows[c] = [f4mul(v, iv) for v in rows[c]]
        for r in range(k):
            if r != c and rows[r][c].any():
                f = rows[r][c]
                rows[r] = [(x + _P - f4mul(f, y)) % _P for x, y in zip(rows[r], rows[c])]
    return [rows[i][k] for i in range(k)]


# ------------------------------------------------------------------------------------------------ circuit (C.12, D.1)
OP_CONST, OP_GET, OP_GET_GLOBAL, OP_ADD, OP_SUB, OP_MUL, OP_TRUE, OP_AND_EQZ, OP_AND_COND = range(9)
G_ACCUM, G_CODE, G_DATA = 0, 1, 2


class Circuit:
    def __init__(self, blob):
        w = [int(v) for v in np.asarray(blob, dtype=np.uint32)]
        assert w[0] == 0x5A4B4331
        self.group_size = w[1:4]; self.mix_size, self.out_size = w[4], w[5]
        n_taps, n_steps, self.ret = w[6], w[7], w[8]
        self.info = b"".join(int(v).to_bytes(4, "little") for v in w[12:16])
        p = 16
        self.taps = [tuple(w[p + 3 * i: p + 3 * i + 3]) for i in range(n_taps)]; p += 3 * n_taps
        self.steps = [tuple(w[p + 4 * i: p + 4 * i + 4]) for i in range(n_steps)]
        assert self.taps == sorted(set(self.taps))
        # registers: one per (group, column), with its list of backs; combos: distinct back-lists, numbered in sorted order
        self.regs = []
        for i, (g, c, b) in enumerate(self.taps):
            if self.regs and self.regs[-1]["group"] == g and self.regs[-1]["column"] == c:
                self.regs[-1]["backs"].append(b)
            else:
                self.regs.append({"group": g, "column": c, "pos": i, "backs": [b]})
        self.combos = sorted({tuple(r["backs"]) for r in self.regs})
        for r in self.regs:
            r["combo"] = self.combos.index(tuple(r["backs"]))

    def poly_ext(self, poly_mix, tap_value, mix_g, out_g, ext):
        """Runs the PolyExtStep program.  tap_value(i) gives tap i as an array (Fp, shape (N,)) when ext is False or as Fp4
        (shape (4,)) when ext is True.  Returns the Fp4 `tot` of the result state: (N, 4) or (4,)."""
        fp, mx = [], []
        lift = (lambda v: f4(int(v))) if ext else (lambda v: U64(int(v) % P))
        mul = f4mul if ext else fmul
        for op, a, b, c in self.steps:
            if op == OP_CONST: fp.append(lift(a))
            elif op == OP_GET: fp.append(tap_value(a))
            elif op == OP_GET_GLOBAL: fp.append(lift((mix_g if a == 0 else out_g)[b]))
            elif op == OP_ADD: fp.append(fadd(fp[a], fp[b]))
            elif op == OP_SUB: fp.append(fsub(fp[a], fp[b]))
            elif op == OP_MUL: fp.append(mul(fp[a], fp[b]))
            elif op == OP_TRUE: mx.append((f4(0), f4(1)))
            elif op == OP_AND_EQZ:
                tot, m = mx[a]
                v = fp[b]
                term = f4mul(m, v) if ext else f4scale(m, v) if np.ndim(v) == 0 else np.asarray(v, dtype=U64)[:, None] * m % _P
                mx.append(((tot + term) % _P, f4mul(m, poly_mix)))
            elif op == OP_AND_COND:
                tot, m = mx[a]; itot, im = mx[c]
                cond = fp[b]
                inner = f4mul(itot, m)
                term = f4mul(inner, cond) if ext else (np.asarray(cond, dtype=U64)[..., None] * inner % _P)
                mx.append(((tot + term) % _P, f4mul(m, im)))
            else:
                raise ValueError(op)
        return mx[self.ret][0]


def eval_check(blob, accum_w, code_w, data_w, mix_g_w, out_g_w, poly_mix_w, po2):
    """C.12 over the whole LDE domain at once; groups are (cols x 4n) Montgomery words.  Returns 4 planar columns x 4n words."""
    c = Circuit(blob)
    n = 1 << po2; dom = 4 * n
    groups = [dec(g).reshape(-1, dom) if len(g) else np.zeros((0, dom), dtype=U64) for g in (accum_w, code_w, data_w)]
    mix_g, out_g = [int(v) for v in dec(mix_g_w)], [int(v) for v in dec(out_g_w)]
    poly_mix = dec(poly_mix_w)
    cyc = np.arange(dom)

    def tap_value(i):
        g, col, back = c.taps[i]
        return groups[g][col][(cyc - 4 * back) % dom]
    tot = c.poly_ext(poly_mix, tap_value, mix_g, out_g, ext=False)
    tot = np.broadcast_to(tot, (dom, 4)) if tot.ndim == 1 else tot
    # x = w_4n^c ; y = (3x)^n takes four values (c mod 4): 3^n * w_4^(c mod 4)
    y4 = [pow(3, n, P) * pow(rou(2), r, P) % P for r in range(4)]
    inv4 = np.array([pow((y - 1) % P, -1, P) for y in y4], dtype=U64)
    ret = tot * inv4[cyc % 4][:, None] % _P
    return enc(ret.T).ravel()


# ------------------------------------------------------------------------------------------------ prover (App. D)
class WriteIOP:
    def __init__(self):
        self.proof = []; self.rng = Rng()

    def write_words(self, words):
        self.proof.extend(int(v) for v in np.asarray(words, dtype=np.uint32).ravel())

    def commit(self, digest_words):
        self.rng.mix(digest_words)


class MerkleProver:
    def __init__(self, matrix_words, rows, cols):
        self.rows, self.cols = rows, cols
        self.matrix = np.asarray(matrix_words, dtype=np.uint32).reshape(cols, rows)
        self.nodes = merkle_build(hash_rows(matrix_words, rows, cols), rows)
        self.layers, self.top_layer, self.top_size = merkle_params(rows)
        self.root = self.nodes[1]

    def commit(self, iop):
        iop.write_words(self.nodes[self.top_size: 2 * self.top_size])
        iop.commit(self.root)

    def prove(self, iop, idx):
        iop.write_words(self.matrix[:, idx])
        i = idx + self.rows
        while i >= 2 * self.top_size:
            iop.write_words(self.nodes[i ^ 1]); i >>= 1


class PolyGroup:
    """coeffs: bit-reversed coefficient words (count x n) as produced by C.1 (+ C.2)."""

    def __init__(self, coeff_words, count, po2):
        self.count, self.po2 = count, po2
        n = 1 << po2
        self.evaluated = batch_expand_into_evaluate_ntt(coeff_words, count, po2, 2)
        self.coeffs = batch_bit_reverse(coeff_words, count, po2)          # natural order from here on
        self.merkle = MerkleProver(self.evaluated, 4 * n, count)


class Prover:
    """Same interface as oracle.Prover (begin / finish / seal / roots)."""

    def __init__(self, blob):
        self.c = Circuit(blob)
        self.blob = np.asarray(blob, dtype=np.uint32)

    def _info_digest(self, info):
        return hash_words(enc(np.frombuffer(info, dtype=np.uint8).astype(U64)))

    def _commit_group(self, g, trace_words):
        cols = self.c.group_size[g]
        coeffs = zk_shift(batch_interpolate_ntt(trace_words, cols, self.po2), cols, self.po2)
        pg = PolyGroup(coeffs, cols, self.po2)
        pg.merkle.commit(self.iop)
        self._roots.append(pg.merkle.root.copy())
        self.groups[g] = pg

    def begin(self, po2, io_words, code_words, data_words):
        self.po2, self.n = po2, 1 << po2
        self.iop = WriteIOP(); self._roots = []; self.groups = [None, None, None]
        self.iop.commit(self._info_digest(b"RISC0_STARK:v1__"))
        self.iop.commit(self._info_digest(self.c.info))
        self.io = np.asarray(io_words, dtype=np.uint32)[: self.c.out_size]
        hdr = np.concatenate([self.io, np.array([po2], dtype=np.uint32)])          # po2 as a RAW word (Elem::from_u32_slice)
        self.iop.commit(hash_words(hdr)); self.iop.write_words(hdr)
        self._commit_group(G_CODE, code_words)
        self._commit_group(G_DATA, data_words)
        self.mix = enc(np.array([self.iop.rng.random_elem() for _ in range(self.c.mix_size)], dtype=U64)) if self.c.mix_size else np.zeros(0, np.uint32)
        return self.mix.copy()

    def finish(self, accum_words):
        self._commit_group(G_ACCUM, accum_words)
        self._finalize()
        return self.seal()

    def seal(self):
        return np.array(self.iop.proof, dtype=np.uint32)

    def roots(self):
        return np.array(self._roots, dtype=np.uint32).reshape(-1, 8)

    def _finalize(self):
        c, iop, n, po2 = self.c, self.iop, self.n, self.po2
        dom = 4 * n
        # 1. check polynomial
        poly_mix = iop.rng.random_ext_elem()
        check = eval_check(self.blob, self.groups[0].evaluated, self.groups[1].evaluated, self.groups[2].evaluated, self.mix, self.io, enc(poly_mix), po2)
        check = batch_interpolate_ntt(check, 4, po2 + 2)          # no zk_shift
        check_group = PolyGroup(check, CHECK_SIZE, po2)             # the same words, seen as 16 polynomials of length n
        check_group.merkle.commit(iop); self._roots.append(check_group.merkle.root.copy())
        # 2. DEEP evaluations
        z = iop.rng.random_ext_elem()
        back_one = pow(rou(po2), -1, P)
        xs = [f4scale(z, pow(back_one, b, P)) for (_, _, b) in c.taps]
        eval_u = [None] * len(c.taps)
        for g in range(3):
            idx = [i for i, t in enumerate(c.taps) if t[0] == g]
            if not idx:
                continue
            ev = dec(batch_evaluate_any(self.groups[g].coeffs, c.group_size[g], po2, [c.taps[i][1] for i in idx], enc(np.array([xs[i] for i in idx])))).reshape(-1, 4)
            for k, i in enumerate(idx):
                eval_u[i] = ev[k]
        # 3. interpolate per register, then the check evaluations at z^4
        coeff_u = []
        for r in c.regs:
            k = len(r["backs"])
            coeff_u += poly_interpolate(xs[r["pos"]: r["pos"] + k], eval_u[r["pos"]: r["pos"] + k])
        z4 = f4pow(z, 4)
        ev = dec(batch_evaluate_any(check_group.coeffs, CHECK_SIZE, po2, list(range(CHECK_SIZE)), enc(np.array([z4] * CHECK_SIZE)))).reshape(-1, 4)
        coeff_u += [ev[i] for i in range(CHECK_SIZE)]
        # 4.
        cu_words = enc(np.array(coeff_u, dtype=U64)).ravel()
        iop.write_words(cu_words); iop.commit(hash_words(cu_words))
        mix = iop.rng.random_ext_elem()
        # 5. combos
        ncomb = len(c.combos)
        combos = np.zeros((ncomb + 1, n, 4), dtype=U64)
        cur = f4(1)
        for g in range(3):
            co = dec(self.groups[g].coeffs).reshape(c.group_size[g], n)
            regs = [r for r in c.regs if r["group"] == g]
            assert [r["column"] for r in regs] == list(range(c.group_size[g])), "every column must be tapped"
            for r in regs:
                combos[r["combo"]] = (combos[r["combo"]] + co[r["column"]][:, None] * cur % _P) % _P
                cur = f4mul(cur, mix)
        co = dec(check_group.coeffs).reshape(CHECK_SIZE, n)
        for i in range(CHECK_SIZE):
            combos[ncomb] = (combos[ncomb] + co[i][:, None] * cur % _P) % _P
            cur = f4mul(cur, mix)
        # 6. subtract the interpolants, divide
        cur = f4(1)
        for r in c.regs:
            for i in range(len(r["backs"])):
                combos[r["combo"]][i] = (combos[r["combo"]][i] + _P - f4mul(cur, coeff_u[r["pos"] + i])) % _P
            cur = f4mul(cur, mix)
        for i in range(CHECK_SIZE):
            combos[ncomb][0] = (combos[ncomb][0] + _P - f4mul(cur, coeff_u[len(c.taps) + i])) % _P
            cur = f4mul(cur, mix)
        for i, backs in enumerate(c.combos):
            for b in backs:
                rem = poly_divide(combos[i], f4scale(z, pow(back_one, b, P)))
                assert not rem.any(), "DEEP quotient remainder"
        rem = poly_divide(combos[ncomb], z4)
        assert not rem.any(), "check quotient remainder"
        # 7. final polynomial -> FRI
        final = combos.sum(axis=0) % _P                                    # (n, 4), natural coefficient order
        final_words = batch_bit_reverse(enc(final.T).ravel(), 4, po2)      # 4 planar columns, bit-reversed
        self._fri(final_words, [self.groups[0], self.groups[1], self.groups[2], check_group])

    def _fri(self, coeff_words, inner_groups):
        iop = self.iop
        orig_domain = 4 * self.n
        length = self.n
        rounds = []
        while length > FRI_MIN_DEGREE:
            lpo2 = length.bit_length() - 1
            evaluated = batch_expand_into_evaluate_ntt(coeff_words, 4, lpo2, 2)
            rows = 4 * length // FRI_FOLD
            mk = MerkleProver(evaluated, rows, FRI_FOLD * EXT)
            mk.commit(iop); self._roots.append(mk.root.copy())
            fold_mix = iop.rng.random_ext_elem()
            coeff_words = fri_fold(coeff_words, enc(fold_mix), length // FRI_FOLD)
            rounds.append((4 * length, mk))
            length //= FRI_FOLD
        final = batch_bit_reverse(coeff_words, 4, length.bit_length() - 1)
        iop.write_words(final); iop.commit(hash_words(final))
        for _ in range(QUERIES):
            pos = iop.rng.random_bits(orig_domain.bit_length() - 1)
            for pg in inner_groups:
                pg.merkle.prove(iop, pos)
            for domain_r, mk in rounds:
                group = pos % (domain_r // FRI_FOLD)
                mk.prove(iop, group)
                pos = group


# ------------------------------------------------------------------------------------------------ word-level wrappers
# Same call shapes as oracle/oracle.py (Montgomery words in, Montgomery words out), so tests and make_golden.py can run
# either oracle through one code path.
def poseidon2_mix_words(cells_words):
    return enc(poseidon2_mix(dec(cells_words)[None, :])[0])


def merkle_build_words(nodes_words, rows):
    """nodes: 2*rows digests with the leaves in the upper half (C.6 layout); returns the filled array (entry 0 untouched)."""
    nodes = np.asarray(nodes_words, dtype=np.uint32).reshape(2 * rows, 8).copy()
    built = merkle_build(nodes[rows:], rows)
    nodes[1:] = built[1:]
    return nodes.ravel()


def mix_poly_coeffs(out_words, mix_start_words, mix_words, in_words, combos, input_size, count):
    out = dec(out_words).reshape(-1, count, 4)
    x = dec(in_words).reshape(input_size, count)
    cur, mix = dec(mix_start_words), dec(mix_words)
    for i in range(input_size):
        k = int(combos[i])
        out[k] = (out[k] + x[i][:, None] * cur % _P) % _P
        cur = f4mul(cur, mix)
    return enc(out).ravel()


def poly_divide_words(p_words, z_words):
    p = dec(p_words).reshape(-1, 4).copy()
    rem = poly_divide(p, dec(z_words))
    return enc(p).ravel(), enc(rem)


def eltwise_sum_extelem(in_words, count, to_add):
    x = dec(in_words).reshape(to_add, count, 4)
    return enc((x.sum(axis=0) % _P).T).ravel()


def prefix_products(io_words):
    x = dec(io_words).reshape(-1, 4)
    out = np.empty_like(x); cur = f4(1)
    for i in range(x.shape[0]):
        cur = f4mul(cur, x[i]); out[i] = cur
    return enc(out).ravel()

"""CPU oracle for the `poseidon_254` hash suite (test infrastructure; only tests/ may import it).

Poseidon over the BN254 scalar field, t = 3, x^5, R_F = 8, R_P = 57 -- circomlib's 2-input `poseidon`, the instance risc0-zkp 1.2.5
`core/hash/poseidon_254` uses for identity_p254 (un-vendored).  Plain Python integers: nothing shared with the 8-limb Montgomery code of
zktls_b200/csrc/poseidon254.cuh.  PINNED: the constants are regenerated from the Poseidon reference procedure (Grain LFSR + Cauchy
matrix) and reproduce circomlib's public known answers (`KAT` below).  The row-packing rule of hash_rows is PROVISIONAL (upstream's is not
recoverable offline) and simply restates the rule documented in zktls_b200/csrc/k_poseidon254.cu."""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N, T, RF, RP = 254, 3, 8, 57
BABYBEAR = 2013265921
KAT = {(1, 2): 7853200120776062878684798364095072458815029376092732009249414926327459813530,
       (3, 4): 14763215145315200506921711489642608356394854266165572616578112107564877678998}      # circomlib / circomlibjs test vectors


def _constants():
    bits = []
    for value, width in ((1, 2), (0, 4), (N, 12), (T, 12), (RF, 10), (RP, 10)):
        bits += [(value >> (width - 1 - i)) & 1 for i in range(width)]
    state = bits + [1] * 30

    def clock():
        nonlocal state
        new = state[62] ^ state[51] ^ state[38] ^ state[23] ^ state[13] ^ state[0]
        state = state[1:] + [new]
        return new
    for _ in range(160):
        clock()

    def word():
        v = 0
        for _ in range(N):
            while True:
                a, b = clock(), clock()
                if a:
                    break
            v = (v << 1) | b
        return v
    rc = []
    while len(rc) < T * (RF + RP):
        v = word()
        if v < P:
            rc.append(v)
    while True:
        xy = [word() % P for _ in range(2 * T)]
        if len(set(xy)) != 2 * T or any((xy[i] + xy[T + j]) % P == 0 for i in range(T) for j in range(T)):
            continue
        mds = [[pow((xy[i] + xy[T + j]) % P, -1, P) for j in range(T)] for i in range(T)]
        return rc, mds


RC, MDS = _constants()


def permute(state):
    s, r = list(state), 0
    for rnd in range(RF + RP):
        s = [(x + RC[r + i]) % P for i, x in enumerate(s)]; r += T
        if rnd < RF // 2 or rnd >= RF // 2 + RP:
            s = [pow(x, 5, P) for x in s]
        else:
            s[0] = pow(s[0], 5, P)
        s = [sum(MDS[i][j] * s[j] for j in range(T)) % P for i in range(T)]
    return s


def to_digest(x):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)


def from_digest(w):
    return sum(int(v) << (32 * i) for i, v in enumerate(np.asarray(w, dtype=np.uint32).ravel()[:8])) % P


def hash_pair(a, b):
    return to_digest(permute([0, from_digest(a), from_digest(b)])[0])


def hash_fold(nodes, input_size, output_size):
    nodes = np.asarray(nodes, dtype=np.uint32).reshape(-1, 8).copy()
    for i in range(output_size):
        nodes[output_size + i] = hash_pair(nodes[input_size + 2 * i], nodes[input_size + 2 * i + 1])
    return nodes.ravel()


def merkle_build(nodes, rows):
    nodes = np.asarray(nodes, dtype=np.uint32).copy()
    size = rows
    while size >= 2:
        nodes = hash_fold(nodes, size, size // 2); size //= 2
    return nodes


def hash_rows(matrix_words, rows, cols):
    """provisional packing (see module docstring): canonical BabyBear values, 8 per word in radix 2^31, rate 2, overwrite, zero padding"""
    rinv = pow(1 << 32, -1, BABYBEAR)
    m = (np.asarray(matrix_words, dtype=np.uint64).reshape(cols, rows) * np.uint64(rinv) % np.uint64(BABYBEAR)) if cols else np.zeros((0, rows), np.uint64)
    out = np.zeros((rows, 8), dtype=np.uint32)
    for r in range(rows):
        words = []
        for c0 in range(0, cols, 8):
            words.append(sum(int(m[c0 + k][r]) << (31 * k) for k in range(min(8, cols - c0))))
        s = [0, 0, 0]
        if not words:
            s = permute(s)
        for i in range(0, len(words), 2):
            s[1] = words[i]; s[2] = words[i + 1] if i + 1 < len(words) else 0
            s = permute(s)
        out[r] = to_digest(s[0])
    return out.ravel()

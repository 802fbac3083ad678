"""ctypes loader + thin numpy wrappers for the CPU oracle (oracle/*.hpp).

ORACLE = test infrastructure, "parity unpinned" against the Rust crates (their source is not in
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under zktls_b200/ does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

P = 2013265921
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) or f == "Makefile"]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_seal_words.restype = C.c_size_t
        _lib.orc_root_count.restype = C.c_size_t
        for name in ("orc_eval_check", "orc_accumulate", "orc_prover_new", "orc_segment_begin", "orc_segment_finish", "orc_verify_segment"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_void_p
        _lib.orc_rng_new.restype = C.c_void_p
        for name in ("orc_fp_encode", "orc_fp_decode", "orc_fp_mul", "orc_fp_add", "orc_fp_sub", "orc_fp_inv", "orc_rng_random_elem", "orc_rng_random_bits"):
            getattr(_lib, name).restype = C.c_uint32
    return _lib


def _p(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


def _sz(x):
    return C.c_size_t(int(x))


def _check(err):
    if err:
        msg = C.cast(err, C.c_char_p).value.decode()
        C.CDLL(None).free(C.c_void_p(err))
        raise RuntimeError("oracle: " + msg)


def u32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


# ---- field ------------------------------------------------------------------------------------------
def encode(x):
    """canonical value(s) -> Montgomery word(s) (vectorised in numpy, checked against orc_fp_encode)."""
    x = np.asarray(x, dtype=np.uint64) % P
    return ((x << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def decode(w):
    w = np.asarray(w, dtype=np.uint64)
    rinv = pow(1 << 32, -1, P)
    return ((w * np.uint64(rinv % P)) % np.uint64(P)).astype(np.uint32) if rinv < (1 << 31) else _decode_slow(w)


def _decode_slow(w):
    rinv = pow(1 << 32, -1, P)
    return np.array([(int(v) * rinv) % P for v in np.asarray(w).ravel()], dtype=np.uint32).reshape(np.shape(w))


def random_fp(rng, shape):
    """uniform canonical Montgomery words < P"""
    return rng.integers(0, P, size=shape, dtype=np.uint32)


def fp4_mul(a, b):
    out = np.zeros(4, np.uint32); lib().orc_fp4_mul(_p(out), _p(u32(a)), _p(u32(b))); return out


def fp4_inv(a):
    out = np.zeros(4, np.uint32); lib().orc_fp4_inv(_p(out), _p(u32(a))); return out


def rou_tables():
    f = np.zeros(28, np.uint32); r = np.zeros(28, np.uint32); lib().orc_rou_tables(_p(f), _p(r)); return f, r


def poseidon2_round_constants():
    c = np.zeros(213, np.uint32); lib().orc_poseidon2_round_constants(_p(c)); return c


def poseidon2_mix(cells):
    c = u32(cells).copy(); lib().orc_poseidon2_mix(_p(c)); return c


def hash_elem_slice(elems):
    e = u32(elems); d = np.zeros(8, np.uint32); lib().orc_hash_elem_slice(_p(d), _p(e) if e.size else None, _sz(e.size)); return d


def hash_pair(a, b):
    d = np.zeros(8, np.uint32); lib().orc_hash_pair(_p(d), _p(u32(a)), _p(u32(b))); return d


class Rng:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_rng_new())

    def __del__(self):
        lib().orc_rng_free(self.h)

    def mix(self, digest):
        lib().orc_rng_mix(self.h, _p(u32(digest)))

    def random_elem(self):
        return int(lib().orc_rng_random_elem(self.h))

    def random_bits(self, bits):
        return int(lib().orc_rng_random_bits(self.h, C.c_int(bits)))


# ---- Hal operators (all take/return numpy uint32 arrays of Montgomery words) --------------------------
def batch_interpolate_ntt(io, count, po2):
    io = u32(io).copy(); lib().orc_batch_interpolate_ntt(_p(io), _sz(count), C.c_int(po2)); return io


def zk_shift(io, count, po2):
    io = u32(io).copy(); lib().orc_zk_shift(_p(io), _sz(count), C.c_int(po2)); return io


def batch_expand(inp, count, in_po2, expand_bits):
    inp = u32(inp); out = np.zeros(inp.size << expand_bits, np.uint32)
    lib().orc_batch_expand(_p(out), _p(inp), _sz(count), C.c_int(in_po2), C.c_int(expand_bits)); return out


def batch_evaluate_ntt(io, count, po2, expand_bits):
    io = u32(io).copy(); lib().orc_batch_evaluate_ntt(_p(io), _sz(count), C.c_int(po2), C.c_int(expand_bits)); return io


def batch_expand_into_evaluate_ntt(inp, count, in_po2, expand_bits):
    inp = u32(inp); out = np.zeros(inp.size << expand_bits, np.uint32)
    lib().orc_batch_expand_into_evaluate_ntt(_p(out), _p(inp), _sz(count), C.c_int(in_po2), C.c_int(expand_bits)); return out


def batch_bit_reverse(io, count, po2):
    io = u32(io).copy(); lib().orc_batch_bit_reverse(_p(io), _sz(count), C.c_int(po2)); return io


def hash_rows(matrix, rows, cols):
    m = u32(matrix); out = np.zeros(rows * 8, np.uint32); lib().orc_hash_rows(_p(out), _p(m), _sz(rows), _sz(cols)); return out


def hash_fold(io, input_size, output_size):
    io = u32(io).copy(); lib().orc_hash_fold(_p(io), _sz(input_size), _sz(output_size)); return io


def merkle_build(nodes, rows):
    nodes = u32(nodes).copy(); lib().orc_merkle_build(_p(nodes), _sz(rows)); return nodes


def batch_evaluate_any(coeffs, poly_count, po2, which, xs):
    which = u32(which); xs = u32(xs); out = np.zeros(which.size * 4, np.uint32)
    lib().orc_batch_evaluate_any(_p(u32(coeffs)), _sz(poly_count), C.c_int(po2), _p(which), _p(xs), _p(out), _sz(which.size)); return out


def mix_poly_coeffs(out, mix_start, mix, inp, combos, input_size, count):
    out = u32(out).copy()
    lib().orc_mix_poly_coeffs(_p(out), _p(u32(mix_start)), _p(u32(mix)), _p(u32(inp)), _p(u32(combos)), _sz(input_size), _sz(count)); return out


def eltwise_sum_extelem(inp, count, to_add):
    out = np.zeros(4 * count, np.uint32); lib().orc_eltwise_sum_extelem(_p(out), _p(u32(inp)), _sz(count), _sz(to_add)); return out


def fri_fold(inp, mix, out_count):
    out = np.zeros(4 * out_count, np.uint32); lib().orc_fri_fold(_p(out), _p(u32(inp)), _p(u32(mix)), _sz(out_count)); return out


def eltwise_add_elem(a, b):
    a = u32(a); out = np.zeros_like(a); lib().orc_eltwise_add_elem(_p(out), _p(a), _p(u32(b)), _sz(a.size)); return out


def eltwise_zeroize_elem(x):
    x = u32(x).copy(); lib().orc_eltwise_zeroize_elem(_p(x), _sz(x.size)); return x


def gather_sample(src, idx, size, stride):
    out = np.zeros(size, np.uint32); lib().orc_gather_sample(_p(out), _p(u32(src)), _sz(idx), _sz(size), _sz(stride)); return out


def scatter(into, index, offsets, values):
    into = u32(into).copy(); index = u32(index)
    lib().orc_scatter(_p(into), _p(index), _sz(index.size - 1), _p(u32(offsets)), _p(u32(values))); return into


def prefix_products(io):
    io = u32(io).copy(); lib().orc_prefix_products(_p(io), _sz(io.size // 4)); return io


def poly_divide(p, z):
    p = u32(p).copy(); rem = np.zeros(4, np.uint32); lib().orc_poly_divide(_p(p), _sz(p.size // 4), _p(u32(z)), _p(rem)); return p, rem


def eval_check(blob, accum, code, data, mix_g, out_g, poly_mix, po2):
    blob = u32(blob); check = np.zeros(16 << po2, np.uint32)
    _check(lib().orc_eval_check(_p(check), _p(blob), _sz(blob.size), _p(u32(accum)), _p(u32(code)), _p(u32(data)), _p(u32(mix_g)), _p(u32(out_g)), _p(u32(poly_mix)), C.c_int(po2)))
    return check


def accumulate(blob, accum, code, data, mix_g, out_g, po2):
    """CircuitHal::accumulate: runs the circuit blob's witness program over `accum` (returns the new accum group)."""
    blob = u32(blob); accum = u32(accum).copy()
    _check(lib().orc_accumulate(_p(blob), _sz(blob.size), _p(accum), _p(u32(code)), _p(u32(data)), _p(u32(mix_g)), _p(u32(out_g)), C.c_int(po2)))
    return accum


class Prover:
    """Two-phase segment prover: begin() commits code+data and returns the `mix`
    globals; finish(accum) commits accum, finalizes, and returns the seal words."""

    def __init__(self, blob):
        self.blob = u32(blob); self.h = C.c_void_p()
        _check(lib().orc_prover_new(_p(self.blob), _sz(self.blob.size), C.byref(self.h)))
        self.mix_size = int(self.blob[4])

    def __del__(self):
        if self.h:
            lib().orc_prover_free(self.h)

    def begin(self, po2, io, code, data):
        mix = np.zeros(self.mix_size, np.uint32)
        _check(lib().orc_segment_begin(self.h, C.c_int(po2), _p(u32(io)), _p(u32(code)), _p(u32(data)), _p(mix)))
        return mix

    def finish(self, accum):
        _check(lib().orc_segment_finish(self.h, _p(u32(accum))))
        return self.seal()

    def seal(self):
        n = lib().orc_seal_words(self.h); s = np.zeros(n, np.uint32); lib().orc_seal_copy(self.h, _p(s)); return s

    def roots(self):
        n = lib().orc_root_count(self.h); r = np.zeros(n * 8, np.uint32); lib().orc_roots_copy(self.h, _p(r)); return r.reshape(n, 8)

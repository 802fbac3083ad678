"""Host-side mirror of the risc0_zkp `Hal` operator surface over the libzkb200 C-ABI.

Method names and argument meaning follow risc0-zkp 1.2.5 `hal::Hal` (risc0-zkp/src/hal/mod.rs; reached from the
reference at /root/reference/crates/guest-prover-r0/src/prover.rs:90): buffers are flat arrays of field elements,
a batch of `count` polynomials is `count` consecutive columns, and operators return nothing and raise on failure
(the Rust trait panics).  Buffers are device-resident; `Buffer.to_numpy()` is `Buffer::to_vec`, `slice` is
`Buffer::slice`.  This is the object the parity tests drive; a Rust `B200Hal` would bind the same C functions
(see INTEGRATION.md).
"""
import ctypes as C
import numpy as np

from ._lib import lib, check, ZkbError  # noqa: F401

P = 2013265921
u32p = C.POINTER(C.c_uint32)


def _sz(x):
    return C.c_size_t(int(x))


def _hp(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


class Buffer:
    """Device buffer of `size` elements of `words` u32 each (Fp: 1, Fp4: 4, Digest: 8)."""

    def __init__(self, hal, ptr, size, words, owner=None):
        self.hal, self.ptr, self.size, self.words, self._owner = hal, ptr, int(size), words, owner

    @property
    def nbytes(self):
        return self.size * self.words * 4

    def slice(self, offset, size):
        assert 0 <= offset and offset + size <= self.size
        return Buffer(self.hal, self.ptr + offset * self.words * 4, size, self.words, owner=self._owner or self)

    def to_numpy(self):
        out = np.empty(self.size * self.words, np.uint32)
        check(lib().zkb_d2h(self.hal.ctx, _hp(out), C.c_void_p(self.ptr), _sz(self.nbytes)))
        return out

    def copy_from(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        assert arr.size == self.size * self.words
        check(lib().zkb_h2d(self.hal.ctx, C.c_void_p(self.ptr), _hp(arr), _sz(self.nbytes)))
        self.hal.sync()

    def __del__(self):
        if self._owner is None and self.ptr and self.hal.ctx:
            try:
                lib().zkb_free(self.hal.ctx, C.c_void_p(self.ptr))
            except Exception:
                pass


class B200Hal:
    def __init__(self, device=0, stream=None):
        self.ctx = C.c_void_p()
        if stream is None:
            check(lib().zkb_init(C.c_int(device), C.byref(self.ctx)))
        else:
            check(lib().zkb_init_on_stream(C.c_int(device), C.c_void_p(int(stream)), C.byref(self.ctx)))
        self.device = device

    def close(self):
        if self.ctx:
            check(lib().zkb_destroy(self.ctx))
            self.ctx = C.c_void_p()

    # ---- memory -------------------------------------------------------------------------------------------
    def _alloc(self, size, words):
        p = C.c_void_p()
        check(lib().zkb_alloc(self.ctx, _sz(max(size * words * 4, 16)), C.byref(p)))
        return Buffer(self, p.value, size, words)

    def alloc_elem(self, size):
        b = self._alloc(size, 1); check(lib().zkb_memset0(self.ctx, C.c_void_p(b.ptr), _sz(b.nbytes))); return b

    def alloc_extelem(self, size):
        b = self._alloc(size, 4); check(lib().zkb_memset0(self.ctx, C.c_void_p(b.ptr), _sz(b.nbytes))); return b

    def alloc_digest(self, size):
        b = self._alloc(size, 8); check(lib().zkb_memset0(self.ctx, C.c_void_p(b.ptr), _sz(b.nbytes))); return b

    def alloc_u32(self, size):
        return self._alloc(size, 1)

    def copy_from_elem(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint32); b = self._alloc(arr.size, 1); b.copy_from(arr); return b

    def copy_from_extelem(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint32); b = self._alloc(arr.size // 4, 4); b.copy_from(arr); return b

    def copy_from_digest(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint32); b = self._alloc(arr.size // 8, 8); b.copy_from(arr); return b

    copy_from_u32 = copy_from_elem

    def sync(self):
        check(lib().zkb_sync(self.ctx))

    def kernel_launches(self):
        n = C.c_uint64(); check(lib().zkb_kernel_launches(self.ctx, C.byref(n))); return n.value

    def timer_start(self):
        check(lib().zkb_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float(); check(lib().zkb_timer_stop(self.ctx, C.byref(ms))); return ms.value

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(lib().zkb_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    # ---- NTT family -----------------------------------------------------------------------------------------
    @staticmethod
    def _po2(n):
        k = int(n).bit_length() - 1
        if n <= 0 or (1 << k) != n:
            raise ZkbError("size is not a power of two")
        return k

    def batch_interpolate_ntt(self, io, count):
        check(lib().zkb_batch_interpolate_ntt(self.ctx, C.c_void_p(io.ptr), _sz(count), C.c_int(self._po2(io.size // count))))

    def zk_shift(self, io, count):
        check(lib().zkb_zk_shift(self.ctx, C.c_void_p(io.ptr), _sz(count), C.c_int(self._po2(io.size // count))))

    def batch_interpolate_ntt_zk_shift(self, io, count):
        check(lib().zkb_batch_interpolate_ntt_zk_shift(self.ctx, C.c_void_p(io.ptr), _sz(count), C.c_int(self._po2(io.size // count))))

    def batch_expand(self, out, inp, count):
        in_po2 = self._po2(inp.size // count); eb = self._po2(out.size // count) - in_po2
        check(lib().zkb_batch_expand(self.ctx, C.c_void_p(out.ptr), C.c_void_p(inp.ptr), _sz(count), C.c_int(in_po2), C.c_int(eb)))

    def batch_evaluate_ntt(self, io, count, expand_bits):
        check(lib().zkb_batch_evaluate_ntt(self.ctx, C.c_void_p(io.ptr), _sz(count), C.c_int(self._po2(io.size // count)), C.c_int(expand_bits)))

    def batch_expand_into_evaluate_ntt(self, out, inp, count, expand_bits):
        check(lib().zkb_batch_expand_into_evaluate_ntt(self.ctx, C.c_void_p(out.ptr), C.c_void_p(inp.ptr), _sz(count),
                                                       C.c_int(self._po2(inp.size // count)), C.c_int(expand_bits)))

    def batch_bit_reverse(self, io, count):
        check(lib().zkb_batch_bit_reverse(self.ctx, C.c_void_p(io.ptr), _sz(count), C.c_int(self._po2(io.size // count))))

    # ---- hashing ---------------------------------------------------------------------------------------------
    def hash_rows(self, out, matrix):
        rows = out.size; cols = matrix.size // rows if rows else 0
        check(lib().zkb_poseidon2_hash_rows(self.ctx, C.c_void_p(out.ptr), C.c_void_p(matrix.ptr), _sz(rows), _sz(cols)))

    def hash_fold(self, io, input_size, output_size):
        check(lib().zkb_poseidon2_hash_fold(self.ctx, C.c_void_p(io.ptr), _sz(input_size), _sz(output_size)))

    def merkle_build(self, nodes, rows):
        check(lib().zkb_poseidon2_merkle_build(self.ctx, C.c_void_p(nodes.ptr), _sz(rows)))

    # ---- poseidon_254 suite (identity_p254's hash; first slice, see include/zkb200.h) ----------------------------------------
    def p254_hash_rows(self, out, matrix):
        rows = out.size; cols = matrix.size // rows if rows else 0
        check(lib().zkb_poseidon254_hash_rows(self.ctx, C.c_void_p(out.ptr), C.c_void_p(matrix.ptr), _sz(rows), _sz(cols)))

    def p254_hash_fold(self, io, input_size, output_size):
        check(lib().zkb_poseidon254_hash_fold(self.ctx, C.c_void_p(io.ptr), _sz(input_size), _sz(output_size)))

    def p254_merkle_build(self, nodes, rows):
        check(lib().zkb_poseidon254_merkle_build(self.ctx, C.c_void_p(nodes.ptr), _sz(rows)))

    # ---- mixing / DEEP / FRI -----------------------------------------------------------------------------------
    def batch_evaluate_any(self, coeffs, poly_count, which, xs, out):
        po2 = self._po2(coeffs.size // poly_count)
        check(lib().zkb_batch_evaluate_any(self.ctx, C.c_void_p(coeffs.ptr), _sz(poly_count), C.c_int(po2), C.c_void_p(which.ptr),
                                           C.c_void_p(xs.ptr), C.c_void_p(out.ptr), _sz(which.size)))

    def mix_poly_coeffs(self, out, mix_start, mix, inp, combos, input_size, count):
        ms = np.ascontiguousarray(mix_start, np.uint32); m = np.ascontiguousarray(mix, np.uint32)
        check(lib().zkb_mix_poly_coeffs(self.ctx, C.c_void_p(out.ptr), _hp(ms), _hp(m), C.c_void_p(inp.ptr), C.c_void_p(combos.ptr),
                                        _sz(input_size), _sz(count)))

    def poly_divide(self, poly, z):
        """Device synthetic division by (x - z); returns the remainder (4 words)."""
        rem = self.alloc_extelem(1); zz = np.ascontiguousarray(z, np.uint32)
        check(lib().zkb_poly_divide(self.ctx, C.c_void_p(poly.ptr), _sz(poly.size), _hp(zz), C.c_void_p(rem.ptr)))
        return rem.to_numpy()

    def combos_divide(self, combos, n, combo, points):
        """Prover::finalize's division step in one call: polynomial combo[k] of `combos` (n Fp4 coefficients each) is divided in place
        by (x - points[k]); returns the remainders, one row of 4 words per division."""
        combo = np.ascontiguousarray(combo, np.uint32); pts = np.ascontiguousarray(points, np.uint32).reshape(-1, 4)
        rem = np.zeros((combo.size, 4), np.uint32)
        check(lib().zkb_combos_divide(self.ctx, C.c_void_p(combos.ptr), _sz(n), _sz(combos.size // n), _hp(combo), _hp(pts), _sz(combo.size), _hp(rem)))
        return rem

    def eltwise_sum_extelem(self, out, inp):
        count = out.size // 4
        check(lib().zkb_eltwise_sum_extelem(self.ctx, C.c_void_p(out.ptr), C.c_void_p(inp.ptr), _sz(count), _sz(inp.size // count)))

    def fri_fold(self, out, inp, mix):
        m = np.ascontiguousarray(mix, np.uint32)
        check(lib().zkb_fri_fold(self.ctx, C.c_void_p(out.ptr), C.c_void_p(inp.ptr), _hp(m), _sz(out.size // 4)))

    def eltwise_add_elem(self, out, a, b):
        check(lib().zkb_eltwise_add_elem(self.ctx, C.c_void_p(out.ptr), C.c_void_p(a.ptr), C.c_void_p(b.ptr), _sz(out.size)))

    def eltwise_copy_elem(self, out, inp):
        check(lib().zkb_eltwise_copy_elem(self.ctx, C.c_void_p(out.ptr), C.c_void_p(inp.ptr), _sz(out.size)))

    def eltwise_zeroize_elem(self, io):
        check(lib().zkb_eltwise_zeroize_elem(self.ctx, C.c_void_p(io.ptr), _sz(io.size)))

    def gather_sample(self, dst, src, idx, size, stride):
        check(lib().zkb_gather_sample(self.ctx, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _sz(idx), _sz(size), _sz(stride)))

    def gather_rows(self, dst, src, idx, size, stride):
        """dst[q * size + i] = src[idx[q] + i * stride]: every query's row of one tree in one launch (batched gather_sample)"""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        check(lib().zkb_gather_rows(self.ctx, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _sz(src.size), idx.ctypes.data_as(C.POINTER(C.c_uint32)), _sz(idx.size), _sz(size), _sz(stride)))

    def prefix_products(self, io):
        check(lib().zkb_prefix_products(self.ctx, C.c_void_p(io.ptr), _sz(io.size)))

    def scatter(self, into, index, offsets, values):
        ix = np.ascontiguousarray(index, np.uint32); of = np.ascontiguousarray(offsets, np.uint32); va = np.ascontiguousarray(values, np.uint32)
        check(lib().zkb_scatter(self.ctx, C.c_void_p(into.ptr), _sz(into.size), _hp(ix), _sz(max(ix.size - 1, 0)), _hp(of), _hp(va)))

    # ---- CircuitHal ---------------------------------------------------------------------------------------------
    def accumulate(self, circuit_blob, accum, code, data, mix_g, out_g, po2):
        """CircuitHal::accumulate: fills `accum` (device, column-major) from the code / data traces and the mix / io globals."""
        blob = np.ascontiguousarray(circuit_blob, np.uint32)
        mg = np.ascontiguousarray(mix_g, np.uint32); og = np.ascontiguousarray(out_g, np.uint32)
        check(lib().zkb_accumulate(self.ctx, _hp(blob), _sz(blob.size), C.c_void_p(accum.ptr), C.c_void_p(code.ptr), C.c_void_p(data.ptr),
                                   _hp(mg), _hp(og), C.c_int(po2)))

    def eval_check(self, check_buf, circuit_blob, accum, code, data, mix_g, out_g, poly_mix, po2):
        blob = np.ascontiguousarray(circuit_blob, np.uint32)
        mg = np.ascontiguousarray(mix_g, np.uint32); og = np.ascontiguousarray(out_g, np.uint32); pm = np.ascontiguousarray(poly_mix, np.uint32)
        check(lib().zkb_eval_check(self.ctx, C.c_void_p(check_buf.ptr), _hp(blob), _sz(blob.size), C.c_void_p(accum.ptr), C.c_void_p(code.ptr),
                                   C.c_void_p(data.ptr), _hp(mg), _hp(og), _hp(pm), C.c_int(po2)))

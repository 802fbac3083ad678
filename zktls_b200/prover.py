"""Host-side mirror of the segment prover over the libzkb200 C-ABI.

`SegmentProver` corresponds to risc0-zkp `prove::Prover` driven by the circuit crate's `prove_segment` (SURVEY.md
App. D.2): begin() commits `code` and `data` and returns the Fiat-Shamir `mix` globals the caller's accumulate step
needs; finish() commits `accum` and finalizes; prove() is the one-shot form.  Traces may be numpy arrays (host) or
`hal.Buffer`s (device).
"""
import ctypes as C
import numpy as np

from ._lib import lib, check
from .hal import B200Hal, Buffer, _hp, _sz


def _trace_arg(t):
    if isinstance(t, Buffer):
        return C.c_void_p(t.ptr), True
    arr = np.ascontiguousarray(t, dtype=np.uint32)
    return arr, False


class SegmentProver:
    def __init__(self, hal: B200Hal, circuit_blob):
        self.hal = hal
        self.blob = np.ascontiguousarray(circuit_blob, dtype=np.uint32)
        self.mix_size = int(self.blob[4])
        self.h = C.c_void_p()
        check(lib().zkb_prover_new(hal.ctx, _hp(self.blob), _sz(self.blob.size), C.byref(self.h)))

    def close(self):
        if self.h:
            check(lib().zkb_prover_free(self.h)); self.h = C.c_void_p()

    @staticmethod
    def _ptr(a, dev):
        return a if dev else a.ctypes.data_as(C.c_void_p)

    def begin(self, po2, io, code, data):
        io = np.ascontiguousarray(io, dtype=np.uint32)
        (c, cd), (d, dd) = _trace_arg(code), _trace_arg(data)
        assert cd == dd, "code and data must both be host or both be device"
        mix = np.zeros(max(self.mix_size, 1), np.uint32)
        check(lib().zkb_prover_segment_begin(self.h, C.c_int(po2), _hp(io), self._ptr(c, cd), self._ptr(d, dd), C.c_int(int(cd)), _hp(mix)))
        return mix[: self.mix_size]

    def finish(self, accum):
        a, ad = _trace_arg(accum)
        check(lib().zkb_prover_segment_finish(self.h, self._ptr(a, ad), C.c_int(int(ad))))
        return self.seal()

    def prove(self, po2, io, code, data, accum):
        io = np.ascontiguousarray(io, dtype=np.uint32)
        (c, cd), (d, dd), (a, ad) = _trace_arg(code), _trace_arg(data), _trace_arg(accum)
        assert cd == dd == ad
        check(lib().zkb_prove_segment(self.h, C.c_int(po2), _hp(io), self._ptr(c, cd), self._ptr(d, dd), self._ptr(a, ad), C.c_int(int(cd))))
        return self.seal()

    def stage(self, po2, code, data, accum=None):
        """Starts the asynchronous upload of one segment's host traces (numpy arrays, ideally over pinned memory) into a
        staging slot; returns immediately.  The arrays must stay alive and unmodified until the matching prove_staged().
        accum=None: the accum group is computed on the device by the circuit's witness program (CircuitHal::accumulate) inside
        prove_staged and never uploaded."""
        arrs = [None if t is None else np.ascontiguousarray(t, dtype=np.uint32) for t in (code, data, accum)]
        self._staged = getattr(self, "_staged", []) + [arrs]
        check(lib().zkb_prover_stage_traces(self.h, C.c_int(po2), *[C.c_void_p(None) if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]))

    def stage_wait(self):
        check(lib().zkb_prover_stage_wait(self.h))

    def prove_staged(self, io):
        io = np.ascontiguousarray(io, dtype=np.uint32)
        check(lib().zkb_prove_staged(self.h, _hp(io)))
        self._staged.pop(0)
        return self.seal()

    def seal(self):
        n = C.c_size_t(); check(lib().zkb_prover_seal_words(self.h, C.byref(n)))
        s = np.zeros(n.value, np.uint32); check(lib().zkb_prover_seal_copy(self.h, _hp(s))); return s

    def roots(self):
        n = C.c_size_t(); check(lib().zkb_prover_root_count(self.h, C.byref(n)))
        r = np.zeros(n.value * 8, np.uint32); check(lib().zkb_prover_roots_copy(self.h, _hp(r))); return r.reshape(n.value, 8)


def control_id(po2, code_root):
    """One control-ID entry (po2, code root[8]) for verify_segment: the Merkle root of the code group identifies the program
    (risc0-zkp `check_code(po2, root)`); `code_root` is `roots()[0]` of a prover that committed the genuine code trace."""
    return np.concatenate([np.array([po2], np.uint32), np.ascontiguousarray(code_root, dtype=np.uint32).ravel()[:8]])


def verify_segment(circuit_blob, seal, control_ids):
    """Raises ZkbError unless `seal` verifies AND its code root is one of `control_ids` (array of 9-word entries, see
    control_id()).  Returns (po2, code_root)."""
    blob = np.ascontiguousarray(circuit_blob, dtype=np.uint32); seal = np.ascontiguousarray(seal, dtype=np.uint32)
    ids = np.ascontiguousarray(control_ids, dtype=np.uint32).ravel()
    if ids.size == 0 or ids.size % 9:
        raise ValueError("control_ids must hold 9-word entries (po2, code root[8])")
    out = np.zeros(9, np.uint32)
    check(lib().zkb_verify_segment(_hp(blob), _sz(blob.size), _hp(seal), _sz(seal.size), _hp(ids), _sz(ids.size // 9), _hp(out)))
    return int(out[0]), out[1:].copy()


def seal_code_root(circuit_blob, seal):
    """(po2, code root) a seal commits to, with every other check of verify_segment applied -- for building a control-ID table
    from a trusted seal; NOT a verification of the program."""
    blob = np.ascontiguousarray(circuit_blob, dtype=np.uint32); seal = np.ascontiguousarray(seal, dtype=np.uint32)
    out = np.zeros(9, np.uint32)
    check(lib().zkb_verify_segment(_hp(blob), _sz(blob.size), _hp(seal), _sz(seal.size), None, _sz(0), _hp(out)))
    return int(out[0]), out[1:].copy()

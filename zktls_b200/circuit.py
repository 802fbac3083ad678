"""Circuit blobs: TapSet + PolyExtStep program as data (layout documented in DESIGN.md, "circuit blob").

Mirrors risc0-zkp `taps::TapSet` / `adapter::PolyExtStep`.  The rv32im circuit's generated
`poly_fp` is not available offline, so the benchmark circuit is the synthetic SYN family defined here
(SYN-280 = accum 40 / code 16 / data 224 columns).
"""
import numpy as np

MAGIC = 0x5A4B4331
OP_CONST, OP_GET, OP_GET_GLOBAL, OP_ADD, OP_SUB, OP_MUL, OP_TRUE, OP_AND_EQZ, OP_AND_COND = range(9)
GROUP_ACCUM, GROUP_CODE, GROUP_DATA = 0, 1, 2
GLOBAL_MIX, GLOBAL_OUT = 0, 1


class CircuitBuilder:
    def __init__(self, accum_cols, code_cols, data_cols, mix_size, out_size, info=b"SYN:v1__________"):
        assert len(info) == 16
        self.group_size = [accum_cols, code_cols, data_cols]
        self.mix_size, self.out_size, self.info = mix_size, out_size, info
        self.taps, self.steps = [], []
        self.n_fp, self.n_mix = 0, 0
        self.ret = None

    def add_tap(self, group, column, back):
        self.taps.append((group, column, back))

    def finish_taps(self):
        self.taps = sorted(set(self.taps))
        self.tap_index = {t: i for i, t in enumerate(self.taps)}

    def _fp(self, op, a=0, b=0, c=0):
        self.steps.append((op, a, b, c)); self.n_fp += 1; return self.n_fp - 1

    def _mix(self, op, a=0, b=0, c=0):
        self.steps.append((op, a, b, c)); self.n_mix += 1; return self.n_mix - 1

    def const(self, v): return self._fp(OP_CONST, v % 2013265921)
    def get(self, group, column, back): return self._fp(OP_GET, self.tap_index[(group, column, back)])
    def get_global(self, base, offset): return self._fp(OP_GET_GLOBAL, base, offset)
    def add(self, a, b): return self._fp(OP_ADD, a, b)
    def sub(self, a, b): return self._fp(OP_SUB, a, b)
    def mul(self, a, b): return self._fp(OP_MUL, a, b)
    def true(self): return self._mix(OP_TRUE)
    def and_eqz(self, x, val): return self._mix(OP_AND_EQZ, x, val)
    def and_cond(self, x, cond, inner): return self._mix(OP_AND_COND, x, cond, inner)

    def blob(self):
        assert self.ret is not None
        hdr = [MAGIC] + self.group_size + [self.mix_size, self.out_size, len(self.taps), len(self.steps), self.ret, self.n_fp, self.n_mix, 0]
        hdr += [int.from_bytes(self.info[4 * i: 4 * i + 4], "little") for i in range(4)]
        words = hdr + [w for t in self.taps for w in t] + [w for s in self.steps for w in s]
        return np.array(words, dtype=np.uint32)


ZK_ROWS = 64   # trailing rows left unconstrained (the reference fills ZK_CYCLES rows with noise)


def syn_circuit(accum_cols=40, code_cols=16, data_cols=224, mix_size=20, out_size=32):
    """SYN family: selector-gated quadratic recurrences.

    code[0] = sel; code[1..] = per-row constants k.  With `m = mix` and `o = out` globals:
      sel * (sel - 1)                                                  == 0
      sel * (d_j[i] - d_j[i-1]^2 - k_{1 + j mod (C-1)}[i])             == 0     every data column j
      sel * (a_j[i] - m_{j mod M} * d_{j mod D}[i] - d_{(j+1) mod D}[i-1])      == 0     even accum column j
      sel * (a_j[i] - a_{j-1}[i-1] * a_{j-1}[i] - m_{j mod M} - o_{j mod O})    == 0     odd accum column j
    Every accum/data column is tapped at back 0 and 1, every code column at back 0 => combos {0} and {0,1}.
    """
    assert code_cols >= 2 and data_cols >= 1 and accum_cols >= 1
    b = CircuitBuilder(accum_cols, code_cols, data_cols, mix_size, out_size, info=("SYN%d:v1" % (accum_cols + code_cols + data_cols)).ljust(16, "_").encode())
    for c in range(accum_cols):
        b.add_tap(GROUP_ACCUM, c, 0); b.add_tap(GROUP_ACCUM, c, 1)
    for c in range(code_cols):
        b.add_tap(GROUP_CODE, c, 0)
    for c in range(data_cols):
        b.add_tap(GROUP_DATA, c, 0); b.add_tap(GROUP_DATA, c, 1)
    b.finish_taps()
    sel = b.get(GROUP_CODE, 0, 0)
    one = b.const(1)
    top = b.and_eqz(b.true(), b.mul(sel, b.sub(sel, one)))
    inner = b.true()
    for j in range(data_cols):
        cur, prev = b.get(GROUP_DATA, j, 0), b.get(GROUP_DATA, j, 1)
        k = b.get(GROUP_CODE, 1 + j % (code_cols - 1), 0)
        inner = b.and_eqz(inner, b.sub(b.sub(cur, b.mul(prev, prev)), k))
    for j in range(accum_cols):
        a = b.get(GROUP_ACCUM, j, 0)
        m = b.get_global(GLOBAL_MIX, j % mix_size)
        if j % 2 == 0:
            d0 = b.get(GROUP_DATA, j % data_cols, 0); d1 = b.get(GROUP_DATA, (j + 1) % data_cols, 1)
            inner = b.and_eqz(inner, b.sub(b.sub(a, b.mul(m, d0)), d1))
        else:
            p1 = b.get(GROUP_ACCUM, j - 1, 1); p0 = b.get(GROUP_ACCUM, j - 1, 0)
            o = b.get_global(GLOBAL_OUT, j % out_size)
            inner = b.and_eqz(inner, b.sub(b.sub(b.sub(a, b.mul(p1, p0)), m), o))
    b.ret = b.and_cond(top, sel, inner)
    return b


SYN280 = dict(accum_cols=40, code_cols=16, data_cols=224, mix_size=20, out_size=32)

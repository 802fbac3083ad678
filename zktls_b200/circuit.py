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
# witness ("accumulate") program: the same value opcodes as PolyExtStep plus Set / Barrier / PrefixProduct (DESIGN.md "circuit blob")
W_CONST, W_GET, W_GET_GLOBAL, W_ADD, W_SUB, W_MUL, W_SET, W_BARRIER, W_PREFIX_PRODUCT = range(9)
W_ALWAYS = 0xFFFFFFFF


class CircuitBuilder:
    def __init__(self, accum_cols, code_cols, data_cols, mix_size, out_size, info=b"SYN:v1__________"):
        assert len(info) == 16
        self.group_size = [accum_cols, code_cols, data_cols]
        self.mix_size, self.out_size, self.info = mix_size, out_size, info
        self.taps, self.steps = [], []
        self.n_fp, self.n_mix = 0, 0
        self.ret = None
        self.wsteps, self.n_w = [], 0          # CircuitHal::accumulate as data: per-row step program over the trace domain

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

    # ---- witness program (CircuitHal::accumulate): values are numbered in definition order; a value may only be used inside the
    # phase (stretch between barriers) that defines it.  Every row of the trace domain runs the same steps.
    def _w(self, op, a=0, b=0, c=0):
        self.wsteps.append((op, a, b, c)); self.n_w += 1; return self.n_w - 1

    def w_const(self, v): return self._w(W_CONST, v % 2013265921)
    def w_get(self, group, column, back=0): return self._w(W_GET, group, column, back)
    def w_get_global(self, base, offset): return self._w(W_GET_GLOBAL, base, offset)
    def w_add(self, a, b): return self._w(W_ADD, a, b)
    def w_sub(self, a, b): return self._w(W_SUB, a, b)
    def w_mul(self, a, b): return self._w(W_MUL, a, b)
    def w_set(self, accum_column, value, cond=None): self.wsteps.append((W_SET, accum_column, value, W_ALWAYS if cond is None else cond))
    def w_barrier(self): self.wsteps.append((W_BARRIER, 0, 0, 0))
    def w_prefix_product(self, first_accum_column): self.wsteps.append((W_PREFIX_PRODUCT, first_accum_column, 0, 0))

    def blob(self):
        assert self.ret is not None
        hdr = [MAGIC] + self.group_size + [self.mix_size, self.out_size, len(self.taps), len(self.steps), self.ret, self.n_fp, self.n_mix, len(self.wsteps)]
        hdr += [int.from_bytes(self.info[4 * i: 4 * i + 4], "little") for i in range(4)]
        words = hdr + [w for t in self.taps for w in t] + [w for s in self.steps for w in s] + [w for s in self.wsteps for w in s]
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
    # witness program of the accum group (what risc0's circuit crates generate as step_compute_accum): on live rows (sel != 0)
    #   phase 0: even column j = m_{j mod M} * d_{j mod D}[i] + d_{(j+1) mod D}[i-1]
    #   phase 1: odd column j  = a_{j-1}[i-1] * a_{j-1}[i] + m_{j mod M} + o_{j mod O}      (reads the finished even columns)
    # the other rows keep what the caller put there (the reference fills its trailing ZK rows with noise the same way)
    for parity in (0, 1):
        wsel = b.w_get(GROUP_CODE, 0, 0)
        for j in range(parity, accum_cols, 2):
            m = b.w_get_global(GLOBAL_MIX, j % mix_size)
            if parity == 0:
                v = b.w_add(b.w_mul(m, b.w_get(GROUP_DATA, j % data_cols, 0)), b.w_get(GROUP_DATA, (j + 1) % data_cols, 1))
            else:
                v = b.w_add(b.w_add(b.w_mul(b.w_get(GROUP_ACCUM, j - 1, 1), b.w_get(GROUP_ACCUM, j - 1, 0)), m), b.w_get_global(GLOBAL_OUT, j % out_size))
            b.w_set(j, v, wsel)
        if parity == 0:
            b.w_barrier()
    return b


SYN280 = dict(accum_cols=40, code_cols=16, data_cols=224, mix_size=20, out_size=32)


def syn_heavy_circuit(accum_cols=40, code_cols=16, data_cols=224, mix_size=20, out_size=32, majors=16, fanout=(4, 4, 2), leaf_constraints=40, seed=0x5EA7):
    """SYN-HEAVY: an eval_check stress circuit with the SHAPE of rv32im's generated `poly_fp` (hard part 4; VERDICT r1
    next #5) on the SYN-280 column layout: by default 16 x 4 x 4 x 2 = 512 leaf blocks of 40 constraints = 20,480 constraints plus the
    selector constraints, > 100 k `PolyExtStep`s, total degree 5, `AndCond` nested four deep (major / minor / sub / leaf selectors taken
    from code columns at back 0 and 2), taps back 0..4 on a third of the data and half of the accum columns, back {0,1,3} / {0,1} on
    the rest, {0,2} on four code columns => five combos.  Constraint operands are drawn pseudo-randomly over ALL columns and taps
    (every column is read hundreds of times, from every part of the program), which is what defeats a streaming column ring.
    Deterministic in `seed`.  The constraints are not meant to be satisfiable: parity and timing use uniform traces (Trace A)."""
    import random
    rnd = random.Random(seed)
    b = CircuitBuilder(accum_cols, code_cols, data_cols, mix_size, out_size, info=("SYNHEAVY%d:v1" % (accum_cols + code_cols + data_cols)).ljust(16, "_").encode()[:16])
    backs = {}
    for c in range(accum_cols):
        backs[(GROUP_ACCUM, c)] = (0, 1, 2, 3, 4) if c % 2 == 0 else (0, 1)
    for c in range(code_cols):
        backs[(GROUP_CODE, c)] = (0, 2) if c < 4 else (0,)
    for c in range(data_cols):
        backs[(GROUP_DATA, c)] = ((0, 1, 2, 3, 4), (0, 1), (0, 1, 3))[c % 3]
    for (g, c), bs in backs.items():
        for k in bs:
            b.add_tap(g, c, k)
    b.finish_taps()
    cols = sorted(backs)
    cache = {}

    def tap(g, c, k):
        if (g, c, k) not in cache:
            cache[(g, c, k)] = b.get(g, c, k)
        return cache[(g, c, k)]

    def any_tap():
        g, c = cols[rnd.randrange(len(cols))]
        return tap(g, c, rnd.choice(backs[(g, c)]))

    def leaf_constraint(chain):
        # degree-1 relation over 3..4 taps and a global: a - (b + k * c) [- m]   (total degree 5 under four selectors)
        a, x, y = any_tap(), any_tap(), any_tap()
        k = b.const(rnd.randrange(1, 1 << 20))
        e = b.sub(a, b.add(x, b.mul(k, y)))
        if rnd.random() < 0.3:
            e = b.sub(e, b.get_global(GLOBAL_MIX, rnd.randrange(mix_size)) if rnd.random() < 0.5 else b.get_global(GLOBAL_OUT, rnd.randrange(out_size)))
        return b.and_eqz(chain, e)

    def selector(level, idx):
        c = (level * 3 + idx) % code_cols
        return tap(GROUP_CODE, c, 2 if (c < 4 and (idx & 1)) else 0)

    top = b.true()
    one = b.const(1)
    for c in range(code_cols):          # selectors are bits: degree-2 constraints at the top level
        s = tap(GROUP_CODE, c, 0)
        top = b.and_eqz(top, b.mul(s, b.sub(s, one)))
    # a few degree-4 / degree-3 relations at shallow depth (products of taps), like the memory / multiply constraints of rv32im
    shallow = b.true()
    for _ in range(64):
        shallow = b.and_eqz(shallow, b.sub(b.mul(b.mul(any_tap(), any_tap()), b.mul(any_tap(), any_tap())), any_tap()))
    top = b.and_cond(top, selector(0, 0), shallow)
    for mj in range(majors):
        lvl1 = b.true()
        for mn in range(fanout[0]):
            lvl2 = b.true()
            for sb in range(fanout[1]):
                lvl3 = b.true()
                for lf in range(fanout[2]):
                    leaf = b.true()
                    for _ in range(leaf_constraints):
                        leaf = leaf_constraint(leaf)
                    lvl3 = b.and_cond(lvl3, selector(3, mj + mn + sb + lf), leaf)
                lvl2 = b.and_cond(lvl2, selector(2, mj + mn + sb), lvl3)
            lvl1 = b.and_cond(lvl1, selector(1, mj + mn), lvl2)
        top = b.and_cond(top, selector(0, mj + 1), lvl1)
    b.ret = top
    return b


SYNHEAVY = dict(SYN280)

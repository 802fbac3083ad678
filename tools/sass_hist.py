"""Dynamic per-permutation opcode histogram of a Poseidon2 kernel from its static SASS.  Loop bodies are found from backward
branches; the outermost loop (the per-row / per-repetition loop) delimits one permutation and the inner loops, in address
order, get the trip counts given with --trips (default 4,21,4: external, internal, external rounds).
usage: cuobjdump -sass -fun <mangled> obj | python tools/sass_hist.py [--trips 4,21,4] [--all]
Pipe model (matches ncu sm__pipe_fmaheavy_cycles_active on k_hash_rows): every IMAD form issues on the 16-lane fmaheavy
pipe, IMAD.WIDE / IMAD.HI occupy it twice as long; every other integer op is one ALU-pipe slot."""
import re, sys, collections
trips = [4, 21, 4]; show_all = False
a = sys.argv[1:]
while a:
    if a[0] == "--trips": trips = [int(x) for x in a[1].split(",")]; a = a[2:]
    elif a[0] == "--all": show_all = True; a = a[1:]
    else: raise SystemExit("unknown arg " + a[0])
pat = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);")
ins = []
for line in sys.stdin:
    m = pat.match(line)
    if m: ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
loops = []
for addr, op, rest in ins:
    if op.startswith("BRA"):
        t = re.search(r"0x([0-9a-f]+)", rest)
        if t and int(t.group(1), 16) < addr: loops.append((int(t.group(1), 16), addr))
loops.sort()
# outermost = the loop containing the most inner loops; use the first maximal one with >= len(trips) inner loops
def inner(l): return [k for k in loops if k != l and l[0] <= k[0] and k[1] <= l[1]]
outer = [l for l in loops if len(inner(l)) >= len(trips)]
rng = outer[0] if outer else (ins[0][0], ins[-1][0])
inn = inner(rng)[: len(trips)] if outer else loops[: len(trips)]
hist = collections.Counter()
for addr, op, _ in ins:
    if not (rng[0] <= addr <= rng[1]): continue
    w = 1
    for (lo, hi), t in zip(inn, trips):
        if lo <= addr <= hi: w *= t
    hist[op] += w
tot = sum(hist.values())
f = sum(n * (2 if (op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI")) else 1) for op, n in hist.items() if op.startswith("IMAD"))
alu = sum(n for op, n in hist.items() if op.split(".")[0] in ("VIADDMNMX", "IADD3", "LOP3", "SHF", "LEA", "VIADD", "ISETP", "SEL", "IMNMX", "VIMNMX", "PRMT", "MOV", "IADD", "VIMNMX3", "SGXT", "BMSK"))
print(f"instr {tot}  fmaheavy-slots {f}  alu-slots {alu}  loops {[(hex(l), hex(h)) for l, h in inn]} in {hex(rng[0])}-{hex(rng[1])}")
for op, n in hist.most_common(40 if show_all else 12): print(f"  {op:22s} {n:8d} {100.0*n/tot:6.2f}%")

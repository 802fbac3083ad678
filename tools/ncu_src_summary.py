"""Summarises an `ncu --page source --csv` dump: executed-instruction mix by opcode and stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); total = 0; samples = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") else parts[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n; total += n
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
    samples += int(r[ix["# Samples"]] or 0)
print("total warp-instructions", total)
for op, n in ops.most_common(25): print(f"  {op:28s} {n:14d} {100.0*n/total:6.2f}%")
print("stall samples", samples)
for s, n in stalls.most_common(10): print(f"  {s:28s} {n:10d} {100.0*n/max(samples,1):6.2f}%")

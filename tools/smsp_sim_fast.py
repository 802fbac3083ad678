"""Front end for the compiled issue simulator (see tools/smsp_sim.py for the model): parses SASS from stdin, flattens the loops,
feeds tools/smsp_sim_core (g++ -O2 of tools/smsp_sim_core.cpp, the loop of tools/smsp_sim.py; built on first use)."""
import subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import smsp_sim as S
_HERE = os.path.dirname(os.path.abspath(__file__))
def _sim_binary():
    exe = os.path.join(_HERE, "smsp_sim_core")
    src = exe + ".cpp"
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-o", exe, src], check=True)
    return exe
def run(sass_text, warps=12, policy="gto", trips=(4, 21, 4), reps=3, sim=None):
    sim = sim or _sim_binary()
    ins = S.parse(sass_text.splitlines(True))
    seq, per = S.build_program(ins, list(trips), reps)
    pid = {"F": 0, "A": 1, "U": 2, "X": 3}
    info = []
    for x in ins:
        (p, occ) = S.pipe_of(x["op"]); info.append((pid[p], occ, max(x["stall"], 1)))
    data = str(len(seq)) + "\n" + "\n".join("%d %d %d" % info[i] for i in seq) + "\n"
    out = subprocess.run([sim, str(warps), "0" if policy == "gto" else "1"], input=data, capture_output=True, text=True).stdout.split()
    t, bf, ba = int(out[0]), int(out[1]), int(out[2])
    return {"slot_times": t / (warps * reps) / 2, "F": 100.0 * bf / t, "A": 100.0 * ba / t, "instr": per}
if __name__ == "__main__":
    a = sys.argv[1:]; warps = int(a[0]) if a else 12; pol = a[1] if len(a) > 1 else "gto"
    r = run(sys.stdin.read(), warps, pol)
    print("slot-times/perm %.0f  F %.1f%%  A %.1f%%  instr %d" % (r["slot_times"], r["F"], r["A"], r["instr"]))

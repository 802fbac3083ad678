"""Builds zktls_b200/libzkb200.so (hand-written CUDA for sm_100a + the C-ABI of include/zkb200.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libzkb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX, "--expt-relaxed-constexpr"] + ARCH


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "zkb200.h"))
    return hs


def build(force=False, verbose=False, ptxas_v=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(h) for h in headers())
    objs, procs = [], []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_mtime):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-x", "cu", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif (verbose or ptxas_v) and out.strip():
            print(f"--- {s} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(SO):
        cmd = [NVCC, "-shared", "-o", SO] + objs + ARCH + ["-ccbin", HOST_CXX, "-Xcompiler", "-fPIC", "-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_v="--ptxas" in sys.argv)
    print(SO)

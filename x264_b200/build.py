"""Builds x264_b200/csrc/libx264_b200.so (sm_100a only) with nvcc.  In-tree, so the .so travels with gpurun."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libx264_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"] + (["-DME_DEBUG"] if os.environ.get("ME_DEBUG") else []) + \
        (["-DLA_PROFILE"] if os.environ.get("LA_PROFILE") else [])


def sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    c = sorted(f for f in os.listdir(CSRC) if f.endswith(".c"))
    return cu, c


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".c"))]
    deps.append(os.path.join(HERE, "..", "include", "x264_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cu, c = sources()
    objs = []
    procs = []
    for f in cu:
        o = os.path.join(CSRC, f[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for f in c:
        o = os.path.join(CSRC, f[:-2] + ".o")
        cmd = ["gcc", "-O2", "-fPIC", "-std=gnu99", "-Wall", "-I" + os.path.join(HERE, "..", "include"),
               "-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (f, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("x264_b200: CUDA build failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

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


def build(force=False, verbose=False, variant=None, defines=()):
    """variant: tuning builds (tools/): libx264_b200_<variant>.so with extra -D flags, objects kept apart; the binding
    loads it instead of the product library when X264CU_LIB points at it."""
    if not variant and not force and not needs_build():
        return LIB
    cu, c = sources()
    objs = []
    procs = []
    lib = LIB if not variant else os.path.join(CSRC, "libx264_b200_%s.so" % variant)
    odir = CSRC if not variant else os.path.join(CSRC, "build", variant)
    os.makedirs(odir, exist_ok=True)
    for f in cu:
        o = os.path.join(odir, f[:-3] + ".o")
        cmd = [NVCC] + FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for f in c:
        o = os.path.join(odir, f[:-2] + ".o")
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
    cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lm", "-ldl"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))

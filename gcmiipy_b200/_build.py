"""Build the C-ABI library `gcmiipy_b200/_lib/libgcm_b200.so` (hand-written sm_100a CUDA) in-tree.

    python -m gcmiipy_b200._build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Objects are rebuilt only when a source or header
is newer.  Per-file flags: the `+ - * /`-only schemes (sw2d, pe2d, ops) and the exact 2.5-D
operators are compiled with -fmad=false so that they reproduce numpy bit for bit.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(HERE, "_lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBNAME = "libgcm_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]
# source -> extra flags
SOURCES = {
    "geom.cu": [],
    "prof.cu": [],
    "tma.cu": [],
    "comm.cu": [],
    "host_step.cu": [],
    "pe25.cu": ["-fmad=false"],
    "pe25_fast.cu": [],
    "pe25_fast_l3.cu": [],
    "pe25_fast_l9.cu": [],
    "pe25_fast_l17.cu": [],
    "pe25_fast_l18.cu": [],
    "pe25_extras.cu": [],
    "sw2d.cu": ["-fmad=false"],
    "pe2d.cu": ["-fmad=false"],
    "ops.cu": ["-fmad=false"],
    "physics.cu": ["-fmad=false"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    headers.append(os.path.abspath(__file__))
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [path] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stdout.write(out)
        if p.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed for: " + ", ".join(failed))
    lib = lib_path()
    if force or procs or _newer(lib, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""TEST INFRASTRUCTURE ONLY: compile the .cu kernel sources for the CPU against tests/emu/cuda_emu.h.

There is no GPU in the build container; the `-m "not gpu"` suite runs the very same kernel sources through
this shim to check indexing / halo / scan logic against the oracle before any GPU minute is spent.  The
library lands in tests/emu/_build/ and is loaded only by tests/emu/emu_lib.py -- never by gcmiipy_b200.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gcmiipy_b200", "csrc")
# GCM_EMU_ASAN=1: AddressSanitizer build (run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so) and
# ASAN_OPTIONS=detect_leaks=0): an out-of-bounds read or write of any kernel shows up on the CPU -- the bounds
# checker for the kernel sources (compute-sanitizer is not available on every GPU pool)
ASAN = os.environ.get("GCM_EMU_ASAN") == "1"
OUT = os.path.join(HERE, "_build_asan" if ASAN else "_build")
LIB = os.path.join(OUT, "libgcm_emu.so")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    deps += [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "build_emu.py")]
    deps.append(os.path.join(ROOT, "include", "gcm_b200.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs, procs = [], []
    flags = ["-O2", "-std=c++17", "-fPIC", "-DGCM_EMU", "-ffp-contract=off", "-I", HERE, "-I", CSRC, "-I",
             os.path.join(ROOT, "include"), "-Wno-unused-value"]
    if ASAN:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "-g", "-O1"]
    for s in srcs + [os.path.join(HERE, "cuda_emu.cpp")]:
        o = os.path.join(OUT, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", s, "-o", o]))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("emu build failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lpthread"] + (["-fsanitize=address"] if ASAN else []))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Builds libpwgs_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpwgs_b200.so")
SOURCES = ["pwgs_api.cu", "pwgs_host_sampler.cpp", "pwgs_host_pickle.cpp"]
DEPS = ["pwgs_api.cu", "pwgs_host_sampler.cpp", "pwgs_host_pickle.cpp", "pwgs_kernels.cuh", "pwgs_device.cuh", os.path.join("..", "..", "include", "pwgs_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu -> phylowgs_b200/libpwgs_b200.so.  Returns the library path.
    `defines` / `out`: tuning builds of kernel variants (-DNAME=VALUE) into another file, loaded with PWGS_LIB=<path>."""
    out = out or LIB
    if out == LIB and not force and not needs_build():
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + ["-D" + d for d in defines]
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build_library(force=True, verbose="-v" in sys.argv, defines=defs, out=os.path.abspath(outs[0]) if outs else None))

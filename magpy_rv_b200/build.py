"""In-tree build of the CUDA shared library (nvcc, sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "abi.cu")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libmagpy_rv_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    d = os.path.join(HERE, "csrc")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    mt = os.path.getmtime(OUT)
    hdr = os.path.join(os.path.dirname(HERE), "include", "magpy_rv_b200.h")
    return any(os.path.getmtime(s) > mt for s in sources() + [hdr])


def build(force=False, verbose=False):
    """Compile csrc/abi.cu (which includes every .cuh) into _lib/libmagpy_rv_b200.so."""
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    extra = os.environ.get("MRV_NVCC_EXTRA", "").split()      # e.g. -DMID_TIMING for the phase timers
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))

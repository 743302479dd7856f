"""In-tree nvcc build of the C-ABI library (sm_100a only).

    python "dl-channel-estimation-mamimo_b200/build.py"      # or __graft_entry__.build()

Produces  dl-channel-estimation-mamimo_b200/libmamimo_b200.so  (git-ignored, travels
with gpurun snapshots).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmamimo_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["fc.cuh", "ls.cuh", "lmmse.cuh", "ofdm.cuh", "ptx.cuh", "schemes.cuh", "tables.h", os.path.join("..", "..", "include", "mamimo.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the sm_100a library")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the library if missing or older than its sources.  Returns the .so path."""
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("MAMIMO_NVCC_EXTRA", "").split()      # e.g. -DMAMIMO_FC_DEBUG_COUNTERS
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

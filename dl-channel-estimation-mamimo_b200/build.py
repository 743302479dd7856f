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


# The MEX gateway executed without MATLAB: mex_gateway.cpp + the functional mini runtime (csrc/mex_runtime) + a C
# driver for the tests, linked against the product library.  Test infrastructure; built here so it travels with it.
MEX_HARNESS_PATH = os.path.join(PKG_DIR, "libmamimo_mex_harness.so")
MEX_SOURCES = ["mex_gateway.cpp", os.path.join("mex_runtime", "mex_runtime.cpp"), os.path.join("mex_runtime", "mex_harness.cpp")]
MEX_HEADERS = [os.path.join("mex_runtime", "mex.h"), os.path.join("mex_runtime", "mex_runtime.hpp"),
               os.path.join("..", "..", "include", "mamimo.h")]


def build_mex_harness(force=False):
    """g++ -shared: the gateway's mexFunction behind a ctypes-callable driver.  Returns the .so path."""
    deps = [os.path.join(CSRC, s) for s in MEX_SOURCES + MEX_HEADERS] + [LIB_PATH]
    if not force and os.path.exists(MEX_HARNESS_PATH) and \
            all(os.path.getmtime(d) <= os.path.getmtime(MEX_HARNESS_PATH) for d in deps if os.path.exists(d)):
        return MEX_HARNESS_PATH
    gxx = shutil.which("g++")
    if not gxx:
        raise RuntimeError("g++ not found: cannot build the MEX harness")
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I" + os.path.join(CSRC, "mex_runtime"), "-I" + os.path.join(PKG_DIR, "..", "include"),
           "-o", MEX_HARNESS_PATH] + [os.path.join(CSRC, s) for s in MEX_SOURCES] + \
          ["-L" + PKG_DIR, "-lmamimo_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (MEX harness) failed:\n" + res.stdout + res.stderr)
    return MEX_HARNESS_PATH


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build the sm_100a library")


def is_stale(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


DEBUG_LIB_PATH = os.path.join(PKG_DIR, "libmamimo_b200_dbg.so")   # -DMAMIMO_FC_DEBUG_COUNTERS build (diagnostics only)


def build_debug(force=False):
    """The same library with the FC role counters compiled in (tools/fc_power_probe.py loads it via MAMIMO_LIB)."""
    if not force and os.path.exists(DEBUG_LIB_PATH) and not is_stale(DEBUG_LIB_PATH):
        return DEBUG_LIB_PATH
    return build(force=True, extra=["-DMAMIMO_FC_DEBUG_COUNTERS"], out=DEBUG_LIB_PATH)


def build(force=False, verbose=False, extra=None, out=None):
    """Compile the library if missing or older than its sources.  Returns the .so path."""
    out = out or LIB_PATH
    if not force and not is_stale(out):
        return out
    extra = list(extra or []) + os.environ.get("MAMIMO_NVCC_EXTRA", "").split()      # e.g. -DMAMIMO_FC_DEBUG_COUNTERS
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_mex_harness(force="--force" in sys.argv))

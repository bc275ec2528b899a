"""In-tree build of the CUDA engine: nvcc -> speedy-ml_b200/lib/libspeedyml_b200.so (sm_100a only).

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only container (the "does it build" check).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libspeedyml_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "speedyml_engine.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "engine.cu")]
    # the image exports CC/CXX wrappers that break nvcc's host pass in some shells; use the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed for speedy-ml_b200/csrc/engine.cu")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


HOST_DRIVER = os.path.join(LIB_DIR, "replay_main")


def build_host_driver(force: bool = False) -> str:
    """the compiled host program on the C++ host layer (g++ only; links the engine library)"""
    srcs = [os.path.join(HERE, "drivers", "replay_main.cpp"), os.path.join(HERE, "host", "speedyml_host.hpp"),
            os.path.join(os.path.dirname(HERE), "include", "speedyml_engine.h")]
    if not force and os.path.exists(HOST_DRIVER) and all(os.path.getmtime(s) < os.path.getmtime(HOST_DRIVER) for s in srcs) \
            and os.path.getmtime(LIB) < os.path.getmtime(HOST_DRIVER):
        return HOST_DRIVER
    build()
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-o", HOST_DRIVER, srcs[0], "-L" + LIB_DIR, "-lspeedyml_b200",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed for speedy-ml_b200/drivers/replay_main.cpp")
    return HOST_DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

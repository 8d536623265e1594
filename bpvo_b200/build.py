"""Builds libbpvo_b200.so (hand-written sm_100a CUDA + the C++ host shim) in-tree with nvcc.

The built .so is git-ignored but travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbpvo_b200.so")
SOURCES = ["engine.cu", "comm.cu", "stereo.cu", os.path.join("host", "vo_shim.cpp")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_source_mtime() -> float:
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    m = max(m, os.path.getmtime(os.path.join(HERE, "..", "include", "bpvo_b200.h")))
    return m


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant "fine": a second library with the fine-grained in-kernel profile marks compiled in
    (-DBP_FINE_PROFILE -> libbpvo_b200_fine.so, loaded when BPVO_B200_LIB points at it); never the product path."""
    lib = LIB if not variant else LIB.replace(".so", f"_{variant}.so")
    extra = {"": [], "fine": ["-DBP_FINE_PROFILE"]}.get(variant)
    if extra is None:                      # ad-hoc A/B variants: "name:-DFOO=1,-DBAR=2"
        variant, _, flags = variant.partition(":")
        extra = [x for x in flags.split(",") if x]
        lib = LIB.replace(".so", f"_{variant}.so")
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest_source_mtime():
        return lib
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
    env = dict(os.environ)
    # the image's default CC points at a gcc wrapper without OpenMP specs; nvcc only needs a host g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant="fine" if "--fine" in sys.argv else next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), "")))

"""Build libcrvae_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcrvae_b200.so")
STAMP = os.path.join(PKG_DIR, ".libcrvae_b200.stamp")
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v", "-lcudart"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(os.path.dirname(PKG_DIR), "include", "crvae_b200.h"))
    for f in files:                      # names relative to the package: the tree is copied to other paths (GPU box)
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    dig = _digest()

    def fresh():
        return os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read() == dig

    if not force and fresh():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_PATH):       # GPU box without a toolchain: use the shipped binary
            return LIB_PATH
        raise RuntimeError("nvcc not found and no prebuilt libcrvae_b200.so present")
    # one builder at a time (torchrun starts N ranks at once); the library is written under a temporary name and renamed,
    # so a concurrent dlopen never sees a half-written file
    import fcntl
    with open(os.path.join(PKG_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():      # another process built it while we waited
                return LIB_PATH
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + _sources()
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = res.stdout + res.stderr
            with open(os.path.join(PKG_DIR, "build.log"), "w") as fh:
                fh.write(" ".join(cmd) + "\n" + log)
            if res.returncode != 0:
                sys.stderr.write(log)
                raise RuntimeError("nvcc failed building libcrvae_b200.so")
            if verbose:
                print(log)
            os.replace(tmp, LIB_PATH)
            with open(STAMP, "w") as fh:
                fh.write(dig)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

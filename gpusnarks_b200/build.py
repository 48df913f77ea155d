"""Build libgpusnarks_b200.so in-tree with nvcc for sm_100a (no JIT cache, no other arch)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libgpusnarks_b200.so")
SRC = os.path.join(HERE, "csrc", "gsn_lib.cu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    out = [SRC]
    for d in (os.path.join(HERE, "csrc"), os.path.join(ROOT, "include")):
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith((".cuh", ".h", ".inl", ".cu"))]
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

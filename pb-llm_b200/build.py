"""Compile libpbllm.so (sm_100a only) in-tree with nvcc. No JIT cache: the .so lives next to
the sources so it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpbllm.so")
SOURCES = ["pbllm_abi.cu", "pbllm_pack.cu", "pbllm_stream.cu", "pbllm_gemv.cu", "pbllm_bireal.cu", "pbllm_decode.cu", "pbllm_gemm_tt.cu", "pbllm_gptq.cu", "pbllm_select.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libpbllm.so cannot be built")


HASHFILE = LIB + ".srchash"


def source_hash() -> str:
    """sha256 over the kernel sources, the public header and the compiler flags: what the .so was built from.
    (mtimes do not survive the snapshot copy to the GPU box; contents do.)"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "pbllm.h")]
    for d in deps:
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()


def built_hash() -> str:
    try:
        with open(HASHFILE) as f:
            return f.read().strip()
    except OSError:
        return ""


def needs_build() -> bool:
    return not os.path.exists(LIB) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(HASHFILE, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

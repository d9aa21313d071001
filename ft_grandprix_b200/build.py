"""In-tree build of libftgp.so (hand-written sm_100a CUDA behind the C ABI of include/ftgp.h).

    python -m ft_grandprix_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands next to this file so that it
travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libftgp.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false" if False else "-Xcompiler", "-fPIC",
    "-shared", "-Xptxas", "-v", "--fmad=true",
]

def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))

def deps():
    out = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "ftgp.h"))
    return out

def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in deps())

def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libftgp.so")
    if verbose:
        print(log)
    return LIB

if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)

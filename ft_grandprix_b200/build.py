"""In-tree build of libftgp.so (hand-written sm_100a CUDA behind the C ABI of include/ftgp.h).

    python ft_grandprix_b200/build.py [--force]

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
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--fmad=true",
]
OBJ = os.path.join(HERE, "_obj")

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
    """One nvcc per translation unit (in parallel), then one link: a change in one kernel recompiles one file."""
    from concurrent.futures import ThreadPoolExecutor
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    headers = [d for d in deps() if not d.endswith(".cu")]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, "", 0
        cmd = [NVCC] + FLAGS + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(compile_one, sources()))
    log = "".join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [r[0] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libftgp.so")
    if verbose:
        print(log)
    return LIB

if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)

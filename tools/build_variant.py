#!/usr/bin/env python
"""A/B variants of libftgp.so for experiments (never shipped, never loaded by the package):
    python tools/build_variant.py NAME [-DMACRO=VALUE ...]   ->  gpurun_out/variants/libftgp_NAME.so
tools/step_ab.py times them on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ft_grandprix_b200", "csrc")
name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "variants"); os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libftgp_{name}.so")
srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
       "-shared", "--fmad=true"] + defs + ["-o", out] + srcs
subprocess.check_call(cmd)
print(out)

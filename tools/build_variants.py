#!/usr/bin/env python
"""Build experiment variants of libmcmil_b200.so (extra -D flags) into build/variants/ (git-ignored, shipped by
gpurun).  Select one at run time with MCMIL_LIB_PATH=build/variants/<name>.so.
    python tools/build_variants.py name1:-DFOO name2:-DBAR=7,-DBAZ"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "montecarlo-gated-mil_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")
SRC = ["api.cu", "pack.cu", "proj_tc.cu", "proj_simt.cu", "reduce.cu", "attnmap.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]

os.makedirs(OUT, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    lib = os.path.join(OUT, name + ".so")
    cmd = ["nvcc", *FLAGS, *defs, "-shared", "-o", lib] + [os.path.join(CSRC, s) for s in SRC]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    out, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + out)

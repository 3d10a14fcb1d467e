"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a.

    python -m gesture2vec_b200.build [--force] [--verbose]

The .so is written next to the sources (gesture2vec_b200/csrc/libg2v_vq.so) so that it
travels to the GPU box with the repo snapshot.  Objects are rebuilt when a source or
header is newer than them.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "include"))
LIB = os.path.join(CSRC, "libg2v_vq.so")
SOURCES = ["g2v_api.cu", "g2v_simt.cu", "g2v_finalize.cu", "g2v_audit.cu", "g2v_gemm.cu", "g2v_soft.cu", "g2v_tc.cu"]
HEADERS = [os.path.join(CSRC, "g2v_common.cuh"), os.path.join(INCLUDE, "g2v_vq.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; the g2v_vq library cannot be built")
    return p


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), suffix: str = "") -> str:
    """Compile csrc/*.cu and link csrc/libg2v_vq<suffix>.so.  `defines` / `suffix` build an experiment
    variant next to the product library (selected at run time with G2V_LIB_PATH)."""
    nvcc = nvcc_path()
    lib = LIB.replace(".so", f"{suffix}.so")
    dflags = [f"-D{d}" for d in defines]
    extra_headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", f"{suffix}.o"))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS + extra_headers):
            cmd = [nvcc] + NVCC_FLAGS + dflags + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stdout.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stdout.write(r.stdout)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    sfx = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--suffix=")), "")
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, suffix=sfx))

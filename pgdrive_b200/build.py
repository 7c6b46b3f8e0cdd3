"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "pgd_step.cu")
SRC_GEN = os.path.join(HERE, "csrc", "pgd_mapgen.cu")
OUT = os.path.join(HERE, "csrc", "libpgdrive_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    # IEEE arithmetic without FMA contraction: the step is checked against a scalar C oracle built with
    # -ffp-contract=off, and contact / done flags depend on exact comparisons
    "-fmad=false", "-Xptxas", "-v"
]


def build_cuda(force=False, verbose=False):
    deps = [SRC, SRC_GEN, os.path.join(HERE, "..", "include", "pgdrive_b200.h"), os.path.join(HERE, "..", "include", "pgd_tables.h")]
    deps += [os.path.join(HERE, "csrc", f) for f in ("pgd_internal.h", "pgd_mapgen.cuh", "pgd_rng.cuh", "pgd_dd.cuh")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT, SRC, SRC_GEN]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
        f.write(res.stdout)
    return OUT


if __name__ == "__main__":
    build_cuda(force=True, verbose=True)

"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(HERE, "..", "include")
OUT = os.path.join(CSRC, "libpgdrive_b200.so")
COMMON = ["pgd_internal.h", os.path.join(INC, "pgdrive_b200.h"), os.path.join(INC, "pgd_tables.h"),
          os.path.join(INC, "pgd_math.h")]
# translation unit -> extra dependencies
UNITS = {
    "pgd_abi.cu": [],
    "pgd_rows.cu": [],
    "pgd_hostpath.cu": ["pgd_hostpool.h"],
    "pgd_step_kernel.cu": ["pgd_step.cuh"],
    "pgd_mapgen.cu": ["pgd_mapgen.cuh", "pgd_rng.cuh", "pgd_dd.cuh"],
}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
    # IEEE arithmetic without FMA contraction: the step is checked against a scalar C oracle built with
    # -ffp-contract=off, contact / done flags depend on exact comparisons, and the map generator must give the
    # same bits as its host build (explicit fma() calls stay fused on both)
    "-fmad=false", "-Xptxas", "-v"
]


def _run(cmd, verbose):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return res.stdout


def build_cuda(force=False, verbose=False, out=OUT, defines=()):
    """Compile every translation unit to an object (re-used while its sources are unchanged) and link the shared
    library.  ``defines`` (e.g. ["-DMIN_CTAS_PER_SM=5"]) builds a kernel variant into ``out``."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = "".join(c if c.isalnum() else "_" for c in "".join(defines))
    objs, jobs, relink = [], [], force or not os.path.exists(out)
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        utag = tag if unit == "pgd_step_kernel.cu" else ""  # variants only differ in the step kernel
        obj = os.path.join(CSRC, unit.replace(".cu", utag + ".o"))
        deps = [src] + [d if os.path.isabs(d) else os.path.join(CSRC, d) for d in COMMON + extra]
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(d) for d in deps):
            cmd = [nvcc] + NVCC_FLAGS + (list(defines) if utag else []) + ["-c", "-o", obj, src]
            jobs.append((cmd, os.path.join(CSRC, unit.replace(".cu", utag + ".ptxas.log"))))
        objs.append(obj)
    if jobs:  # the translation units compile side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(jobs)) as pool:
            texts = list(pool.map(lambda j: _run(j[0], verbose), jobs))
        for (cmd, log_path), text in zip(jobs, texts):
            with open(log_path, "w") as f:
                f.write(text)
        relink = True
    relink = relink or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs)
    if relink:
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lpthread"], verbose)
    return out


if __name__ == "__main__":
    build_cuda(force=True, verbose=True)

"""Build compile-time variants of the step kernel next to the default library and (on a GPU box) time each with
tools/quick_bench.py.  Used for the occupancy / CTA-size experiments recorded in profiles/r01k_experiments.md.

    python tools/kernel_variants.py build      # here (nvcc cross-compiles): writes pgdrive_b200/csrc/libvar_<name>.so
    python tools/kernel_variants.py bench      # on the GPU box: one line per variant and action policy
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {  # step kernel (pgd_step_kernel.cu): warps per CTA, resident CTAs the register budget is set for, L2 hints
    "clk": ["-DPGS_ROLES=4", "-DPGS_MIN_CTAS=4", "-DPGS_PHASE_CLOCKS"],
    "r4c3": ["-DPGS_ROLES=4", "-DPGS_MIN_CTAS=3"],
    "outline": ["-DPGS_OUTLINE_HELPERS"],
    "arcout": ["-DPGS_OUTLINE_ARC"],
}


def main():
    from pgdrive_b200.build import CSRC, build_cuda
    what = sys.argv[1] if len(sys.argv) > 1 else "build"
    if what == "build":
        build_cuda()
        for name, defs in VARIANTS.items():
            print(name, build_cuda(out=os.path.join(CSRC, "libvar_%s.so" % name), defines=defs))
        return
    libs = [("default", "")] + [(n, os.path.join(CSRC, "libvar_%s.so" % n)) for n in VARIANTS]
    for name, lib in libs:
        if lib and not os.path.exists(lib):
            continue
        for actions in ("uniform", "forward"):
            env = dict(os.environ, PGDRIVE_B200_LIB=lib, ACTIONS=actions)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py")], env=env,
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout.strip().split("\n")
            print(name, out[-1], flush=True)


if __name__ == "__main__":
    main()

"""
Builds ``libtsb.so`` (the C-ABI library, include/tsb.h) for sm_100a with nvcc, in-tree.

    python -m tensorsat_b200.build            # build if sources are newer than the library
    python -m tensorsat_b200.build --force

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the
repository snapshot (it is git-ignored, not gpurun-ignored).
"""

from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtsb.so")
SOURCES = ["tsb_ops.cu", "tsb_plan.cu", "tsb_gemm.cu", "tsb_merge_sat.cu", "tsb_merge_gen.cu",
           "tsb_merge_cnt_88.cu", "tsb_merge_cnt_84.cu", "tsb_merge_cnt_48.cu", "tsb_merge_cnt_44.cu", "tsb_merge_wide.cu", "tsb_merge_bits.cu", "tsb_merge_bits_f.cu", "tsb_merge_bits_f42.cu", "tsb_merge_bits_f32.cu", "tsb_merge_bits_g42.cu", "tsb_merge_bits_g32.cu", "tsb_merge_bits_gx1.cu", "tsb_merge_bits_h42.cu", "tsb_merge_bits_i42.cu", "tsb_merge_bits_h32.cu", "tsb_frontend.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtsb.so cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def all_deps() -> list[str]:
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    deps += glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "emu", "*.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tsb.h"))
    return deps


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB, all_deps()):
        build_host_helper()
        return LIB
    objs, procs = [], []
    for src in sources():
        obj = src[:-3] + ".o"
        if force or _stale(obj, all_deps()):
            cmd = [_nvcc(), *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "--expt-relaxed-constexpr", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd)))  # translation units compile in parallel
        objs.append(obj)
    for src, proc in procs:
        if proc.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([_nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-lcuda"], check=True)
    build_host_helper(force)
    return LIB


def build_host_helper(force: bool = False) -> str | None:
    """The CPython helper of the binding (csrc/tsb_pyhost.c: leaf packing for the end-to-end call), in-tree next to the
    package.  Optional: without it ``contract_from_host`` packs with ``bytearray().join``."""
    import sysconfig

    src = os.path.join(CSRC, "tsb_pyhost.c")
    out = os.path.join(HERE, "_tsb_host" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    inc = sysconfig.get_paths().get("include")
    cc = shutil.which("gcc") or shutil.which("cc")
    try:
        import numpy

        np_inc = numpy.get_include()
    except Exception:
        return None
    if not (cc and inc and os.path.exists(os.path.join(inc, "Python.h"))):
        return None
    if force or _stale(out, [src]):
        subprocess.run([cc, "-O2", "-shared", "-fPIC", f"-I{inc}", f"-I{np_inc}", src, "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

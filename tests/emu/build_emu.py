"""
TEST INFRASTRUCTURE: builds ``tests/emu/_build/libtsb_emu.so`` — the same sources as libtsb.so
compiled by g++ with ``-DTSB_HOST_EMU`` against the host emulation of the CUDA execution model
in ``tensorsat_b200/csrc/emu/cuda_emu.h``.  It lets the CPU test-suite run the SIMT kernels and
the whole lowering / plan executor without a GPU.  The product never loads this library.
"""

from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tensorsat_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libtsb_emu.so")
SOURCES = ["tsb_ops.cu", "tsb_plan.cu", "tsb_gemm.cu", "tsb_merge_sat.cu", "tsb_merge_gen.cu",
           "tsb_merge_cnt_88.cu", "tsb_merge_cnt_84.cu", "tsb_merge_cnt_48.cu", "tsb_merge_cnt_44.cu", "tsb_merge_wide.cu", "tsb_merge_bits.cu", "tsb_merge_bits_f.cu", "tsb_merge_bits_f42.cu", "tsb_merge_bits_f32.cu", "tsb_merge_bits_g42.cu", "tsb_merge_bits_g32.cu", "tsb_merge_bits_gx1.cu", "tsb_merge_bits_h42.cu", "tsb_merge_bits_i42.cu", "tsb_merge_bits_h32.cu", "tsb_frontend.cu"]


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    deps = glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(CSRC, "emu", "*"))
    deps.append(os.path.join(ROOT, "include", "tsb.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps if os.path.isfile(d)):
        return LIB
    objs, procs = [], []
    for s in SOURCES:
        obj = os.path.join(OUT, s[:-3] + ".o")
        procs.append(subprocess.Popen(["g++", "-O1", "-std=c++17", "-fPIC", "-DTSB_HOST_EMU", "-x", "c++", "-c",
                                       os.path.join(CSRC, s), "-o", obj, "-Wall", "-Wno-unused-function",
                                       "-Wno-unknown-pragmas", "-Wno-sign-compare"]))
        objs.append(obj)
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("g++ failed")
    subprocess.run(["g++", "-shared", "-o", LIB, *objs], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

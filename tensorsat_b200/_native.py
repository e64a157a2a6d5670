"""
ctypes binding of ``libtsb.so`` (C-ABI in include/tsb.h) for the host-side mirror.

PyTorch supplies device memory (``torch.Tensor`` storage) and the current CUDA stream;
every numeric operation is a call through the C-ABI.  There is no CPU fallback: if the
library has not been built or no CUDA device is present, the first operation raises
``RuntimeError``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Any, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSB_LIB") or os.path.join(_HERE, "csrc", "libtsb.so")   # TSB_LIB: A/B builds of the library

TSB_U8, TSB_I32, TSB_I64 = 0, 1, 2
MODE_SAT, MODE_COUNT = 0, 1
PLAN_NO_GRAPH, PLAN_FORCE_GENERIC, PLAN_NO_GEMM, PLAN_NO_PACK, PLAN_NO_FUSE = 1, 2, 4, 8, 16

_STATUS_EXC = {1: ValueError, 2: RuntimeError, 3: MemoryError, 4: NotImplementedError, 5: OverflowError}

_lib: C.CDLL | None = None
_device: str = "cuda"
_ctx: dict[int, C.c_void_p] = {}


class ArenaStats(C.Structure):
    _fields_ = [
        ("reserved_bytes", C.c_size_t),
        ("in_use_bytes", C.c_size_t),
        ("peak_in_use_bytes", C.c_size_t),
        ("device_allocs", C.c_size_t),
        ("cache_hits", C.c_size_t),
        ("cross_stream_waits", C.c_size_t),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_int32),
        ("leaf_ranks", C.POINTER(C.c_int32)),
        ("leaf_wires", C.POINTER(C.c_int32)),
        ("n_out", C.c_int32),
        ("out_wires", C.POINTER(C.c_int32)),
        ("n_wires", C.c_int32),
        ("wire_extents", C.POINTER(C.c_int32)),
        ("ssa_path", C.POINTER(C.c_int32)),
        ("n_sliced", C.c_int32),
        ("sliced_wires", C.POINTER(C.c_int32)),
        ("mode", C.c_int32),
        ("wire_is_free", C.POINTER(C.c_uint8)),
        ("batch_log2", C.c_int32),
        ("flags", C.c_uint32),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("n_slices", C.c_int64),
        ("peak_bytes", C.c_int64),
        ("out_numel", C.c_int64),
        ("out_dtype", C.c_int32),
        ("n_steps", C.c_int32),
        ("n_kernels_shared", C.c_int32),
        ("n_kernels_per_slice", C.c_int32),
        ("alg_bytes_shared", C.c_double),
        ("alg_bytes_per_slice", C.c_double),
        ("alg_ops_shared", C.c_double),
        ("alg_ops_per_slice", C.c_double),
        ("log2_max_tensor", C.c_int32),
        ("may_overflow", C.c_int32),
        ("alg_bytes_u8_shared", C.c_double),
        ("alg_bytes_u8_per_slice", C.c_double),
        ("n_packed_steps", C.c_int32),
        ("n_conversions", C.c_int32),
        ("n_fused_steps", C.c_int32),
        ("reserved", C.c_int32),
    ]


EXPORTS = [
    "tsb_version", "tsb_last_error", "tsb_ctx_create", "tsb_ctx_destroy", "tsb_ctx_launch_count", "tsb_ctx_trace", "tsb_ctx_trace_json",
    "tsb_arena_reserve", "tsb_arena_alloc", "tsb_arena_release", "tsb_arena_alloc_async", "tsb_arena_release_async", "tsb_arena_trim", "tsb_arena_stats",
    "tsb_permute", "tsb_rewire", "tsb_contract2", "tsb_slice", "tsb_stack", "tsb_union", "tsb_spider",
    "tsb_fill", "tsb_convert", "tsb_all_ones", "tsb_all_binary", "tsb_scalar_to_host", "tsb_plan_create", "tsb_plan_info",
    "tsb_plan_describe", "tsb_plan_execute", "tsb_plan_execute_host", "tsb_plan_set_timing", "tsb_plan_timing_json",
    "tsb_plan_destroy", "tsb_network_from_clauses", "tsb_network_from_dimacs", "tsb_network_set_inputs", "tsb_network_propagate", "tsb_network_info",
    "tsb_network_arrays", "tsb_network_leaf_tensors", "tsb_network_destroy", "tsb_tree_search", "tsb_tree_info",
    "tsb_tree_arrays", "tsb_tree_destroy", "tsb_plan_create_from_tree", "tsb_gemm_u8", "tsb_pack_bits", "tsb_unpack_bits", "tsb_contract2_bits",
]
"""Every symbol include/tsb.h declares (checked by tests/test_abi.py)."""


def load_library(path: str | None = None) -> C.CDLL:
    """Loads the C-ABI library; raises ``RuntimeError`` if it has not been built."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA extension with `python -m tensorsat_b200.build` "
            "(there is no CPU fallback)."
        )
    lib = C.CDLL(path)
    lib.tsb_last_error.restype = C.c_char_p
    lib.tsb_version.restype = C.c_int
    return lib


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def _testing_bind(path: str, device: str) -> None:
    """TEST HOOK: bind to an explicitly given library build (tests/emu) — never called by the product."""
    global _lib, _device
    for c in _ctx.values():
        lib().tsb_ctx_destroy(c)
    _ctx.clear()
    _lib = load_library(path)
    _device = device


def _check(status: int) -> None:
    if status != 0:
        msg = lib().tsb_last_error().decode(errors="replace")
        raise _STATUS_EXC.get(status, RuntimeError)(msg)


def device() -> Any:
    import torch

    if _device == "cuda":
        if not torch.cuda.is_available():
            raise RuntimeError("tensorsat_b200 needs a CUDA device (there is no CPU fallback).")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(_device)


def ctx() -> C.c_void_p:
    """Per-device context (created on first use)."""
    import torch

    dev = torch.cuda.current_device() if _device == "cuda" else 0
    if dev not in _ctx:
        device()
        h = C.c_void_p()
        _check(lib().tsb_ctx_create(dev, C.byref(h)))
        _ctx[dev] = h
    return _ctx[dev]


def stream() -> C.c_void_p:
    import torch

    if _device == "cuda":
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(0)


def _i32(xs: Sequence[int]) -> Any:
    return (C.c_int32 * max(1, len(xs)))(*[int(x) for x in xs])


def _ptr(t: Any) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _torch_dtype(code: int) -> Any:
    import torch

    return {TSB_U8: torch.uint8, TSB_I32: torch.int32, TSB_I64: torch.int64}[code]


def dtype_code(t: Any) -> int:
    import torch

    return {torch.uint8: TSB_U8, torch.int32: TSB_I32, torch.int64: TSB_I64}[t.dtype]


def mode_dtype(mode: int) -> int:
    return TSB_U8 if mode == MODE_SAT else TSB_I64


# ---------------------------------------------------------------------------------- storage
def empty(n: int, mode: int, like: Any = None, dtype: int | None = None) -> Any:
    import torch

    code = mode_dtype(mode) if dtype is None else dtype
    return torch.empty(int(n), dtype=_torch_dtype(code), device=device())


def upload(host: np.ndarray) -> Any:
    import torch

    t = torch.from_numpy(np.array(host, order="C").reshape(-1).copy())
    return t.to(device())


# ---------------------------------------------------------------------------------- box hooks
def permute(src: Any, out: Any, dims: Sequence[int], axes: Sequence[int], elem_size: int) -> None:
    """out (logical layout) = transpose(src viewed with physical ``dims``, axes)."""
    _check(lib().tsb_permute(ctx(), _ptr(src), _ptr(out), len(dims), _i32(dims), _i32(axes), int(elem_size), stream()))


def contract2(
    lhs: Any, ldims: Sequence[int], lw: Sequence[int],
    rhs: Any, rdims: Sequence[int], rw: Sequence[int],
    out: Any, odims: Sequence[int], ow: Sequence[int], mode: int,
) -> None:
    labels: dict[int, int] = {}
    ext: list[int] = []
    for w, d in list(zip(lw, ldims)) + list(zip(rw, rdims)):
        if w not in labels:
            labels[w] = len(labels)
            ext.append(int(d))
    _check(lib().tsb_contract2(
        ctx(), _ptr(lhs), len(lw), _i32([labels[w] for w in lw]), dtype_code(lhs),
        _ptr(rhs), len(rw), _i32([labels[w] for w in rw]), dtype_code(rhs),
        _ptr(out), len(ow), _i32([labels[w] for w in ow]), dtype_code(out),
        len(ext), _i32(ext), int(mode), stream(),
    ))


def rewire(src: Any, dims: Sequence[int], ports: Sequence[int], out: Any, mode: int) -> None:
    _check(lib().tsb_rewire(ctx(), _ptr(src), len(dims), _i32(dims), dtype_code(src), _ptr(out), len(ports),
                            _i32(ports), dtype_code(out), int(mode), stream()))


def slice_(src: Any, dims: Sequence[int], fixed: Sequence[int], out_axes: Sequence[int], out: Any, mode: int) -> None:
    _check(lib().tsb_slice(ctx(), _ptr(src), len(dims), _i32(dims), _i32(fixed), len(out_axes), _i32(out_axes),
                           _ptr(out), src.element_size(), stream()))


def stack(srcs: Sequence[Any], out: Any, dims: Sequence[int], port: int, mode: int) -> None:
    arr = (C.c_void_p * len(srcs))(*[s.data_ptr() for s in srcs])
    _check(lib().tsb_stack(ctx(), arr, len(srcs), len(dims), _i32(dims), int(port), _ptr(out),
                           srcs[0].element_size(), stream()))


def union(a: Any, b: Any, out: Any, mode: int) -> None:
    _check(lib().tsb_union(ctx(), _ptr(a), _ptr(b), _ptr(out), C.c_int64(a.numel()), dtype_code(a), int(mode), stream()))


def spider(out: Any, size: int, num_ports: int, mode: int) -> None:
    _check(lib().tsb_spider(ctx(), _ptr(out), int(size), int(num_ports), dtype_code(out), stream()))


def fill(out: Any, value: int, mode: int) -> None:
    _check(lib().tsb_fill(ctx(), _ptr(out), C.c_int64(out.numel()), C.c_int64(int(value)), dtype_code(out), stream()))


def widen(src: Any, out: Any) -> None:
    _check(lib().tsb_convert(ctx(), _ptr(src), dtype_code(src), _ptr(out), dtype_code(out), C.c_int64(src.numel()), 0,
                             stream()))


def packed_nbytes(n_entries: int) -> int:
    """Bytes of a bit-packed buffer of ``n_entries`` boolean entries (include/tsb.h)."""
    return max(16, (int(n_entries) + 7) // 8)


def pack_bits(src: Any, dst: Any) -> None:
    """``dst`` bit i = ``src[i] != 0`` (``src`` uint8, ``dst`` a uint8 buffer of ``packed_nbytes`` bytes)."""
    _check(lib().tsb_pack_bits(ctx(), _ptr(src), _ptr(dst), C.c_int64(src.numel()), stream()))


def unpack_bits(src: Any, dst: Any, accumulate: bool = False) -> None:
    _check(lib().tsb_unpack_bits(ctx(), _ptr(src), _ptr(dst), C.c_int64(dst.numel()), int(bool(accumulate)), stream()))


def contract2_bits(lhs: Any, lhs_wires: Sequence[int], rhs: Any, rhs_wires: Sequence[int], out: Any,
                   out_wires: Sequence[int], n_wires: int) -> None:
    """Boolean ``_contract2`` on three bit-packed buffers (all wires of extent 2, one wire order)."""
    _check(lib().tsb_contract2_bits(ctx(), _ptr(lhs), len(lhs_wires), _i32(lhs_wires), _ptr(rhs), len(rhs_wires),
                                    _i32(rhs_wires), _ptr(out), len(out_wires), _i32(out_wires), int(n_wires),
                                    _i32([2] * int(n_wires)), stream()))


def all_ones(t: Any, mode: int) -> bool:
    res = C.c_int(0)
    _check(lib().tsb_all_ones(ctx(), _ptr(t), C.c_int64(t.numel()), dtype_code(t), C.byref(res), stream()))
    return bool(res.value)


def all_binary(t: Any) -> bool:
    """True iff every element of the device tensor is 0 or 1 (one kernel, one synchronising read-back)."""
    res = C.c_int(0)
    _check(lib().tsb_all_binary(ctx(), _ptr(t), C.c_int64(t.numel()), dtype_code(t), C.byref(res), stream()))
    return bool(res.value)


def launch_count() -> int:
    n = C.c_int64(0)
    _check(lib().tsb_ctx_launch_count(ctx(), C.byref(n)))
    return int(n.value)


def trace(enabled: bool = True) -> None:
    """Clears and (re)starts the per-kernel launch counters of the current context."""
    _check(lib().tsb_ctx_trace(ctx(), int(bool(enabled))))


def kernel_counts() -> dict[str, int]:
    """``{kernel name: launches}`` since the last ``trace()``."""
    import json

    need = C.c_size_t(0)
    _check(lib().tsb_ctx_trace_json(ctx(), None, C.c_size_t(0), C.byref(need)))
    buf = C.create_string_buffer(need.value)
    _check(lib().tsb_ctx_trace_json(ctx(), buf, need, C.byref(need)))
    return json.loads(buf.value.decode())


def arena_stats() -> dict[str, int]:
    st = ArenaStats()
    _check(lib().tsb_arena_stats(ctx(), C.byref(st)))
    return {f: int(getattr(st, f)) for f, _ in ArenaStats._fields_}


def arena_reserve(nbytes: int) -> None:
    _check(lib().tsb_arena_reserve(ctx(), C.c_size_t(int(nbytes))))


def arena_trim() -> None:
    _check(lib().tsb_arena_trim(ctx()))


def gemm_u8(A: Any, B: Any, Cout: Any, b_bits: int, m_bits: int, n_bits: int, k_bits: int, mode: int) -> None:
    _check(lib().tsb_gemm_u8(ctx(), _ptr(A), _ptr(B), _ptr(Cout), int(b_bits), int(m_bits), int(n_bits), int(k_bits),
                             dtype_code(Cout), int(mode), stream()))


# ---------------------------------------------------------------------------------- plans
class Plan:
    """Owner of a ``tsb_plan`` handle."""

    def __init__(
        self,
        inputs: Sequence[Sequence[int]],
        output: Sequence[int],
        size_dict: Sequence[int],
        ssa_path: Sequence[Sequence[int]],
        sliced_wires: Sequence[int] = (),
        mode: int = MODE_SAT,
        wire_is_free: Sequence[int] | None = None,
        batch_log2: int = -1,
        flags: int = 0,
    ) -> None:
        self._h = C.c_void_p()
        ranks = [len(ws) for ws in inputs]
        flat = [w for ws in inputs for w in ws]
        path = [x for p in ssa_path for x in p]
        d = PlanDesc()
        keep = [_i32(ranks), _i32(flat), _i32(output), _i32(size_dict), _i32(path), _i32(sliced_wires)]
        d.n_leaves = len(inputs)
        d.leaf_ranks, d.leaf_wires = keep[0], keep[1]
        d.n_out, d.out_wires = len(output), keep[2]
        d.n_wires, d.wire_extents = len(size_dict), keep[3]
        d.ssa_path = keep[4]
        d.n_sliced, d.sliced_wires = len(sliced_wires), keep[5]
        d.mode = int(mode)
        if wire_is_free is not None:
            wf = (C.c_uint8 * max(1, len(size_dict)))(*[int(bool(x)) for x in wire_is_free])
            keep.append(wf)
            d.wire_is_free = wf
        d.batch_log2 = int(batch_log2)
        d.flags = int(flags)
        _check(lib().tsb_plan_create(ctx(), C.byref(d), C.byref(self._h)))
        self.mode = int(mode)
        self.n_leaves = len(inputs)
        info = PlanInfo()
        _check(lib().tsb_plan_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in PlanInfo._fields_}

    @classmethod
    def from_handle(cls, handle: C.c_void_p, n_leaves: int, mode: int) -> "Plan":
        """Wraps a ``tsb_plan*`` created elsewhere (``tsb_plan_create_from_tree``)."""
        self = cls.__new__(cls)
        self._h = handle
        self.mode = int(mode)
        self.n_leaves = int(n_leaves)
        info = PlanInfo()
        _check(lib().tsb_plan_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in PlanInfo._fields_}
        return self

    def execute(self, leaves: Sequence[Any], out: Any, slice_begin: int = 0, slice_end: int | None = None) -> None:
        if len(leaves) != self.n_leaves:
            raise ValueError(f"Expected {self.n_leaves} leaf tensors, got {len(leaves)}.")
        if slice_end is None:
            slice_end = self.info["n_slices"]
        self.execute_ptrs([t.data_ptr() for t in leaves], out.data_ptr(), slice_begin, slice_end)

    def execute_ptrs(self, leaf_ptrs: Sequence[int], out_ptr: int, slice_begin: int, slice_end: int) -> None:
        """Same, on raw device addresses (leaves packed in one staging buffer)."""
        if len(leaf_ptrs) != self.n_leaves:
            raise ValueError(f"Expected {self.n_leaves} leaf tensors, got {len(leaf_ptrs)}.")
        if isinstance(leaf_ptrs, C.Array):   # a pointer table the caller keeps across calls
            arr = leaf_ptrs
        else:
            arr = (C.c_void_p * max(1, len(leaf_ptrs)))(*[int(p) for p in leaf_ptrs])
        _check(lib().tsb_plan_execute(self._h, arr, C.c_int64(int(slice_begin)), C.c_int64(int(slice_end)),
                                      C.c_void_p(int(out_ptr)), stream()))

    def execute_host(self, host_leaves_ptr: int, leaves_bytes: int, host_out_ptr: int, out_bytes: int,
                     slice_begin: int = 0, slice_end: int | None = None) -> None:
        """``tsb_plan_execute_host``: leaves back to back in host memory in, result in host memory out; returns when
        the result has arrived."""
        hi = self.info["n_slices"] if slice_end is None else slice_end
        _check(lib().tsb_plan_execute_host(self._h, C.c_void_p(int(host_leaves_ptr)), C.c_size_t(int(leaves_bytes)),
                                           C.c_int64(int(slice_begin)), C.c_int64(int(hi)), C.c_void_p(int(host_out_ptr)),
                                           C.c_size_t(int(out_bytes)), stream()))

    def set_timing(self, enabled: bool | int) -> None:
        """0 / False: off; 1 / True: per-launch (host-synchronised); 2: phase boundaries only (include/tsb.h)."""
        _check(lib().tsb_plan_set_timing(self._h, int(enabled)))

    def describe(self) -> dict[str, Any]:
        import json

        need = C.c_size_t(0)
        _check(lib().tsb_plan_describe(self._h, None, C.c_size_t(0), C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().tsb_plan_describe(self._h, buf, need, C.byref(need)))
        return json.loads(buf.value.decode())

    def close(self) -> None:
        if self._h:
            lib().tsb_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

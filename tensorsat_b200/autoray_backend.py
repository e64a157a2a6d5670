"""
Backend functions for `autoray <https://github.com/jcmgray/autoray>`_ dispatch: the drop-in for
``tensorsat/_autoray.py`` (einsum :31-62, transpose :64-67, shape :69-73, stack :75-80), plus
``tensordot`` and ``reshape`` so that cotengra's non-einsum route (``prefer_einsum=False``) works.

cotengra reaches these by name through ``autoray.do(name, *boxes)`` once the box classes are
registered (``register()``; the reference does this in its ``BoxMeta``, diagrams.py:682-688).
autoray is optional: without it the functions are still importable and callable directly.
"""

from __future__ import annotations

from collections.abc import Sequence
from math import prod
from typing import Any

from .diagrams import Box, FinSet, TensorLikeBox
from .lang.fin_rel import CudaCountRel, CudaFinRel, _CudaTensorBox

BACKEND_NAME = "tensorsat_b200.autoray_backend"


def einsum(contraction: str, /, lhs: Box, rhs: Box | None = None) -> Box:
    """Contracts boxes using einsum notation (letters -> wire ids as in _autoray.py:53-60)."""
    if not isinstance(contraction, str):
        raise TypeError("Contraction must be a string.")
    _input_wires, _out_wires = contraction.split("->")
    if not isinstance(lhs, Box):
        raise TypeError("Operands must be boxes.")
    if rhs is None:
        # The reference raises NotImplementedError here (_autoray.py:41-50); a single-operand
        # einsum is a rewire with summed / repeated ports, which the box API already defines.
        char_idxs = {c: i for i, c in reversed(list(enumerate(_input_wires)))}
        if len(set(_input_wires)) != len(_input_wires):
            raise NotImplementedError("Repeated input indices in a single-operand einsum.")
        return lhs.rewire([char_idxs[c] for c in _out_wires])
    if not isinstance(rhs, Box):
        raise TypeError("Operands must be boxes.")
    box_class = Box.class_join([type(lhs), type(rhs)])
    _lhs_wires, _rhs_wires = _input_wires.split(",")
    char_idxs = {c: i for i, c in reversed(list(enumerate(_lhs_wires + _rhs_wires)))}
    lhs_wires = [char_idxs[c] for c in _lhs_wires]
    rhs_wires = [char_idxs[c] for c in _rhs_wires]
    out_wires = [char_idxs[c] for c in _out_wires]
    return box_class.contract2(lhs, lhs_wires, rhs, rhs_wires, out_wires)


def tensordot(a: Box, b: Box, axes: Any = 2) -> Box:
    """``numpy.tensordot`` semantics on boxes: sum over the paired axes, keep the rest in order."""
    na, nb = len(a.shape), len(b.shape)
    if isinstance(axes, int):
        axes_a, axes_b = list(range(na - axes, na)), list(range(axes))
    else:
        axes_a, axes_b = axes
        axes_a = [axes_a] if isinstance(axes_a, int) else list(axes_a)
        axes_b = [axes_b] if isinstance(axes_b, int) else list(axes_b)
    if len(axes_a) != len(axes_b):
        raise ValueError("tensordot: axes lists must have the same length.")
    axes_a = [x % na for x in axes_a]
    axes_b = [x % nb for x in axes_b]
    lhs_wires = list(range(na))
    rhs_wires = [na + j for j in range(nb)]
    for x, y in zip(axes_a, axes_b):
        rhs_wires[y] = lhs_wires[x]
    out = [w for i, w in enumerate(lhs_wires) if i not in axes_a] + [w for j, w in enumerate(rhs_wires) if j not in axes_b]
    box_class = Box.class_join([type(a), type(b)])
    return box_class.contract2(a, lhs_wires, b, rhs_wires, out)


def transpose(box: Box, perm: Sequence[int] | None = None, /) -> Box:
    """Rearranges the ports of a box (_autoray.py:64-67)."""
    if not isinstance(box, Box):
        raise TypeError("Operand must be a box.")
    if perm is None:
        perm = list(range(len(box.shape)))[::-1]
    return box.rewire(list(perm))


def reshape(box: Box, shape: Sequence[int]) -> Box:
    """Regroups ports without moving data (needed by cotengra's transpose/reshape/matmul lowering)."""
    if not isinstance(box, _CudaTensorBox):
        raise TypeError("reshape is only supported by device tensor boxes.")
    shape = [int(s) for s in (shape if not isinstance(shape, int) else [shape])]
    n = prod(box.tensor_shape)
    if shape.count(-1) == 1:
        rest = prod(s for s in shape if s != -1)
        shape[shape.index(-1)] = n // max(rest, 1)
    if prod(shape) != n or any(s <= 0 for s in shape):
        raise ValueError(f"cannot reshape box of shape {box.tensor_shape} into {tuple(shape)}")
    return type(box)._new(tuple(shape), dev=box.device_tensor.reshape(-1))


def shape(box: TensorLikeBox) -> tuple[int, ...]:
    """Integer shape (_autoray.py:69-73)."""
    if not isinstance(box, TensorLikeBox):
        raise TypeError("Autoray shape is only supported by tensor-like boxes.")
    return box.tensor_shape


def stack(boxes: Sequence[Box], axis: int = 0) -> Box:
    """Stacks boxes along a new axis (_autoray.py:75-80)."""
    boxes = list(boxes)
    box_class = Box.class_join([type(b) for b in boxes])
    return box_class.stack(boxes, axis)


def register() -> bool:
    """Registers the device box classes with autoray, if autoray is installed."""
    try:
        import autoray  # type: ignore[import-not-found]
    except ModuleNotFoundError:
        return False
    for cls in (CudaFinRel, CudaCountRel):
        autoray.register_backend(cls, BACKEND_NAME)
    try:
        autoray.shape.register(BACKEND_NAME)(shape)
    except Exception:  # pragma: no cover - older autoray
        pass
    return True


register()

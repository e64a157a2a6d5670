"""
Device-resident relation boxes: the B200 drop-in for ``tensorsat.lang.fin_rel.FinRel``.

``CudaFinRel`` implements exactly the protected hooks a box class must supply
(tensorsat/diagrams.py:789-996) with the semantics of the reference's numpy ``FinRel``
(tensorsat/lang/fin_rel.py:255-346,379-393,472-473); every hook that touches tensor data
is one call into the C-ABI of ``libtsb.so`` (include/tsb.h) on the current CUDA stream:

===================  ==========================================  ====================
hook                 reference (tensorsat/lang/fin_rel.py)       C-ABI entry point
===================  ==========================================  ====================
``_contract2``       :255-288 einsum in widened dtype + sign      ``tsb_contract2``
``_rewire``          :379-393 lazy transpose / rewire_array       ``tsb_rewire``
``_getitem``         :342-343 ``tensor[idxs]``                    ``tsb_slice``
``_union``           :345-346 ``a | b``                           ``tsb_union``
``_stack``           :304-308 ``np.stack``                        ``tsb_stack``
``_spider``          :290-298 delta tensor                        ``tsb_spider``
``_scalar``          :300-302                                     ``tsb_spider`` (rank 0)
``_as_scalar``       :472-473 ``np.all(tensor == 1)``             ``tsb_all_ones``
===================  ==========================================  ====================

There is no CPU arithmetic path: a hook raises ``RuntimeError`` if the extension or a
CUDA device is missing.  Host numpy arrays appear only as staging for the tiny leaf
constants (``from_set`` & co., uploaded on first use) and for read-back (``.tensor``).

``CudaCountRel`` is the same box over the counting semiring (no clamp, ``int64``), the
mode the reference only has as a notebook prototype
(notebooks/prototypes/2025-02-13-#sat-tensor-contraction.ipynb:215,244).

Semantic note: the reference clamps with ``np.sign(x, dtype=uint8)``,
which casts a widened count to ``uint8`` *before* taking the sign, so a count that is a
non-zero multiple of 256 clamps to 0.  ``CudaFinRel`` implements the boolean semiring
(count > 0) exactly; the oracle can reproduce the quirk to explain any disagreement.
"""

from __future__ import annotations

from collections.abc import Callable, Iterable, Mapping, Sequence
from math import prod
from typing import Any, ClassVar, Self

import numpy as np

from ..diagrams import FinSet, Port, TensorLikeBox, Type, Wire

Size = int
El = int
Point = tuple[int, ...]

MODE_SAT = 0
MODE_COUNT = 1


def _extract_sizes(sizes_or_finsets: Any, /) -> tuple[int, ...]:
    if isinstance(sizes_or_finsets, int):
        return (sizes_or_finsets,)
    if isinstance(sizes_or_finsets, FinSet):
        return (sizes_or_finsets.tensor_dim,)
    return tuple(s if isinstance(s, int) else s.tensor_dim for s in sizes_or_finsets)


def _wrap_el(el_or_point: Any, /) -> Point:
    if isinstance(el_or_point, (int, np.integer)):
        return (int(el_or_point),)
    return tuple(int(x) for x in el_or_point)


class _CudaTensorBox(TensorLikeBox):
    """Common implementation of the two semirings; not instantiable itself."""

    _MODE: ClassVar[int]
    _NP_DTYPE: ClassVar[Any]

    # ------------------------------------------------------------------ construction
    @classmethod
    def _new(
        cls,
        dims: Sequence[int],
        *,
        host: np.ndarray | None = None,
        dev: Any = None,
        axes: Sequence[int] | None = None,
        name: str | None = None,
        binary: bool | None = None,
    ) -> Self:
        if cls in (_CudaTensorBox,):
            raise TypeError("Only final subclasses of Box can be instantiated.")
        self = object.__new__(cls)
        self._name = name
        self._dims = tuple(int(d) for d in dims)
        self._host = host
        self._dev = dev
        self._axes = None if axes is None else tuple(int(a) for a in axes)
        self._hash_cache = None
        # every entry is 0 or 1: always for the boolean semiring; for counting boxes known for host-built ones,
        # checked on the device on demand otherwise (is_binary)
        self._binary = True if cls._MODE == MODE_SAT else binary
        return self

    def __new__(cls, tensor: Any) -> Self:
        """Constructs a box from a host array (validated, copied) or a CUDA torch tensor."""
        try:
            import torch

            if isinstance(tensor, torch.Tensor):
                return cls.from_device(tensor)
        except ModuleNotFoundError:  # pragma: no cover
            pass
        arr = np.array(tensor, order="C")  # (ascontiguousarray would promote rank 0 to rank 1)
        if cls._MODE == MODE_SAT:
            if arr.dtype != np.uint8:
                raise TypeError("Boolean tensors must have dtype uint8.")
            if not np.all(arr <= 1):
                raise ValueError("Values in a Boolean tensor must be 0 or 1.")
        else:
            if not np.issubdtype(arr.dtype, np.integer):
                raise TypeError("Counting tensors must have an integer dtype.")
        if any(d == 0 for d in arr.shape):
            raise ValueError("Zero dimension in shape.")
        arr = arr.astype(cls._NP_DTYPE, copy=True)
        arr.setflags(write=False)
        return cls._new(arr.shape, host=arr, binary=bool(arr.size == 0 or (arr.min() >= 0 and arr.max() <= 1)))

    @classmethod
    def from_device(cls, tensor: Any, *, name: str | None = None) -> Self:
        """Wraps a CUDA torch tensor (made contiguous; trusted to respect the value range)."""
        import torch

        want = torch.uint8 if cls._MODE == MODE_SAT else torch.int64
        if not tensor.is_cuda:
            raise ValueError("from_device expects a CUDA tensor.")
        if tensor.dtype != want:
            raise TypeError(f"Expected a tensor of dtype {want}.")
        dims = tuple(tensor.shape)
        return cls._new(dims, dev=tensor.contiguous().reshape(-1), name=name)

    @classmethod
    def from_set(cls, shape: Any, points: Iterable[El | Point], *, name: str | None = None) -> Self:
        """Relation from a set of points (fin_rel.py:171-196)."""
        shape = _extract_sizes(shape)
        if any(d <= 0 for d in shape):
            raise ValueError("Zero dimension in shape.")
        data = np.zeros(shape, dtype=cls._NP_DTYPE)
        for point in points:
            point = _wrap_el(point)
            if len(point) != len(shape):
                raise ValueError(f"Length of {point = } is invalid for {shape = }.")
            if not all(0 <= i < d for i, d in zip(point, shape)):
                raise ValueError(f"Values of {point = } are invalid for {shape = }.")
            data[point] = 1
        data.setflags(write=False)
        return cls._new(shape, host=data, name=name, binary=True)

    @classmethod
    def from_mapping(
        cls,
        input_shape: Any,
        output_shape: Any,
        mapping: Mapping[Point, El | Point],
        *,
        name: str | None = None,
    ) -> Self:
        """Function graph from a mapping (fin_rel.py:198-221)."""
        input_shape = _extract_sizes(input_shape)
        output_shape = _extract_sizes(output_shape)
        rel = cls.from_set(
            input_shape + output_shape,
            (tuple(k) + _wrap_el(v) for k, v in mapping.items()),
            name=name,
        )
        if len(mapping) != prod(input_shape):
            raise ValueError("Mapping does not cover the entire input space.")
        return rel

    @classmethod
    def singleton(cls, shape: Any, point: El | Point, *, name: str | None = None) -> Self:
        """Singleton relation (fin_rel.py:223-232)."""
        return cls.from_mapping((), shape, {(): point}, name=name)

    @classmethod
    def from_callable(
        cls,
        input_shape: Any,
        output_shape: Any,
        func: Callable[..., El | Point],
        *,
        name: str | None = None,
    ) -> Self:
        """Function graph from a callable (fin_rel.py:234-253)."""
        input_shape = _extract_sizes(input_shape)
        output_shape = _extract_sizes(output_shape)
        mapping = {idx: func(*idx) for idx in np.ndindex(input_shape)}
        return cls.from_mapping(input_shape, output_shape, mapping, name=name)

    # ------------------------------------------------------------------ storage
    @property
    def shape(self) -> tuple[FinSet, ...]:  # type: ignore[override]
        return tuple(FinSet(d) for d in self._dims)

    @property
    def tensor_shape(self) -> tuple[int, ...]:
        return self._dims

    @property
    def is_resident(self) -> bool:
        return self._dev is not None

    def is_binary(self) -> bool:
        """Whether every entry is 0 or 1 (a relation's indicator).  Whole-tree plans size their intermediates on
        that assumption, so ``TreeContraction`` checks it for counting leaves; device-produced boxes are checked
        by one kernel (``tsb_all_binary``), once."""
        if self._binary is None:
            from .. import _native

            dev, _, _ = self._physical()
            self._binary = _native.all_binary(dev)
        return self._binary

    def _physical(self) -> tuple[Any, tuple[int, ...], tuple[int, ...]]:
        """``(flat device tensor, physical dims, axes)`` with ``logical = transpose(phys, axes)``."""
        from .. import _native

        if self._dev is None:
            assert self._host is not None
            self._dev = _native.upload(self._host)
            self._axes = None
        axes = self._axes if self._axes is not None else tuple(range(len(self._dims)))
        pdims = [0] * len(axes)
        for i, a in enumerate(axes):
            pdims[a] = self._dims[i]
        return self._dev, tuple(pdims), axes

    def _physical_wires(self, wires: Sequence[Wire]) -> tuple[Any, tuple[int, ...], list[int]]:
        dev, pdims, axes = self._physical()
        pw = [0] * len(axes)
        for i, a in enumerate(axes):
            pw[a] = int(wires[i])
        return dev, pdims, pw

    @property
    def device_tensor(self) -> Any:
        """The relation as a CUDA ``torch.Tensor`` of the logical shape (materialised)."""
        from .. import _native

        dev, pdims, axes = self._physical()
        if axes != tuple(range(len(axes))):
            out = _native.empty(dev.numel(), self._MODE, like=dev)
            _native.permute(dev, out, pdims, axes, dev.element_size())
            self._dev, self._axes = out, None
            dev = out
        return dev.view(self._dims) if self._dims else dev.view(())

    @property
    def tensor(self) -> np.ndarray:
        """The tensor as a read-only numpy array (device -> host copy when resident)."""
        if self._host is None:
            arr = self.device_tensor.cpu().numpy()
            arr.setflags(write=False)
            self._host = arr
        return self._host

    # ------------------------------------------------------------------ hooks
    @classmethod
    def _contract2(
        cls,
        lhs: Self,
        lhs_wires: Sequence[Wire],
        rhs: Self,
        rhs_wires: Sequence[Wire],
        out_wires: Sequence[Wire],
    ) -> Self:
        from .. import _native

        if type(lhs) is not type(rhs) or not hasattr(cls, "_MODE") or not isinstance(lhs, cls):
            raise TypeError("Operands must be boxes of the same semiring (CudaFinRel or CudaCountRel).")
        ldev, ldims, lw = lhs._physical_wires(lhs_wires)
        rdev, rdims, rw = rhs._physical_wires(rhs_wires)
        dim_of: dict[int, int] = {}
        for w, d in list(zip(lw, ldims)) + list(zip(rw, rdims)):
            if dim_of.setdefault(w, d) != d:
                raise ValueError(f"Wire {w} is attached to ports of different dimensions.")
        out_wires = [int(w) for w in out_wires]
        odims = tuple(dim_of[w] for w in out_wires)
        out = _native.empty(prod(odims), cls._MODE, like=ldev)
        _native.contract2(ldev, ldims, lw, rdev, rdims, rw, out, odims, out_wires, cls._MODE)
        return cls._new(odims, dev=out)

    def _rewire(self, out_ports: Sequence[Port]) -> Self:
        from .. import _native

        out_ports = [int(p) for p in out_ports]
        n = len(self._dims)
        if len(set(out_ports)) == len(out_ports) == n:
            # pure permutation: lazy, like the reference's np.transpose view (fin_rel.py:381-382)
            axes = self._axes if self._axes is not None else tuple(range(n))
            new = type(self)._new(
                tuple(self._dims[p] for p in out_ports),
                host=None if self._host is None else np.transpose(self._host, out_ports),
                dev=self._dev,
                axes=tuple(axes[p] for p in out_ports),
                binary=self._binary,
            )
            if new._dev is None:
                new._axes = None
                new._host = np.array(new._host, order="C")
                new._host.setflags(write=False)
            return new
        dev, pdims, axes = self._physical()
        odims = tuple(self._dims[p] for p in out_ports)
        out = _native.empty(prod(odims), self._MODE, like=dev)
        _native.rewire(dev, pdims, [axes[p] for p in out_ports], out, self._MODE)
        return type(self)._new(odims, dev=out)

    @classmethod
    def _spider(cls, t: Type, num_ports: int) -> Self:
        from .. import _native

        assert isinstance(t, FinSet)
        size = t.tensor_dim
        out = _native.empty(size**num_ports, cls._MODE)
        _native.spider(out, size, num_ports, cls._MODE)
        return cls._new((size,) * num_ports, dev=out)

    @classmethod
    def _scalar(cls, scalar: bool, /) -> Self:
        from .. import _native

        out = _native.empty(1, cls._MODE)
        _native.fill(out, int(scalar), cls._MODE)
        return cls._new((), dev=out)

    @classmethod
    def _stack(cls, boxes: tuple[Self, ...], port: int = 0, /) -> Self:
        from .. import _native

        dims = boxes[0]._dims
        for b in boxes[1:]:
            if b._dims != dims:
                raise ValueError("Shapes of stacked boxes must coincide.")
        srcs = [b.device_tensor.reshape(-1) for b in boxes]
        odims = dims[:port] + (len(boxes),) + dims[port:]
        out = _native.empty(prod(odims), cls._MODE, like=srcs[0])
        _native.stack(srcs, out, dims, port, cls._MODE)
        return cls._new(odims, dev=out)

    def _getitem(self, idxs: tuple[int | slice, ...]) -> Self:
        from .. import _native

        n = len(self._dims)
        if len(idxs) > n:
            raise IndexError("Too many indices for box.")
        fixed = [-1] * n
        for p, idx in enumerate(idxs):
            if isinstance(idx, (int, np.integer)):
                i = int(idx)
                if not -self._dims[p] <= i < self._dims[p]:
                    raise IndexError(f"Index {i} is out of bounds for port {p}.")
                fixed[p] = i % self._dims[p]
        dev, pdims, axes = self._physical()
        pfixed = [-1] * n
        for p in range(n):
            pfixed[axes[p]] = fixed[p]
        # surviving logical ports, in order, and the physical axes they live on
        keep = [p for p in range(n) if fixed[p] < 0]
        odims = tuple(self._dims[p] for p in keep)
        out = _native.empty(prod(odims), self._MODE, like=dev)
        _native.slice_(dev, pdims, pfixed, [axes[p] for p in keep], out, self._MODE)
        return type(self)._new(odims, dev=out)

    def _union(self, other: Self) -> Self:
        from .. import _native

        a = self.device_tensor.reshape(-1)
        b = other.device_tensor.reshape(-1)
        out = _native.empty(a.numel(), self._MODE, like=a)
        _native.union(a, b, out, self._MODE)
        return type(self)._new(self._dims, dev=out)

    def _as_scalar(self) -> bool:
        from .. import _native

        dev, _, _ = self._physical()
        return bool(_native.all_ones(dev, self._MODE))

    # ------------------------------------------------------------------ inspectors (host side)
    def to_set(self) -> frozenset[Point]:
        """Points of the relation (fin_rel.py:395-403)."""
        tensor = self.tensor
        return frozenset(tuple(int(i) for i in idx) for idx in np.argwhere(tensor != 0))

    def _function_matrix(self, input_ports: tuple[int, ...]) -> tuple[tuple[int, ...], np.ndarray]:
        ports = range(len(self._dims))
        if not all(p in ports for p in input_ports):
            raise ValueError("Invalid input ports.")
        if len(input_ports) != len(set(input_ports)):
            raise ValueError("Input ports cannot be repeated.")
        output_ports = tuple(p for p in ports if p not in input_ports)
        isz = prod(self._dims[i] for i in input_ports)
        osz = prod(self._dims[o] for o in output_ports)
        t = np.transpose(self.tensor, input_ports + output_ports)
        return output_ports, t.reshape(isz, osz)

    def is_function_graph(self, input_ports: Sequence[Port], /) -> bool:
        """fin_rel.py:424-431."""
        _, matrix = self._function_matrix(tuple(input_ports))
        return bool(np.all(np.count_nonzero(matrix, axis=1) == 1))

    def to_mapping(self, input_ports: Sequence[Port], /) -> Mapping[Point, Point]:
        """fin_rel.py:433-455."""
        input_ports = tuple(input_ports)
        output_ports, matrix = self._function_matrix(input_ports)
        if not np.all(np.count_nonzero(matrix, axis=1) == 1):
            raise ValueError("Relation is not a function graph with the given input ports.")
        isz = tuple(self._dims[i] for i in input_ports)
        osz = tuple(self._dims[o] for o in output_ports)
        out_points = list(np.ndindex(osz))
        return {
            in_point: out_points[int(np.argmax(row))]
            for in_point, row in zip(np.ndindex(isz), matrix)
        }

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, _CudaTensorBox) or type(other) is not type(self):
            return NotImplemented
        if self is other:
            return True
        return self._dims == other._dims and np.array_equal(self.tensor, other.tensor)

    def __hash__(self) -> int:
        if self._hash_cache is None:
            self._hash_cache = hash((self._dims, np.array(self.tensor, order="C").tobytes()))
        return self._hash_cache


class CudaFinRel(_CudaTensorBox):
    """Boolean-semiring relation box (0/1 ``uint8``), device resident.  Final."""

    _MODE = MODE_SAT
    _NP_DTYPE = np.uint8

    def __init_subclass__(cls, **kw: Any) -> None:
        raise TypeError("CudaFinRel is final.")


class CudaCountRel(_CudaTensorBox):
    """Counting-semiring box (``int64``, no clamp): #SAT mode.  Final.

    ``_union`` adds (slices partition the assignments), ``int(box)`` reads a rank-0 count.
    """

    _MODE = MODE_COUNT
    _NP_DTYPE = np.int64

    def __init_subclass__(cls, **kw: Any) -> None:
        raise TypeError("CudaCountRel is final.")

    def _as_scalar(self) -> bool:
        return int(self) != 0

    def __int__(self) -> int:
        if self._dims:
            raise ValueError("Box is not a scalar.")
        return int(self.tensor.reshape(()))

    @classmethod
    def from_rel(cls, rel: CudaFinRel) -> "CudaCountRel":
        """The 0/1 indicator of a relation as a counting tensor."""
        if rel._host is not None:
            host = rel._host.astype(np.int64)
            host.setflags(write=False)
            return cls._new(rel._dims, host=host, name=rel.name, binary=True)
        from .. import _native

        src = rel.device_tensor.reshape(-1)
        out = _native.empty(src.numel(), MODE_COUNT, like=src)
        _native.widen(src, out)
        return cls._new(rel._dims, dev=out, name=rel.name, binary=True)


FinRel = CudaFinRel
"""Alias so that code written against ``tensorsat.lang.fin_rel.FinRel`` reads unchanged."""

"""
The file a maintainer adds to the reference as ``tensorsat/lang/cuda_fin_rel.py``:
a ``@final`` subclass of the REAL ``tensorsat.diagrams.TensorLikeBox`` whose protected hooks run on the
device through the C-ABI of ``libtsb.so``.

* The class is created by the reference's own ``BoxMeta`` (tensorsat/diagrams.py:667-690), which registers it
  with autoray under ``"tensorsat._autoray"``; the reference's backend functions ``einsum / transpose / shape /
  stack`` (tensorsat/_autoray.py:31-80) only call ``Box.contract2 / rewire / stack / tensor_shape`` and so work
  on these boxes unchanged, as do ``SimpleContraction.contract`` (tensorsat/contractions/simple.py:251-292),
  ``Box.prod / from_wiring / spider / scalar / union`` (diagrams.py:719-809,868-901) and
  ``Box.__getitem__ / __add__ / __bool__`` (diagrams.py:949-988).
* Hook contracts: diagrams.py:789-795 (``_spider``), :806-809 (``_scalar``), :846-866 (``_contract2``),
  :898-901 (``_stack``), :931-947 (``_rewire`` / ``_as_scalar``), :990-996 (``_getitem`` / ``_union``),
  :1013-1031 (``TensorLikeBox.shape``).
* Every hook is ONE call into ``tensorsat_b200.lang.fin_rel.CudaFinRel``'s hook of the same name, i.e. one
  C-ABI call (``tsb_contract2 / tsb_rewire / tsb_permute / tsb_slice / tsb_union / tsb_stack / tsb_spider /
  tsb_all_ones``, include/tsb.h); this module adds no arithmetic.

Importing this module needs the reference package (``tensorsat``) on the path; it is not imported by the rest of
``tensorsat_b200`` (the GPU boxes do not have the reference).
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any, Self, final

import numpy as np
from tensorsat.diagrams import Port, TensorLikeBox, Type, Wire  # the REAL reference classes
from tensorsat.lang.fin_rel import FinRel, FinSet

from ..lang.fin_rel import CudaFinRel as _DeviceRel


@final
class CudaFinRel(TensorLikeBox):
    """Device-resident drop-in for ``tensorsat.lang.fin_rel.FinRel`` (boolean semiring, 0/1 ``uint8``)."""

    @classmethod
    def _new(cls, impl: _DeviceRel, name: str | None = None) -> Self:
        self = super().__new__(cls, name=name)
        self.__impl = impl
        return self

    __impl: _DeviceRel

    def __new__(cls, tensor: Any) -> Self:
        """From a 0/1 ``uint8`` numpy array (validated like ``FinRel.__new__``, fin_rel.py:328-340) or a CUDA
        ``torch.Tensor``."""
        return cls._new(_DeviceRel(tensor))

    @classmethod
    def from_fin_rel(cls, rel: FinRel) -> Self:
        """Uploads a reference relation."""
        return cls._new(_DeviceRel(np.ascontiguousarray(rel.tensor)), name=rel.name)

    def to_fin_rel(self) -> FinRel:
        """Downloads into a reference relation (device -> host copy)."""
        return FinRel(np.array(self.__impl.tensor, dtype=np.uint8))

    @property
    def tensor(self) -> np.ndarray:
        """Read-only host copy of the tensor, like ``FinRel.tensor`` (fin_rel.py:348-351)."""
        return self.__impl.tensor

    @property
    def device_tensor(self) -> Any:
        return self.__impl.device_tensor

    @property
    def shape(self) -> tuple[FinSet, ...]:
        return tuple(FinSet._new(d) for d in self.__impl.tensor_shape)

    def to_set(self) -> frozenset[tuple[int, ...]]:
        return self.__impl.to_set()

    # ---- protected hooks (tensorsat/diagrams.py:789-996): one device call each --------------------------
    @classmethod
    def _spider(cls, t: Type, num_ports: int) -> Self:
        assert isinstance(t, FinSet)
        from ..diagrams import FinSet as _MirrorFinSet

        return cls._new(_DeviceRel._spider(_MirrorFinSet(t.tensor_dim), num_ports))

    @classmethod
    def _scalar(cls, scalar: bool, /) -> Self:
        return cls._new(_DeviceRel._scalar(scalar))

    @classmethod
    def _contract2(
        cls,
        lhs: Self,
        lhs_wires: Sequence[Wire],
        rhs: Self,
        rhs_wires: Sequence[Wire],
        out_wires: Sequence[Wire],
    ) -> Self:
        return cls._new(_DeviceRel._contract2(lhs.__impl, list(lhs_wires), rhs.__impl, list(rhs_wires), list(out_wires)))

    @classmethod
    def _stack(cls, boxes: tuple[Self, ...], port: int = 0, /) -> Self:
        return cls._new(_DeviceRel._stack(tuple(b.__impl for b in boxes), port))

    def _rewire(self, out_ports: Sequence[Port]) -> Self:
        return CudaFinRel._new(self.__impl._rewire(list(out_ports)))

    def _getitem(self, idxs: tuple[int | slice, ...]) -> Self:
        return CudaFinRel._new(self.__impl._getitem(tuple(idxs)))

    def _union(self, other: Self) -> Self:
        return CudaFinRel._new(self.__impl._union(other.__impl))

    def _as_scalar(self) -> bool:
        return self.__impl._as_scalar()

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, CudaFinRel):
            return NotImplemented
        return self.__impl == other.__impl

    def __hash__(self) -> int:
        return hash(self.__impl)


def to_device(diagram: Any) -> Any:
    """A flat reference ``Diagram`` with every ``FinRel`` box replaced by its device twin (same wiring)."""
    from tensorsat.diagrams import Diagram

    flat = diagram.flatten() if not diagram.is_flat else diagram
    return Diagram._new(flat.wiring, tuple(CudaFinRel.from_fin_rel(b) for b in flat.boxes))

"""
Host-side mirror of the slice of ``tensorsat.diagrams`` that the contraction hot path
touches: the flat-wiring triple that is the executor's input format, a small diagram
builder so that the benchmark networks can be stated the way the reference states them,
and the ``Box`` plug-in protocol that a box class must implement to be contracted.

Reference anchors (paths relative to the reference repo):

* ``Wiring`` data = ``(slot_wires_list, out_wires, wire_types)``, tensorsat/diagrams.py:153-235;
  ``dangling_wires`` :291-295; ``compose`` (wire fusion + renumbering) :387-461.
* ``DiagramBuilder.add_block`` wire/slot creation order :1758-1771, ``add_inputs`` :1790-1795,
  ``box @ builder[wires]`` :1926-1948.
* ``Diagram.flatten`` :1452-1481, ``>>`` sequential composition :1592-1637.
* ``Box`` public wrappers and the protected hooks behind them :693-1038
  (``contract2`` :811-866, ``rewire`` :922-939, ``spider`` :777-795, ``scalar`` :797-809,
  ``prod`` :719-746, ``union`` :868-881, ``stack`` :883-901, ``__bool__`` :949-958,
  ``__add__`` :965-970, ``__getitem__`` :972-992), ``TensorLikeBox.tensor_shape`` :1013-1031.

This is bookkeeping only (O(#wires) Python).  No tensor arithmetic happens here; the only
arithmetic implementation in this package is the CUDA one behind :class:`CudaFinRel`.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Iterable, Mapping, Sequence
from typing import Any, ClassVar

Wire = int
Port = int
Slot = int
Wires = tuple[int, ...]


class Type:
    """Base class for wire/port types (tensorsat/diagrams.py:59-76)."""


class TensorLikeType(Type):
    """Types with an integral tensor dimension (tensorsat/diagrams.py:78-88)."""

    @property
    def tensor_dim(self) -> int:  # pragma: no cover - abstract
        raise NotImplementedError


class FinSet(TensorLikeType):
    """Finite set ``{0..size-1}``, interned by size (tensorsat/lang/fin_rel.py:75-118)."""

    _store: ClassVar[dict[int, "FinSet"]] = {}
    __slots__ = ("_size",)

    def __new__(cls, size: int) -> "FinSet":
        if not isinstance(size, int) or isinstance(size, bool):
            raise TypeError("Finite set size must be an integer.")
        if size <= 0:
            raise ValueError("Finite set size must be strictly positive.")
        self = cls._store.get(size)
        if self is None:
            self = super().__new__(cls)
            self._size = size
            cls._store[size] = self
        return self

    @property
    def size(self) -> int:
        return self._size

    @property
    def tensor_dim(self) -> int:
        return self._size

    def __repr__(self) -> str:
        return f"FinSet({self._size})"


Shape = tuple[Type, ...]


class Shaped:
    @property
    def shape(self) -> Shape:  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def num_ports(self) -> int:
        return len(self.shape)

    @property
    def ports(self) -> Sequence[Port]:
        return range(self.num_ports)


class Wiring(Shaped):
    """Immutable wiring: which wire each slot port / out port is attached to."""

    __slots__ = ("_wire_types", "_slot_wires_list", "_out_wires")

    def __init__(
        self,
        *,
        wire_types: Sequence[Type],
        slot_wires_list: Sequence[Sequence[Wire]],
        out_wires: Sequence[Wire],
    ) -> None:
        self._wire_types = tuple(wire_types)
        self._slot_wires_list = tuple(tuple(int(w) for w in ws) for ws in slot_wires_list)
        self._out_wires = tuple(int(w) for w in out_wires)
        n = len(self._wire_types)
        for ws in self._slot_wires_list + (self._out_wires,):
            for w in ws:
                if not 0 <= w < n:
                    raise ValueError(f"Invalid wire index {w}.")

    @property
    def wire_types(self) -> Shape:
        return self._wire_types

    @property
    def slot_wires_list(self) -> tuple[Wires, ...]:
        return self._slot_wires_list

    @property
    def out_wires(self) -> Wires:
        return self._out_wires

    @property
    def slot_shapes(self) -> tuple[Shape, ...]:
        t = self._wire_types
        return tuple(tuple(t[w] for w in ws) for ws in self._slot_wires_list)

    @property
    def shape(self) -> Shape:
        return tuple(self._wire_types[w] for w in self._out_wires)

    @property
    def num_slots(self) -> int:
        return len(self._slot_wires_list)

    @property
    def slots(self) -> Sequence[Slot]:
        return range(self.num_slots)

    @property
    def num_wires(self) -> int:
        return len(self._wire_types)

    @property
    def wires(self) -> Sequence[Wire]:
        return range(self.num_wires)

    @property
    def dangling_wires(self) -> frozenset[Wire]:
        return frozenset(self.wires) - frozenset(w for ws in self._slot_wires_list for w in ws)

    @property
    def wire_dims(self) -> tuple[int, ...]:
        return tuple(t.tensor_dim for t in self._wire_types)  # type: ignore[attr-defined]

    def compose(self, wirings: Mapping[Slot, "Wiring"]) -> "Wiring":
        """Plug wirings into slots, fusing wires (tensorsat/diagrams.py:387-461).

        Numbering convention kept from the reference: surviving outer wires keep their
        relative order (a fused group is represented by its lowest outer wire), then the
        internal wires of each plugged wiring follow in slot order.
        """
        parent: dict[Any, Any] = {}

        def find(x: Any) -> Any:
            root = x
            while parent.setdefault(root, root) != root:
                root = parent[root]
            while parent[x] != root:
                parent[x], x = root, parent[x]
            return root

        def union(a: Any, b: Any) -> None:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[rb] = ra

        for slot, sub in wirings.items():
            if len(sub.out_wires) != len(self._slot_wires_list[slot]):
                raise ValueError(f"Wiring for slot {slot} has the wrong number of ports.")
            for w, sw in zip(self._slot_wires_list[slot], sub.out_wires):
                union(("o", w), ("s", slot, sw))
        # representative outer wire of each fused group = lowest outer wire in it
        group_min: dict[Any, int] = {}
        for w in self.wires:
            r = find(("o", w))
            if r not in group_min:
                group_min[r] = w
        new_types: list[Type] = []
        remap: dict[Any, int] = {}
        for w in self.wires:
            r = find(("o", w))
            if group_min[r] == w:
                remap[r] = len(new_types)
                new_types.append(self._wire_types[w])
        for slot in sorted(wirings):
            sub = wirings[slot]
            for sw in sub.wires:
                r = find(("s", slot, sw))
                if r not in remap:
                    remap[r] = len(new_types)
                    new_types.append(sub.wire_types[sw])
        new_slots: list[Wires] = []
        for slot in self.slots:
            if slot in wirings:
                for ws in wirings[slot].slot_wires_list:
                    new_slots.append(tuple(remap[find(("s", slot, w))] for w in ws))
            else:
                new_slots.append(tuple(remap[find(("o", w))] for w in self._slot_wires_list[slot]))
        out = tuple(remap[find(("o", w))] for w in self._out_wires)
        return Wiring(wire_types=new_types, slot_wires_list=new_slots, out_wires=out)

    def _key(self) -> tuple[Any, ...]:
        return (self._wire_types, self._slot_wires_list, self._out_wires)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, Wiring) and self._key() == other._key()

    def __hash__(self) -> int:
        return hash(self._key())

    def __repr__(self) -> str:
        return (
            f"<Wiring: {self.num_wires} wires, {self.num_slots} slots, "
            f"{len(self._out_wires)} out ports>"
        )


class Diagram(Shaped):
    """A wiring with a block (box or sub-diagram) in every slot."""

    def __init__(self, wiring: Wiring, blocks: Sequence["Box | Diagram"]) -> None:
        blocks = tuple(blocks)
        if len(blocks) != wiring.num_slots:
            raise ValueError("Exactly one block per slot is required.")
        for slot, block in enumerate(blocks):
            if tuple(block.shape) != wiring.slot_shapes[slot]:
                raise ValueError(f"Block shape does not match slot {slot}.")
        self._wiring = wiring
        self._blocks = blocks

    @staticmethod
    def from_recipe(recipe):  # type: ignore[no-untyped-def]
        """``@Diagram.from_recipe`` (tensorsat/diagrams.py:1224-1254)."""
        builder = DiagramBuilder()
        recipe(builder)
        return builder.diagram()

    @staticmethod
    def recipe(recipe):  # type: ignore[no-untyped-def]
        """``@Diagram.recipe`` with per-argument caching (tensorsat/diagrams.py:1256-1287)."""
        cache: dict[Any, Diagram] = {}

        def make(*args: Any, **kwargs: Any) -> Diagram:
            key = (args, frozenset(kwargs.items()))
            if key not in cache:
                builder = DiagramBuilder()
                recipe(builder, *args, **kwargs)
                cache[key] = builder.diagram()
            return cache[key]

        make.__name__ = getattr(recipe, "__name__", "recipe")
        make.__doc__ = recipe.__doc__
        return make

    @property
    def wiring(self) -> Wiring:
        return self._wiring

    @property
    def blocks(self) -> tuple["Box | Diagram", ...]:
        return self._blocks

    @property
    def shape(self) -> Shape:
        return self._wiring.shape

    @property
    def num_open_slots(self) -> int:
        return 0

    @property
    def subdiagram_slots(self) -> tuple[Slot, ...]:
        return tuple(s for s, b in enumerate(self._blocks) if isinstance(b, Diagram))

    @property
    def boxes(self) -> tuple["Box", ...]:
        return tuple(b for b in self._blocks if isinstance(b, Box))

    @property
    def is_flat(self) -> bool:
        return not self.subdiagram_slots

    @property
    def box_classes(self) -> frozenset[type]:
        return frozenset(type(b) for b in self.boxes)

    def flatten(self) -> "Diagram":
        if self.is_flat:
            return self
        subs = {s: self._blocks[s].flatten() for s in self.subdiagram_slots}  # type: ignore[union-attr]
        wiring = self._wiring.compose({s: d.wiring for s, d in subs.items()})
        blocks: list[Box | Diagram] = []
        for slot, block in enumerate(self._blocks):
            blocks.extend(subs[slot].blocks if slot in subs else (block,))
        return Diagram(wiring, blocks)

    def __rshift__(self, other: "Box | Diagram") -> "Diagram":
        """All ports of ``self`` feed the first ports of ``other`` (diagrams.py:1487-1493,1592-1637)."""
        if not isinstance(other, (Box, Diagram)):
            return NotImplemented
        if other.num_ports < self.num_ports:
            raise ValueError(
                f"Cannot sequentially compose: {self.num_ports} selected on LHS,"
                f" {other.num_ports} ports available on RHS."
            )
        builder = DiagramBuilder()
        wires: Wires = ()
        for block in (self, other):
            wires = block @ builder[wires]
        builder.add_outputs(wires)
        return builder.diagram()

    def __repr__(self) -> str:
        attrs = [f"{self._wiring.num_wires} wires", f"{len(self._blocks)} blocks"]
        if self.num_ports:
            attrs.append(f"{self.num_ports} ports")
        return f"<Diagram: {', '.join(attrs)}>"


class DiagramBuilder:
    """Incremental diagram construction (tensorsat/diagrams.py:1647-1869)."""

    def __init__(self) -> None:
        self._wire_types: list[Type] = []
        self._slot_wires: list[Wires] = []
        self._out_wires: list[Wire] = []
        self._blocks: list[Box | Diagram] = []

    @property
    def num_wires(self) -> int:
        return len(self._wire_types)

    @property
    def wires(self) -> Sequence[Wire]:
        return range(len(self._wire_types))

    def add_wire(self, t: Type) -> Wire:
        self._wire_types.append(t)
        return len(self._wire_types) - 1

    def add_input(self, t: Type) -> Wire:
        return self.add_inputs((t,))[0]

    def add_inputs(self, ts: Iterable[Type]) -> Wires:
        wires = tuple(self.add_wire(t) for t in ts)
        self._out_wires.extend(wires)
        return wires

    def add_output(self, wire: Wire) -> Port:
        return self.add_outputs((wire,))[0]

    def add_outputs(self, wires: Iterable[Wire]) -> tuple[Port, ...]:
        wires = tuple(wires)
        for w in wires:
            if not 0 <= w < len(self._wire_types):
                raise ValueError(f"Invalid wire index {w}.")
        start = len(self._out_wires)
        self._out_wires.extend(wires)
        return tuple(range(start, start + len(wires)))

    def add_block(self, block: "Box | Diagram", inputs: Mapping[Port, Wire] | None = None) -> Wires:
        inputs = dict(inputs or {})
        shape = block.shape
        for port, wire in inputs.items():
            if not 0 <= port < len(shape):
                raise ValueError(f"Invalid port {port} for block.")
            if not 0 <= wire < len(self._wire_types):
                raise ValueError(f"Invalid wire index {wire}.")
            if shape[port] != self._wire_types[wire]:
                raise ValueError(f"Incompatible wire type for port {port}.")
        out_ports = tuple(p for p in range(len(shape)) if p not in inputs)
        out_wires = tuple(self.add_wire(shape[p]) for p in out_ports)
        mapping = {**inputs, **dict(zip(out_ports, out_wires))}
        self._slot_wires.append(tuple(mapping[p] for p in range(len(shape))))
        self._blocks.append(block)
        return out_wires

    def __getitem__(self, wires: Wire | Sequence[Wire]) -> "SelectedBuilderWires":
        if isinstance(wires, int):
            wires = (wires,)
        return SelectedBuilderWires(self, tuple(wires))

    def __rmatmul__(self, block: "Box | Diagram") -> Wires:
        return self.add_block(block)

    def diagram(self) -> Diagram:
        wiring = Wiring(
            wire_types=self._wire_types,
            slot_wires_list=self._slot_wires,
            out_wires=self._out_wires,
        )
        return Diagram(wiring, self._blocks)


class SelectedBuilderWires:
    def __init__(self, builder: DiagramBuilder, wires: Wires) -> None:
        for w in wires:
            if not 0 <= w < builder.num_wires:
                raise ValueError(f"Invalid wire index {w}.")
        self.builder = builder
        self.wires = wires

    def __rmatmul__(self, other: "Box | Diagram") -> Wires:
        if not isinstance(other, (Box, Diagram)):
            return NotImplemented
        if other.num_ports < len(self.wires):
            raise ValueError(
                f"Cannot apply block: {len(self.wires)} wires selected,"
                f" {other.num_ports} ports available on RHS."
            )
        return self.builder.add_block(other, dict(enumerate(self.wires)))


class Box(Shaped, ABC):
    """
    The plug-in protocol of a contractible box class (tensorsat/diagrams.py:693-1010).

    Public methods validate and then call the protected hooks; a concrete box class
    supplies the hooks.  Error behaviour follows the reference: ``ValueError`` for
    contract violations, ``TypeError`` for a foreign box class.
    """

    _name: str | None = None

    @property
    def name(self) -> str | None:
        return self._name

    @staticmethod
    def class_join(bases: Iterable[type]) -> type:
        """Most specific common ancestor under ``Box`` (tensorsat/diagrams.py:698-703)."""
        bases = list(bases)
        if not bases:
            raise ValueError("Need at least one box class.")
        for cand in bases[0].__mro__:
            if isinstance(cand, type) and issubclass(cand, Box) and all(issubclass(b, cand) for b in bases):
                return cand
        return Box

    @classmethod
    def can_be_contracted(cls) -> bool:
        abstract = getattr(cls, "__abstractmethods__", frozenset())
        return not ({"_contract2", "_rewire", "_spider"} & set(abstract))

    @classmethod
    def prod(cls, boxes: Iterable["Box"]) -> "Box":
        boxes = tuple(boxes)
        for box in boxes:
            if not isinstance(box, cls):
                raise TypeError(
                    f"Boxes must be instances of box class {cls},"
                    f" found instance of box class {type(box)} instead."
                )
        if not boxes:
            raise ValueError("Products must involve at least one box.")
        res = boxes[0]
        for box in boxes[1:]:
            nl, nr = len(res.shape), len(box.shape)
            res = cls._contract2(res, range(nl), box, range(nl, nl + nr), range(nl + nr))
        return res

    @classmethod
    def from_wiring(cls, wiring: "Wiring") -> "Box":
        """The box of this class that a wiring denotes on its own (tensorsat/diagrams.py:750-775): one spider per
        wire over the ports attached to it, their product, ports ordered as the slot ports in slot order followed
        by the out ports."""
        attached: dict[Wire, list[tuple[int, int]]] = {}   # wire -> (slot or -1 for an out port, port), in that order
        for slot, ws in enumerate(wiring.slot_wires_list):
            for port, w in enumerate(ws):
                attached.setdefault(w, []).append((slot, port))
        for port, w in enumerate(wiring.out_wires):
            attached.setdefault(w, []).append((-1, port))
        wired = [w for w in wiring.wires if w in attached]
        types = wiring.wire_types
        res = cls.prod(cls.spider(types[w], len(attached[w])) for w in wired)
        have = [label for w in wired for label in attached[w]]
        want = [(slot, port) for slot, ws in enumerate(wiring.slot_wires_list) for port in range(len(ws))]
        want += [(-1, port) for port in range(len(wiring.out_wires))]
        return res._rewire([have.index(label) for label in want])

    @classmethod
    def spider(cls, t: Type, num_ports: int) -> "Box":
        if not isinstance(num_ports, int):
            raise TypeError("Number of ports must be an integer.")
        if num_ports <= 0:
            raise ValueError("Number of ports must be strictly positive.")
        return cls._spider(t, num_ports)

    @classmethod
    @abstractmethod
    def _spider(cls, t: Type, num_ports: int) -> "Box": ...

    @classmethod
    def scalar(cls, scalar: bool, /) -> "Box":
        box = cls._scalar(scalar)
        assert box.shape == ()
        return box

    @classmethod
    @abstractmethod
    def _scalar(cls, scalar: bool, /) -> "Box": ...

    @classmethod
    def contract2(
        cls,
        lhs: "Box",
        lhs_wires: Sequence[Wire],
        rhs: "Box",
        rhs_wires: Sequence[Wire],
        out_wires: Sequence[Wire] | None = None,
    ) -> "Box":
        if not isinstance(lhs, cls) or not isinstance(rhs, cls):
            raise TypeError(f"Operands must be instances of box class {cls}.")
        if len(lhs_wires) != len(lhs.shape):
            raise ValueError(
                f"Number of wires in lhs ({len(lhs_wires)}) does not match"
                f" the number of ports in lhs shape ({len(lhs.shape)})."
            )
        if len(rhs_wires) != len(rhs.shape):
            raise ValueError(
                f"Number of wires in rhs ({len(rhs_wires)}) does not match"
                f" the number of ports in rhs shape ({len(rhs.shape)})."
            )
        if out_wires is None:
            out_wires = sorted(set(lhs_wires).symmetric_difference(rhs_wires))
        else:
            extra = set(out_wires) - set(lhs_wires) - set(rhs_wires)
            if extra:
                raise ValueError("Every output wire must appear in LHR or RHS.")
        return cls._contract2(lhs, lhs_wires, rhs, rhs_wires, out_wires)

    @classmethod
    @abstractmethod
    def _contract2(
        cls,
        lhs: "Box",
        lhs_wires: Sequence[Wire],
        rhs: "Box",
        rhs_wires: Sequence[Wire],
        out_wires: Sequence[Wire],
    ) -> "Box": ...

    @classmethod
    def union(cls, boxes: Iterable["Box"]) -> "Box":
        boxes = tuple(boxes)
        if not boxes:
            raise ValueError("Cannot take an empty union: no shape info available.")
        res = boxes[0]
        for box in boxes[1:]:
            if box.shape != res.shape:
                raise ValueError("Shapes of boxes in a union must coincide.")
            res = res._union(box)
        return res

    @classmethod
    def stack(cls, boxes: Iterable["Box"], port: int = 0, /) -> "Box":
        boxes = tuple(boxes)
        if not boxes:
            raise ValueError("Need at least one box to stack.")
        shape = boxes[0].shape
        if port not in range(len(shape) + 1):
            raise ValueError(f"Invalid port index {port} for stacking.")
        box = cls._stack(boxes, port)
        assert box.shape == shape[:port] + (FinSet(len(boxes)),) + shape[port:]
        return box

    @classmethod
    @abstractmethod
    def _stack(cls, boxes: tuple["Box", ...], port: int = 0, /) -> "Box": ...

    def rewire(self, out_ports: Sequence[Port]) -> "Box":
        ports = self.ports
        if not all(port in ports for port in out_ports):
            raise ValueError("Output ports must be valid for this box.")
        return self._rewire(out_ports)

    @abstractmethod
    def _rewire(self, out_ports: Sequence[Port]) -> "Box": ...

    @abstractmethod
    def _as_scalar(self) -> bool: ...

    def __bool__(self) -> bool:
        if self.shape:
            raise ValueError("Box is not a scalar.")
        return self._as_scalar()

    def __add__(self, other: "Box") -> "Box":
        if self.shape != other.shape:
            raise ValueError("Boxes in a union must have the same shape")
        return self._union(other)

    def __getitem__(self, idxs: tuple[int | slice, ...]) -> "Box":
        if not isinstance(idxs, tuple):
            raise TypeError("Box indices must be a tuple.")
        shape = self.shape
        for port, idx in enumerate(idxs):
            if isinstance(idx, int):
                if not isinstance(shape[port], FinSet):
                    raise ValueError("Port indexing is only allowed for finite enumerated sets.")
            elif not isinstance(idx, slice):
                raise TypeError("Box indices must be integers or slices.")
            elif idx != slice(None, None, None):
                raise NotImplementedError("Non-trivial slicing is not implemented.")
        return self._getitem(idxs)

    @abstractmethod
    def _getitem(self, idxs: tuple[int | slice, ...]) -> "Box": ...

    @abstractmethod
    def _union(self, other: "Box") -> "Box": ...

    def __repr__(self) -> str:
        n = len(self.shape)
        if self.name is None:
            return f"<{type(self).__name__}: {n} ports>"
        return f"<{type(self).__name__}: {n} ports, named {self.name!r}>"


class TensorLikeBox(Box):
    """Boxes with an integral tensor shape (tensorsat/diagrams.py:1013-1031)."""

    @property
    def tensor_shape(self) -> tuple[int, ...]:
        return tuple(t.tensor_dim for t in self.shape)  # type: ignore[attr-defined]

"""
Flat network = the executor's input format: one tuple of wire ids per leaf tensor, the
output wires and the wire dimensions — exactly the triple the reference hands to
``cotengra.array_contract_tree`` (tensorsat/contractions/cotengra.py:78-82).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from ..diagrams import Diagram, Wiring


@dataclass(frozen=True)
class Network:
    inputs: tuple[tuple[int, ...], ...]
    output: tuple[int, ...]
    size_dict: tuple[int, ...]  # indexed by wire id

    @classmethod
    def from_wiring(cls, wiring: Wiring) -> "Network":
        return cls(
            tuple(wiring.slot_wires_list),
            tuple(wiring.out_wires),
            tuple(wiring.wire_dims),
        )

    @classmethod
    def from_diagram(cls, diagram: Diagram) -> "Network":
        return cls.from_wiring(diagram.flatten().wiring)

    @property
    def num_leaves(self) -> int:
        return len(self.inputs)

    @property
    def num_wires(self) -> int:
        return len(self.size_dict)

    def to_json(self) -> dict[str, Any]:
        return {
            "inputs": [list(ws) for ws in self.inputs],
            "output": list(self.output),
            "size_dict": list(self.size_dict),
        }

    @classmethod
    def from_json(cls, d: dict[str, Any]) -> "Network":
        return cls(
            tuple(tuple(ws) for ws in d["inputs"]),
            tuple(d["output"]),
            tuple(d["size_dict"]),
        )

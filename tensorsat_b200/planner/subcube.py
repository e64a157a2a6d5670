"""
Sampled sub-cubes of the wide BASELINE configurations as self-contained plan files.

Random 3-SAT n=120 / n=200 and the 16x16 Wallace multiplier need 2^53 / 2^106 / 2^88 slices of 2^30-entry
intermediates: no dense method contracts them exactly.  What is measured (and checked) instead is a SUB-CUBE of
the assignment space: all but ``keep`` of the tree's sliced wires are pinned on the host, the remaining
``2**keep`` slices run exactly like any other plan.  The pins follow a known solution of the instance (a
committed witness: a model of the CNF found offline, the factors 65521 x 65519 for the multiplier), so the
sub-cube's exact answer is NOT trivially zero: SAT = 1, and its model count is checked against an independent
exact counter (oracle/sharp_sat.py).

A plan file (``bench_plans/subcube_*.json``, written by tools/make_subcube_plans.py) holds the flat network, the
leaf tensors, the SSA path, the full list of sliced wires and the witness value of each: building the Python
diagram of the n=200 instance takes half a minute, loading the file milliseconds.
"""

from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .network import Network
from .sampling import sample_subcube
from .tree import ContractionTree

PLAN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "bench_plans")


@dataclass
class SubcubeFile:
    config: str
    description: str
    mode: str                      # "sat" | "count"
    n_vars: int
    tree: ContractionTree          # full slicing (all sliced wires)
    leaves: list[np.ndarray]
    witness: dict[int, int]        # sliced wire -> value in the known solution
    extra: dict[str, Any] = field(default_factory=dict)

    @property
    def wire_is_free(self) -> list[bool]:
        return [w < self.n_vars for w in range(self.tree.network.num_wires)]

    def subcube(self, keep: int, *, leaves: list[np.ndarray] | None = None, witness: bool = True, seed: int = 0):
        """``(subtree, subleaves, assignment)`` with ``2**keep`` executable slices; pins follow the witness (or
        seeded random values with ``witness=False``: an almost surely empty sub-cube)."""
        return sample_subcube(self.tree, self.leaves if leaves is None else leaves, keep, seed=seed,
                              witness=self.witness if witness else None)

    def to_json(self) -> dict[str, Any]:
        return {
            "config": self.config, "description": self.description, "mode": self.mode, "n_vars": self.n_vars,
            "tree": self.tree.to_json(),
            "leaves": [[list(t.shape), [int(x) for x in t.reshape(-1)]] for t in self.leaves],
            "witness": {str(k): int(v) for k, v in self.witness.items()},
            "extra": self.extra,
        }

    @classmethod
    def from_json(cls, d: dict[str, Any]) -> "SubcubeFile":
        leaves = [np.array(v, dtype=np.uint8).reshape(shape) for shape, v in d["leaves"]]
        return cls(d["config"], d["description"], d["mode"], int(d["n_vars"]), ContractionTree.from_json(d["tree"]),
                   leaves, {int(k): int(v) for k, v in d["witness"].items()}, d.get("extra", {}))


def path_of(config: str) -> str:
    return os.path.join(PLAN_DIR, f"subcube_{config}.json")


def load(config: str) -> SubcubeFile:
    with open(path_of(config)) as f:
        return SubcubeFile.from_json(json.load(f))


def decode_leaves(extra_leaves: list[Any]) -> list[np.ndarray]:
    return [np.array(v, dtype=np.uint8).reshape(shape) for shape, v in extra_leaves]

"""Host-side planning: flat networks, contraction trees, slicing (not on the timed path)."""

from .network import Network
from .sampling import fix_wires, propagate_wire_values, sample_subcube
from .search import find_tree
from .tree import ContractionTree, TreeStats, linear_to_ssa, ssa_to_linear

__all__ = [
    "Network",
    "ContractionTree",
    "TreeStats",
    "find_tree",
    "fix_wires",
    "sample_subcube",
    "propagate_wire_values",
    "linear_to_ssa",
    "ssa_to_linear",
]

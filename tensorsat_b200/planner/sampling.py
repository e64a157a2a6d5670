"""
Sampled slices of contractions that are too wide to finish.

Random 3-SAT at ratio 4.26 has contraction width growing linearly in n; BASELINE configs 2 (n=120),
4 (n=200) and 5 (16x16 Wallace multiplier) need hundreds of sliced wires to fit a GPU, i.e. more slices
than can ever be executed.  What CAN be measured honestly is the cost of one slice (all slices of a tree
have identical shapes and cost).  ``fix_wires`` pins most sliced wires to sampled constants on the host
(indexing 2..8-byte leaf tensors), leaving a small network with a few sliced wires that the executor runs
exactly like any other; the result is the exact answer for that sub-cube of the assignment space only.
"""

from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np

from .network import Network
from .tree import ContractionTree


def fix_wires(net: Network, leaves: Sequence[np.ndarray], assignment: Mapping[int, int]) -> tuple[Network, list[np.ndarray]]:
    """Index every leaf at the assigned value of each fixed wire; the wires disappear from the network."""
    for w in assignment:
        if w in net.output:
            raise ValueError(f"Wire {w} is an output wire and cannot be fixed.")
    new_inputs, new_leaves = [], []
    for ws, t in zip(net.inputs, leaves):
        idx = tuple(int(assignment[w]) if w in assignment else slice(None) for w in ws)
        new_leaves.append(np.array(np.asarray(t)[idx], order="C"))
        new_inputs.append(tuple(w for w in ws if w not in assignment))
    return Network(tuple(new_inputs), net.output, net.size_dict), new_leaves


def sample_subcube(tree: ContractionTree, leaves: Sequence[np.ndarray], keep_sliced: int, seed: int = 0):
    """Fix all but the last ``keep_sliced`` sliced wires of ``tree`` to seeded random values.

    Returns ``(subtree, subleaves, assignment)``: the same SSA path over the reduced network with
    ``keep_sliced`` sliced wires left (``2**keep_sliced`` executable slices for binary wires)."""
    rng = np.random.default_rng(seed)
    sliced = list(tree.sliced_wires)
    fixed, kept = sliced[: max(0, len(sliced) - keep_sliced)], sliced[max(0, len(sliced) - keep_sliced):]
    assignment = {w: int(rng.integers(0, tree.network.size_dict[w])) for w in fixed}
    net, sub = fix_wires(tree.network, leaves, assignment)
    return ContractionTree(net, tree.ssa_path, kept), sub, assignment

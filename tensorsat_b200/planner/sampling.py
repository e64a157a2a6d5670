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


def propagate_wire_values(net: Network, leaves: Sequence[np.ndarray], known: Mapping[int, int]) -> dict[int, int]:
    """Values of every wire that the relation tensors DETERMINE once the ``known`` wires (e.g. the variables of
    a CNF, or the input bits of a circuit) are set: a leaf with one open wire left and exactly one value of it
    that keeps the leaf's entry at 1 fixes that wire (gates are function graphs: or_, not_, xor_, ...).  Raises
    ``ValueError`` if a leaf becomes unsatisfiable (the known values are not part of any solution).  Wires that
    stay undetermined are absent from the result."""
    val = {int(w): int(v) for w, v in known.items()}
    pending = list(range(len(net.inputs)))
    while pending:
        nxt, progress = [], False
        for i in pending:
            ws, t = net.inputs[i], np.asarray(leaves[i])
            open_ax = [a for a, w in enumerate(ws) if w not in val]
            idx = tuple(val[w] if w in val else slice(None) for w in ws)
            sub = t[idx]
            if not open_ax:
                if not sub:
                    raise ValueError(f"leaf {i} is violated by the given wire values")
                progress = True
                continue
            if len({ws[a] for a in open_ax}) == 1:
                # one open wire (possibly on several ports: take the diagonal)
                k = len(open_ax)
                diag = np.array([sub[(v,) * k] for v in range(sub.shape[0])])
                ok = np.flatnonzero(diag)
                if len(ok) == 0:
                    raise ValueError(f"leaf {i} is violated by the given wire values")
                if len(ok) == 1:
                    val[ws[open_ax[0]]] = int(ok[0])
                    progress = True
                    continue
            nxt.append(i)
        if not progress:
            break
        pending = nxt
    return val


def sample_subcube(tree: ContractionTree, leaves: Sequence[np.ndarray], keep_sliced: int, seed: int = 0,
                   witness: Mapping[int, int] | None = None):
    """Fix all but the last ``keep_sliced`` sliced wires of ``tree``: to the values of ``witness`` (wire -> value
    of a known solution, see ``propagate_wire_values``: the sub-cube then CONTAINS that solution, so its exact
    answer is SAT / a count >= 1) where given, else to seeded random values (the sub-cube is then almost surely
    empty for a wide instance).

    Returns ``(subtree, subleaves, assignment)``: the same SSA path over the reduced network with
    ``keep_sliced`` sliced wires left (``2**keep_sliced`` executable slices for binary wires)."""
    rng = np.random.default_rng(seed)
    sliced = list(tree.sliced_wires)
    fixed, kept = sliced[: max(0, len(sliced) - keep_sliced)], sliced[max(0, len(sliced) - keep_sliced):]
    assignment = {}
    for w in fixed:
        r = int(rng.integers(0, tree.network.size_dict[w]))   # drawn either way: same stream with and without a witness
        assignment[w] = int(witness[w]) if witness is not None and w in witness else r
    net, sub = fix_wires(tree.network, leaves, assignment)
    return ContractionTree(net, tree.ssa_path, kept), sub, assignment

"""
Slice-parallel execution across the GPUs of one box.

Fixing the sliced wires of a contraction tree to each value combination yields independent
contractions of the same tree; the reference combines them with ``Box.__add__`` (bitwise OR,
tensorsat/diagrams.py:965-970, tensorsat/lang/fin_rel.py:345-346).  Here every rank (one process
per GPU, ``torch.distributed``) contracts a contiguous block of slice ids with one
``tsb_plan_execute`` — accumulating locally on the device — and the partial results are combined
by ONE collective: ``all_reduce(MAX)`` on uint8 for the boolean semiring, ``all_reduce(SUM)`` on
int64 for the counting semiring.  The payload is the result tensor (1 byte / 8 bytes for a
decision / count), so the collective is latency-bound and NCCL over NVLink is used as is.
"""

from __future__ import annotations

from typing import Any

from .contractions.tree import TreeContraction
from .diagrams import Box, Diagram
from .lang.fin_rel import MODE_SAT


def slice_block(n_slices: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of slice ids owned by ``rank``: sizes differ by at most one."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("invalid rank / world size")
    base, rem = divmod(n_slices, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_result(out: Any, mode: int, group: Any = None) -> Any:
    """In-place combination of per-rank partial results (max for SAT, sum for #SAT)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        op = dist.ReduceOp.MAX if mode == MODE_SAT else dist.ReduceOp.SUM
        dist.all_reduce(out, op=op, group=group)
    return out


def contract_distributed(contraction: TreeContraction, diagram: Diagram, group: Any = None) -> Box:
    """Contracts ``diagram`` with the slices of ``contraction.tree`` split over the ranks of ``group``."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = slice_block(contraction.num_slices, rank, world)
    out = contraction.contract_raw(diagram, (lo, hi))
    all_reduce_result(out, contraction.mode, group)
    dims = tuple(contraction.wiring.wire_dims[w] for w in contraction._distinct_out)
    box = contraction.box_class._new(dims, dev=out)
    if contraction._distinct_out != tuple(contraction.wiring.out_wires):
        box = box._rewire([contraction._distinct_out.index(w) for w in contraction.wiring.out_wires])
    return box

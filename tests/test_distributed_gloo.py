"""world_size-2 gloo test of the slice-parallel path (host emulation build, CPU tensors)."""

import os
import sys

import pytest

from tests.conftest import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch, torch.distributed as dist
from tests.backend import bind_emulation
bind_emulation()
from oracle import fin_rel_oracle as O
from tensorsat_b200.contractions.tree import TreeContraction
from tensorsat_b200.distributed import contract_distributed, slice_block
from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel
from tensorsat_b200.lib.sat import CNFInstance
from tensorsat_b200.planner import Network
from tensorsat_b200.planner.search import find_tree_elimination

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
cnf = CNFInstance.random(3, 14, 58, rng=3)
d = (cnf.inputs(None) >> cnf.diagram()).flatten()
tree = find_tree_elimination(Network.from_diagram(d), trials=2, seed=0).slice_to(6)
assert tree.num_slices >= 4
brute = O.cnf_brute_force(14, cnf.clauses)
free = [w < 14 for w in d.wiring.wires]
sat = contract_distributed(TreeContraction(CudaFinRel, d.wiring, tree=tree), d)
cnt = contract_distributed(TreeContraction(CudaCountRel, d.wiring, tree=tree, wire_is_free=free), d)
assert bool(sat) == bool(brute.any()), (bool(sat), brute.sum())
assert int(cnt) == int(brute.sum()), (int(cnt), brute.sum())
# each rank really did only its own block: partial counts differ from the total in general
lo, hi = slice_block(tree.num_slices, dist.get_rank(), 2)
part = TreeContraction(CudaCountRel, d.wiring, tree=tree, wire_is_free=free).contract(d, slice_range=(lo, hi))
t = torch.tensor([int(part)])
dist.all_reduce(t)
assert int(t) == int(brute.sum())
# the end-to-end entry point on host leaves, combined on the device before the read-back (what bench.py's e2e times)
leaves = [b.tensor for b in d.boxes]
c = TreeContraction(CudaCountRel, d.wiring, tree=tree, wire_is_free=free)
for _ in range(2):
    assert int(c.contract_from_host(leaves, (lo, hi), all_reduce=True)) == int(brute.sum())
s = TreeContraction(CudaFinRel, d.wiring, tree=tree)
assert bool(s.contract_from_host(leaves, (lo, hi), all_reduce=True)) == bool(brute.any())
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
'''


def test_two_ranks_gloo(tmp_path):
    import socket
    import subprocess

    from tests.backend import bind_emulation

    bind_emulation()  # make sure the emulated library is built before the ranks start
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True, env={**os.environ, "MASTER_ADDR": "127.0.0.1"}) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o

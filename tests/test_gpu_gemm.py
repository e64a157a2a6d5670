"""tcgen05 int8 contraction kernel against exact integer references (GPU only)."""

import numpy as np
import pytest

from tests.backend import bind_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _bind():
    bind_cuda()


def _ref(A, B):
    import torch

    # exact: values <= 3, K <= 2^13  ->  sums < 2^24, representable in fp32; use fp64 anyway
    return torch.bmm(A.double(), B.double().transpose(1, 2)).round().to(torch.int64)


@pytest.mark.parametrize("b,m,n,k", [(0, 7, 5, 5), (0, 7, 8, 7), (1, 8, 6, 6), (2, 7, 9, 8), (0, 9, 9, 10), (3, 8, 8, 12),
                                     (0, 10, 11, 9), (0, 8, 5, 13)])
def test_gemm_u8_exact(b, m, n, k):
    import torch

    from tensorsat_b200 import _native

    g = torch.Generator(device="cuda").manual_seed(b * 1000 + m * 100 + n * 10 + k)
    Bt, M, N, K = 1 << b, 1 << m, 1 << n, 1 << k
    A = torch.randint(0, 4, (Bt, M, K), device="cuda", dtype=torch.uint8, generator=g)
    B = torch.randint(0, 4, (Bt, N, K), device="cuda", dtype=torch.uint8, generator=g)
    # sparse 0/1 operands so that the clamp is non-trivial
    A01 = (torch.rand((Bt, M, K), device="cuda", generator=g) < 2.0 / K).to(torch.uint8)
    B01 = (torch.rand((Bt, N, K), device="cuda", generator=g) < 0.5).to(torch.uint8)
    want = _ref(A, B)
    C32 = torch.empty((Bt, M, N), device="cuda", dtype=torch.int32)
    _native.gemm_u8(A, B, C32, b, m, n, k, _native.MODE_COUNT)
    torch.cuda.synchronize()
    assert torch.equal(C32.to(torch.int64), want)
    want01 = (_ref(A01, B01) != 0).to(torch.uint8)
    C8 = torch.empty((Bt, M, N), device="cuda", dtype=torch.uint8)
    _native.gemm_u8(A01, B01, C8, b, m, n, k, _native.MODE_SAT)
    torch.cuda.synchronize()
    assert torch.equal(C8, want01)
    assert 0 < int(want01.sum()) < want01.numel()


def test_gemm_rejects_small_shapes():
    import torch

    from tensorsat_b200 import _native

    A = torch.zeros(1 << 10, device="cuda", dtype=torch.uint8)
    with pytest.raises(NotImplementedError):
        _native.gemm_u8(A, A, A, 0, 5, 5, 5, 0)


@pytest.mark.parametrize("seed", range(4))
def test_contract2_above_the_ridge_uses_tensor_cores_and_matches_oracle(seed):
    """A pairwise step with enough arithmetic intensity is lowered to transpose -> tcgen05 GEMM -> transpose."""
    import numpy as np

    from oracle import fin_rel_oracle as O
    from tensorsat_b200.lang.fin_rel import CudaCountRel, CudaFinRel

    rng = np.random.default_rng(seed)
    m, n, k, b = 9, 8, 7, seed % 2
    wires = [int(w) for w in rng.permutation(m + n + k + b)]
    B, M, N, K = wires[:b], wires[b:b + m], wires[b + m:b + m + n], wires[b + m + n:]
    lw = [int(w) for w in rng.permutation(B + M + K)]
    rw = [int(w) for w in rng.permutation(B + N + K)]
    ow = [int(w) for w in rng.permutation(B + M + N)]
    lhs = (rng.random((2,) * len(lw)) < 0.02).astype(np.uint8)
    rhs = (rng.random((2,) * len(rw)) < 0.3).astype(np.uint8)
    got = CudaFinRel.contract2(CudaFinRel(lhs), lw, CudaFinRel(rhs), rw, ow)
    want = O.contract2(lhs, lw, rhs, rw, ow)
    assert np.array_equal(got.tensor, want) and 0 < want.sum() < want.size
    gotc = CudaCountRel.contract2(CudaCountRel(lhs), lw, CudaCountRel(rhs), rw, ow)
    assert np.array_equal(gotc.tensor, O.contract2(lhs, lw, rhs, rw, ow, mode="count"))


@pytest.mark.parametrize("mode", ["sat", "count"])
def test_plan_with_a_step_above_the_ridge(mode):
    """A three-tensor network whose first step is GEMM-shaped (m=9, n=7, k=6): inside a plan the step is lowered
    at run time to transpose -> tcgen05 GEMM -> transpose; the result must equal the oracle's."""
    import numpy as np
    import torch

    from oracle import fin_rel_oracle as O
    from tensorsat_b200 import _native

    rng = np.random.default_rng(5)
    # counting mode stores the step's output as uint8 while its bound fits 7 bits: 8 summed wires make it int32
    nk = 6 if mode == "sat" else 8
    M, N, K = list(range(0, 9)), list(range(9, 16)), list(range(16, 16 + nk))
    inputs = [tuple(int(w) for w in rng.permutation(M + K)), tuple(int(w) for w in rng.permutation(N + K)),
              tuple(M[:4] + N[:3])]
    dens = (0.01, 0.1, 0.02) if mode == "sat" else (0.5, 0.5, 0.5)
    leaves = [(rng.random((2,) * len(ws)) < p).astype(np.uint8) for ws, p in zip(inputs, dens)]
    out = tuple(M[4:] + N[3:])   # every M, N wire outlives the first step, the six K wires are summed there
    path = [(0, 1), (3, 2)]
    code = _native.MODE_SAT if mode == "sat" else _native.MODE_COUNT
    plan = _native.Plan(inputs, out, [2] * (16 + nk), path, (), mode=code)
    kinds = [o["kind"] for o in plan.describe()["ops"]]
    assert "gemm" in kinds
    dev = [_native.upload(t) for t in leaves]
    res = _native.empty(plan.info["out_numel"], code)
    plan.execute(dev, res)
    torch.cuda.synchronize()
    want = O.execute_tree(inputs, out, leaves, path, mode=mode)
    assert np.array_equal(res.cpu().numpy().reshape(want.shape), want)
    assert 0 < np.count_nonzero(want) 

"""GPU: pin the tcgen05/TMEM conventions (smem descriptor, instruction
descriptor, accumulator readback) of csrc/umma.cuh against a CPU fp64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 56), (16, 8), (208, 56), (256, 56), (64, 96)])
@pytest.mark.parametrize("split", [0, 1])
def test_umma_selftest(N, K, split):
    from matdeeplearn_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(N * 1000 + K)
    A = torch.randn(128, K)
    B = torch.randn(N, K)
    D = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.to(dev), B.to(dev)  # keep alive: a freed temporary's block would be recycled
    rc = lib.mdl_selftest_umma(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, split, _lib.stream())
    _lib.check(rc, "mdl_selftest_umma")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.cpu().double() - ref).abs().max().item()
    scale = (A.abs().double() @ B.abs().double().t()).max().item()
    # plain TF32: 10-bit mantissas -> ~1e-3 relative; 3xTF32: fp32-class
    tol = (2e-3 if split == 0 else 2e-6) * scale
    assert err < tol, (split, err, scale)

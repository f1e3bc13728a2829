"""GPU: pin the tcgen05/TMEM conventions (smem descriptor, instruction
descriptor, accumulator readback) of csrc/umma.cuh against a CPU fp64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 56), (16, 8), (208, 56), (256, 56), (64, 96)])
@pytest.mark.parametrize("split", [0, 1])
def test_umma_selftest(N, K, split):
    from matdeeplearn_b200 import _lib
    lib = _lib.load_selftest()
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


@pytest.mark.parametrize("N,K", [(128, 56), (64, 64), (16, 8), (256, 56), (64, 128)])
@pytest.mark.parametrize("split", [0, 1])
def test_umma_selftest_a_operand_in_tensor_memory(N, K, split):
    from matdeeplearn_b200 import _lib
    lib = _lib.load_selftest()
    dev = torch.device("cuda:0")
    torch.manual_seed(N * 1000 + K + 7)
    A = torch.randn(128, K)
    B = torch.randn(N, K)
    D = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.to(dev), B.to(dev)
    rc = lib.mdl_selftest_umma_ts(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, split, _lib.stream())
    _lib.check(rc, "mdl_selftest_umma_ts")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.cpu().double() - ref).abs().max().item()
    scale = (A.abs().double() @ B.abs().double().t()).max().item()
    tol = (2e-3 if split == 0 else 2e-6) * scale
    assert err < tol, (split, err, scale)


def _tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _tf32_rna(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_raw_fp32_operands_are_truncated_not_rounded():
    """kind::tf32 reads fp32 bit patterns from smem; the fused kernels rely on the hardware
    TRUNCATING the low 13 mantissa bits (so that lo = x - trunc(x) is the exact complement)."""
    from matdeeplearn_b200 import _lib
    lib = _lib.load_selftest()
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    N, K = 64, 64
    A = torch.randn(128, K)
    B = torch.randn(N, K)
    D = torch.empty(128, N, device=dev)
    Ad, Bd = A.to(dev), B.to(dev)
    _lib.check(lib.mdl_selftest_umma(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, 0, _lib.stream()), "umma")
    torch.cuda.synchronize()
    got = D.cpu().double()
    e_trunc = (got - _tf32_trunc(A).double() @ _tf32_trunc(B).double().t()).abs().max().item()
    e_rna = (got - _tf32_rna(A).double() @ _tf32_rna(B).double().t()).abs().max().item()
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/tf32_narrowing.txt", "w").write(f"err vs truncation {e_trunc:.3e}  err vs round-to-nearest {e_rna:.3e}\n")
    assert e_trunc < 1e-4 and e_trunc < 0.1 * e_rna, (e_trunc, e_rna)


"""GPU: the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu) against an fp64 matmul, at the edge-level shapes of the
BASELINE configs (SchNet filter network, MEGNet edge MLP, NNConv edge network) and awkward ones."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("R,I,O", [(102086, 128, 128), (102086, 50, 128), (173823, 128, 128), (102086, 64, 4096),
                                   (4096, 64, 64), (2049, 100, 150), (5000, 256, 8), (3000, 12, 40), (70000, 128, 256)])
def test_wgrad_tc_matches_fp64(R, I, O, monkeypatch):
    from matdeeplearn_b200 import functional as MF
    torch.manual_seed(R + I)
    x, g = torch.randn(R, I, device=DEV), torch.randn(R, O, device=DEV)
    ref_w = g.double().t().mm(x.double())
    ref_b = g.double().sum(0)
    wscale = (g.abs().double().t().mm(x.abs().double())).max().item()
    bscale = g.abs().double().sum(0).max().item()
    for impl in ("tc", "simt"):
        monkeypatch.setenv("MDL_WGRAD", impl)
        if impl == "simt" and I + O > 700:
            continue
        dW = torch.full((O, I), float("nan"), device=DEV)
        db = torch.full((O,), float("nan"), device=DEV)
        MF.linear_wgrad_into(x, g, MF._wgrad_map(O, I, [dW.data_ptr()], [db.data_ptr()]))
        assert (dW.double() - ref_w).abs().max().item() <= 2e-6 * wscale, impl
        assert (db.double() - ref_b).abs().max().item() <= 2e-6 * bscale, impl


def test_linear_fn_long_batch_uses_the_kernel_and_matches_autograd():
    from matdeeplearn_b200 import functional as MF, _lib
    torch.manual_seed(3)
    x = torch.randn(20000, 50, device=DEV, requires_grad=True)
    lin = torch.nn.Linear(50, 128).to(DEV)
    w = torch.randn(20000, 128, device=DEV)
    n0 = _lib.launch_count()
    (MF.linear(x, lin.weight, lin.bias) * w).sum().backward()
    assert _lib.launch_count() - n0 >= 2          # wgrad kernel + reduce went through the library
    got = (lin.weight.grad.clone(), lin.bias.grad.clone(), x.grad.clone())
    lin.zero_grad(); x.grad = None
    (torch.nn.functional.linear(x, lin.weight, lin.bias) * w).sum().backward()
    for a, b in zip(got, (lin.weight.grad, lin.bias.grad, x.grad)):
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()


@pytest.mark.parametrize("E,G", [(102086, 50), (5000, 37), (2048, 64), (3001, 8)])
@pytest.mark.parametrize("act1", ["ssp", "relu"])
def test_edge_mlp2_fused_matches_fp64(E, G, act1):
    """The fused SchNet filter network + cutoff (csrc/edge_mlp.cu) against the same arithmetic in fp64:
    forward, and the four weight / bias gradients (reference schnet.py:81 -> PyG InteractionBlock.mlp, CFConv)."""
    from matdeeplearn_b200 import functional as MF
    from tests.util import assert_close
    torch.manual_seed(E + G)
    H = 128
    x = torch.rand(E, G, dtype=torch.float64)
    rs = torch.rand(E, dtype=torch.float64)
    lin1, lin2 = torch.nn.Linear(G, H).double(), torch.nn.Linear(H, H).double()
    act = (lambda t: torch.nn.functional.softplus(t) - 0.6931471805599453) if act1 == "ssp" else torch.relu
    ref = lin2(act(lin1(x))) * rs[:, None]
    w = torch.randn_like(ref)
    ref.backward(w)
    p32 = [t.detach().float().to(DEV).requires_grad_(True) for t in (lin1.weight, lin1.bias, lin2.weight, lin2.bias)]
    xg, rsg = x.float().to(DEV), rs.float().to(DEV)
    assert MF.edge_mlp2_supported(xg, p32[0], p32[2])
    got = MF.edge_mlp2(xg, p32[0], p32[1], p32[2], p32[3], rsg, act1)
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what="edge_mlp2 fwd")
    got.backward(w.float().to(DEV))
    for g, r, name in zip(p32, (lin1.weight, lin1.bias, lin2.weight, lin2.bias), ("dW1", "db1", "dW2", "db2")):
        assert_close(g.grad, r.grad, rtol=1e-4, atol_rel=2e-5, what=f"edge_mlp2 {name}")


@pytest.mark.parametrize("R,K,N", [(102086, 128, 128), (173823, 128, 128), (20000, 50, 128), (16384, 128, 64), (17001, 100, 72)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_tc_forward_and_input_gradient(R, K, N, act):
    """mdl_linear_tc (csrc/linear_tc.cu): y = act(x W^T + b) and dx = g W against fp64."""
    from matdeeplearn_b200 import _lib
    from tests.util import assert_close
    lib = _lib.load()
    torch.manual_seed(R + K + act)
    x = torch.randn(R, K, dtype=torch.float64)
    w = torch.randn(N, K, dtype=torch.float64) * 0.2
    b = torch.randn(N, dtype=torch.float64)
    pre = x @ w.t() + b
    ref = pre if act == 0 else torch.relu(pre) if act == 1 else torch.nn.functional.softplus(pre) - 0.6931471805599453
    xg, wg, bg = x.float().to(DEV), w.float().to(DEV), b.float().to(DEV)
    y = torch.full((R, N), float("nan"), device=DEV)
    _lib.check(lib.mdl_linear_tc(_lib.ptr(xg), _lib.ptr(wg), K, 1, _lib.ptr(bg), _lib.ptr(y), R, K, N, act, _lib.stream()), "fwd")
    assert_close(y, ref, rtol=1e-5, atol_rel=5e-6, what=f"linear_tc fwd act={act}")
    if act == 0:
        g = torch.randn(R, N, dtype=torch.float64)
        gg = g.float().to(DEV)
        dx = torch.full((R, K), float("nan"), device=DEV)
        _lib.check(lib.mdl_linear_tc(_lib.ptr(gg), _lib.ptr(wg), 1, K, None, _lib.ptr(dx), R, N, K, 0, _lib.stream()), "dx")
        assert_close(dx, g @ w, rtol=1e-5, atol_rel=5e-6, what="linear_tc dx")

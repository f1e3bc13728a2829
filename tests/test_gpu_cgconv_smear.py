"""GPU parity of the smearing-fused CGConv form (mdl_cgconv_smear_fwd / _bwd): the edge kernels take the normalised
distance d_hat [E] and expand the reference's GaussianSmearing (process/process.py:580-590, applied at :500-502)
themselves.  Checked against the fp64 oracle fed the [E, G] tensor the reference would store, and against the
materialised-edge_attr kernels on the same inputs."""
import numpy as np
import pytest
import torch

from tests.util import assert_close, random_graph, block_diagonal_graph

pytestmark = pytest.mark.gpu

FWD = dict(rtol=1e-5, atol_rel=2e-6)
BWD = dict(rtol=1e-4, atol_rel=2e-5)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _default_dispatch(monkeypatch):
    for k in ("MDL_CGCONV_IMPL", "MDL_CGCONV_BWD", "MDL_CGCONV_DETERMINISTIC"):
        monkeypatch.delenv(k, raising=False)


def _reference_basis(d_hat64, G, start=0.0, stop=1.0, width=0.2):
    """GaussianSmearing.forward as the reference writes it, in fp64 on the fp32 offsets torch.linspace gives."""
    offset = torch.linspace(start, stop, G).double()
    coeff = -0.5 / ((stop - start) * width) ** 2
    return torch.exp(coeff * (d_hat64.view(-1, 1) - offset.view(1, -1)) ** 2)


def _case(dev, ei, n, C, G, aggr, seed, width=0.2, window="1", monkeypatch=None):
    import matdeeplearn_b200.nn as mnn
    from matdeeplearn_b200 import _lib
    from matdeeplearn_b200.data import GaussianEdgeAttr
    from oracle import pyg_ops as O
    if monkeypatch is not None:
        monkeypatch.setenv("MDL_CGCONV_WINDOW", window)
    torch.manual_seed(seed)
    E = ei.shape[1]
    x = torch.randn(n, C, dtype=torch.float64)
    d_hat = torch.rand(E).float()
    d_hat[::7] = 0.0            # self-loop-like zeros
    d_hat[3::11] = 1.0          # the far end of the normalised range
    ea64 = _reference_basis(d_hat.double(), G, width=width)
    ref_conv = O.CGConv(C, G, aggr=aggr).double()
    conv = mnn.CGConv(C, G, aggr=aggr)
    conv.load_state_dict({k: v.float() for k, v in ref_conv.state_dict().items()})
    conv = conv.to(dev)
    xr = x.clone().requires_grad_(True)
    ref = ref_conv(xr, ei, ea64)
    lazy = GaussianEdgeAttr(d_hat.to(dev), resolution=G, width=width)
    assert lazy.shape == (E, G)
    n0 = _lib.launch_count()
    xg = x.float().to(dev).requires_grad_(True)
    got = conv(xg, ei.to(dev), lazy)
    assert lazy._dense is None, "the fused form must not materialise edge_attr"
    assert_close(got, ref, **FWD, what=f"smear-fused cgconv fwd C={C} G={G} {aggr}")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(dev))
    assert_close(xg.grad, xr.grad, **BWD, what="dx")
    for name, pr in ref_conv.named_parameters():
        assert_close(dict(conv.named_parameters())[name].grad, pr.grad, **BWD, what=f"d{name}")
    # the materialised form on the same inputs (edge_attr expanded by the GaussianSmearing kernel)
    conv.zero_grad()
    xm = x.float().to(dev).requires_grad_(True)
    got_m = conv(xm, ei.to(dev), lazy.materialize())
    assert_close(got, got_m, rtol=2e-5, atol_rel=4e-6, what="fused vs materialised forward")
    return got


@pytest.mark.parametrize("G", [50, 64, 37, 8, 2, 51])
def test_smear_fused_matches_oracle(dev, G):
    ei = random_graph(300, 3000, 1)
    _case(dev, ei, 300, 64, G, "mean", seed=G)


def test_smear_fused_add_aggr_hub_isolated(dev):
    ei = random_graph(600, 4000, 2, hub=(11, 400), isolated=7)
    _case(dev, ei, 600, 64, 50, "add", seed=2)


@pytest.mark.parametrize("window", ["1", "0"])
@pytest.mark.parametrize("sizes,k", [([30] * 40, 12), ([4, 60, 9, 33, 58, 17, 41] * 9, 12), ([100, 20, 130, 7, 64, 64], 12),
                                     ([1] * 200 + [25] * 8, 6)])
def test_smear_fused_crystal_batches(dev, monkeypatch, sizes, k, window):
    ei = block_diagonal_graph(sizes, k, seed=len(sizes))
    _case(dev, ei, sum(sizes), 64, 50, "mean", seed=3, window=window, monkeypatch=monkeypatch)


def test_smear_fused_tiny_and_multi_tile(dev):
    _case(dev, random_graph(2, 1, 0), 2, 64, 50, "mean", seed=0)
    _case(dev, random_graph(6000, 70000, 4), 6000, 64, 50, "mean", seed=4)


def test_smear_fused_narrow_basis(dev):
    """width 0.1: coeff = -50, basis values down to exp(-50) -- the chunked recurrence must not drift"""
    _case(dev, random_graph(300, 3000, 5), 300, 64, 50, "mean", seed=5, width=0.1)


@pytest.mark.parametrize("C", [100, 128, 192])
def test_smear_fused_wide_layers(dev, C):
    """layer widths above 64 run as 64-channel chunks of the same kernels (the last chunk of C = 100 overlaps)"""
    _case(dev, random_graph(300, 3000, 7), 300, C, 50, "mean", seed=C)
    _case(dev, block_diagonal_graph([30] * 20, 12, seed=2), 600, C, 50, "add", seed=C + 1)


def test_smear_unsupported_shapes_materialise(dev):
    """C < 64 is not served by the fused form: CGConv expands the basis and takes the general kernels"""
    import matdeeplearn_b200.nn as mnn
    from matdeeplearn_b200 import functional as MF
    from matdeeplearn_b200.data import GaussianEdgeAttr
    assert MF.cgconv_smear_supported(64, 50) and MF.cgconv_smear_supported(100, 50) and not MF.cgconv_smear_supported(32, 50)
    torch.manual_seed(0)
    ei = random_graph(100, 800, 6).to(dev)
    lazy = GaussianEdgeAttr(torch.rand(ei.shape[1], device=dev), resolution=50)
    conv = mnn.CGConv(32, 50, aggr="mean").to(dev)
    x = torch.randn(100, 32, device=dev)
    a = conv(x, ei, lazy)
    b = conv(x, ei, lazy.materialize())
    assert lazy._dense is not None and torch.equal(a, b)


def test_cgcnn_model_lazy_edge_attr_matches_oracle(dev):
    """the whole model on a batch whose edge_attr is the 4 B/edge form (what TrainStep.from_host / from_store feed it)"""
    from matdeeplearn_b200 import models as M, process as pr
    from oracle import models as OM
    ds = pr.synthetic_dataset("bulk", 24, seed=7)
    batch = ds.batch()
    torch.manual_seed(0)
    cfg = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=2)
    ref_model = OM.CGCNN(ds, **cfg).double()
    model = M.CGCNN(ds, **cfg)
    model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref_model.state_dict().items()})
    model = model.to(dev)
    ref = ref_model(batch.double())
    lazy_batch = batch.with_lazy_edge_attr().to(dev)
    got = model(lazy_batch)
    assert lazy_batch.edge_attr._dense is None
    assert_close(got, ref, rtol=1e-4, atol_rel=1e-5, what="CGCNN forward (lazy edge_attr)")
    torch.nn.functional.l1_loss(ref, batch.y.double()).backward()
    torch.nn.functional.l1_loss(got, batch.y.to(dev)).backward()
    for name, pr_ in ref_model.named_parameters():
        assert_close(dict(model.named_parameters())[name].grad, pr_.grad, rtol=1e-3, atol_rel=5e-4, what=f"grad {name}")


def test_store_step_lazy_equals_materialised(dev):
    """TrainStep.from_store keeps edge_attr in the 4 B/edge form; same losses as eager steps on materialised batches"""
    from matdeeplearn_b200 import models as M, process as pr
    from matdeeplearn_b200.engine import TrainStep
    from matdeeplearn_b200.store import GraphStore
    ds = pr.synthetic_dataset("bulk", 48, seed=11)
    store = GraphStore.from_dataset(ds, dev)
    cfg = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)
    idxs = [np.random.default_rng(s).permutation(48)[:16] for s in range(4)]
    losses = []
    for mode in ("store", "eager"):
        torch.manual_seed(0)
        step = TrainStep(M.CGCNN(ds, **cfg).to(dev).train(), lr=1e-3)
        out = []
        for idx in idxs:
            if mode == "store":
                out.append(step.from_store(store, idx))
            else:
                out.append(float(step.eager(store.batch(idx)).item()))
        losses.append(out)
    assert np.allclose(losses[0], losses[1], rtol=2e-4, atol=1e-6), losses

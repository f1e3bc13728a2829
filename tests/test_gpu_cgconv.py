"""GPU parity: engine layout, segmented reduce and the fused CGConv kernels vs
the CPU oracle (fp64 ground truth), through the C ABI."""
import numpy as np
import pytest
import torch

from tests.util import assert_close, random_graph, contiguous_batch_vector, block_diagonal_graph

pytestmark = pytest.mark.gpu

# fp32 kernels vs fp64 oracle.  Forward values: rtol 1e-5 (+ 2e-6 of the tensor's
# scale as absolute floor); gradients (long fp32 sums over E): rtol 1e-4.
FWD = dict(rtol=1e-5, atol_rel=2e-6)
BWD = dict(rtol=1e-4, atol_rel=2e-5)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _csr_reference(ei, n):
    row, col = ei[0].numpy(), ei[1].numpy()
    eid = np.argsort(col, kind="stable")
    dst_dst, dst_src = col[eid], row[eid]
    pos = np.argsort(dst_src, kind="stable")
    dst_ptr = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=n))])
    src_ptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n))])
    return eid, dst_src, dst_dst, pos, dst_ptr, src_ptr


@pytest.mark.parametrize("n,e,hub,iso", [(40, 300, None, 0), (500, 6000, (7, 300), 5), (3, 2, None, 1)])
def test_csr_matches_stable_sort(dev, n, e, hub, iso):
    from matdeeplearn_b200.csr import GraphCSR
    ei = random_graph(n, e, 1, hub, iso)
    batch = contiguous_batch_vector(n, min(4, n), 2)
    csr = GraphCSR.from_coo(ei.to(dev), batch.to(dev), num_graphs=min(4, n))
    eid, dst_src, dst_dst, pos, dst_ptr, src_ptr = _csr_reference(ei, n)
    assert np.array_equal(csr.dst_eid.cpu().numpy(), eid)
    assert np.array_equal(csr.dst_src.cpu().numpy(), dst_src)
    assert np.array_equal(csr.dst_dst.cpu().numpy(), dst_dst)
    assert np.array_equal(csr.src_slot.cpu().numpy(), pos)
    assert np.array_equal(csr.dst_ptr.cpu().numpy(), dst_ptr)
    assert np.array_equal(csr.src_ptr.cpu().numpy(), src_ptr)
    gp = np.concatenate([[0], np.cumsum(np.bincount(batch.numpy(), minlength=min(4, n)))])
    assert np.array_equal(csr.graph_ptr.cpu().numpy(), gp)
    indeg = np.maximum(1, np.diff(dst_ptr))
    assert np.allclose(csr.inv_deg_dst.cpu().numpy(), 1.0 / indeg)


def test_csr_empty_edges(dev):
    from matdeeplearn_b200.csr import GraphCSR
    ei = torch.zeros(2, 0, dtype=torch.int64, device=dev)
    csr = GraphCSR.from_coo(ei, num_nodes=5)
    assert csr.dst_ptr.cpu().tolist() == [0] * 6 and csr.src_ptr.cpu().tolist() == [0] * 6


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
@pytest.mark.parametrize("width", [1, 64, 100, 300])
def test_scatter_matches_oracle(dev, reduce, width):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(0)
    n, s = 777, 40
    index = contiguous_batch_vector(n, s - 3, 3)  # 3 trailing empty segments
    src = torch.randn(n, width, dtype=torch.float64)
    ref_in = src.clone().requires_grad_(True)
    ref = O.scatter(ref_in, index, 0, s, reduce)
    got_in = src.float().to(dev).requires_grad_(True)
    got = mnn.scatter(got_in, index.to(dev), 0, s, reduce)
    assert_close(got, ref, **FWD, what=f"scatter {reduce}")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(dev))
    assert_close(got_in.grad, ref_in.grad, **FWD, what=f"scatter {reduce} grad")


def test_scatter_unsorted_index(dev):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(1)
    index = torch.randint(0, 50, (1000,))
    src = torch.randn(1000, 32, dtype=torch.float64)
    for red in ("sum", "mean", "max"):
        ref = O.scatter(src, index, 0, None, red)
        got = mnn.scatter(src.float().to(dev), index.to(dev), 0, None, red)
        assert_close(got, ref, **FWD, what=red)


@pytest.fixture(params=["ws", "ws_nowin", "pipe", "pipe_nowin", "tc", "tc_det", "tc_nowin", "tc_rows", "simt"])
def impl(request, monkeypatch):
    """Run a case on the tensor-core kernels (default dispatch: single-pass backward with vector
    atomics for dQ; shapes that do not fit fall back to SIMT inside the library), on the
    tensor-core kernels in deterministic two-pass mode, and with the SIMT kernels forced."""
    # "tc_nowin": tensor-core kernels with the shared-memory node-row window off (per-slot rows only)
    # "tc_rows": edge rows by per-row cp.async instead of one bulk (TMA) copy per round
    monkeypatch.setenv("MDL_CGCONV_WINDOW", "0" if request.param.endswith("_nowin") else "1")
    monkeypatch.setenv("MDL_CGCONV_EA", "rows" if request.param == "tc_rows" else "bulk")
    # "pipe": default dispatch (software-pipelined forward kernel, cgconv_fwd.cu; tc backward)
    # "ws": default dispatch (warp-specialised forward kernel, cgconv_fwd_ws.cu; single-pass tcgen05 backward)
    monkeypatch.setenv("MDL_CGCONV_IMPL", request.param if request.param in ("simt",) else
                       ("ws" if request.param.startswith("ws") else "pipe" if request.param.startswith("pipe") else "tc"))
    monkeypatch.setenv("MDL_CGCONV_DETERMINISTIC", "1" if request.param == "tc_det" else "0")
    # backward: "pipe*" = default single-pass kernel with dW_e on tcgen05 (cgconv_bwd.cu); "tc*" = the mma.sync one
    monkeypatch.setenv("MDL_CGCONV_BWD", "tc" if request.param.startswith("tc") else "pipe")
    return request.param


def _log_err(tag, got, ref):
    import os
    err = (got.detach().double().cpu() - ref.detach().double().cpu()).abs().max().item()
    scale = ref.detach().abs().max().item()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_errors.txt", "a") as f:
        f.write(f"{tag}: max|err|={err:.3e} scale={scale:.3e} rel={err / max(scale, 1e-30):.3e}\n")


def _cgconv_case(dev, n, e, C, G, aggr, hub=None, iso=0, seed=0, tag="", ei=None):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(seed)
    if ei is None:
        ei = random_graph(n, e, seed, hub, iso)
    E = ei.shape[1]
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(E, G, dtype=torch.float64)
    ref_conv = O.CGConv(C, G, aggr=aggr).double()
    conv = mnn.CGConv(C, G, aggr=aggr)
    conv.load_state_dict({k: v.float() for k, v in ref_conv.state_dict().items()})
    conv = conv.to(dev)
    xr = x.clone().requires_grad_(True)
    ref = ref_conv(xr, ei, ea)
    xg = x.float().to(dev).requires_grad_(True)
    got = conv(xg, ei.to(dev), ea.float().to(dev))
    _log_err(f"cgconv fwd {tag} C={C} G={G} {aggr}", got, ref)
    assert_close(got, ref, **FWD, what=f"cgconv fwd C={C} G={G} {aggr}")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(dev))
    _log_err(f"cgconv dx {tag} C={C} G={G}", xg.grad, xr.grad)
    assert_close(xg.grad, xr.grad, **BWD, what="dx")
    for name, pr in ref_conv.named_parameters():
        pg = dict(conv.named_parameters())[name]
        _log_err(f"cgconv d{name} {tag} C={C} G={G}", pg.grad, pr.grad)
        assert_close(pg.grad, pr.grad, **BWD, what=f"d{name}")
    return got


@pytest.mark.parametrize("C,G", [(64, 50), (100, 50), (128, 100), (64, 200), (8, 3), (128, 200)])
def test_cgconv_shapes(dev, impl, C, G):
    _cgconv_case(dev, n=300, e=3000, C=C, G=G, aggr="mean", tag=impl)


def test_wide_layers_dispatch_to_the_tensor_core_kernels():
    """configs[0]'s default width (dim1 = 100, reference config.yml:121-123) and SchNet-style C = 128 CGCNNs run on the
    tcgen05 kernels (64-channel chunks), forward and backward (G = 64: forward only, the backward's ea^T tiles do not fit);
    widths below 64 and edge widths above 64 do not"""
    from matdeeplearn_b200 import _lib
    lib = _lib.load()
    for C, G, want in [(64, 50, 3), (100, 50, 3), (128, 50, 3), (128, 64, 1), (256, 37, 3), (32, 50, 0), (64, 100, 0), (100, 200, 0)]:
        assert lib.mdl_cgconv_tc_supported(C, G) == want, (C, G)


@pytest.mark.parametrize("C,G", [(100, 50), (128, 50), (68, 37), (192, 64), (128, 51)])
@pytest.mark.parametrize("graph", ["random", "crystal", "hub"])
def test_cgconv_wide_layers(dev, impl, C, G, graph):
    """C > 64 on every implementation (tensor-core kernels: 64-channel chunks, overlapping last chunk when C % 64 != 0)"""
    if graph == "crystal":
        ei = block_diagonal_graph([30, 7, 52, 18] * 6, 12, seed=C)
        _cgconv_case(dev, n=int(ei.max()) + 1, e=0, C=C, G=G, aggr="mean", seed=C, tag=impl + " wide", ei=ei)
    elif graph == "hub":
        _cgconv_case(dev, n=500, e=3000, C=C, G=G, aggr="add", hub=(5, 300), iso=4, seed=C, tag=impl + " wide")
    else:
        _cgconv_case(dev, n=300, e=3000, C=C, G=G, aggr="mean", seed=C, tag=impl + " wide")


def test_cgconv_odd_edge_width(dev, impl):
    _cgconv_case(dev, n=150, e=1200, C=32, G=37, aggr="mean", tag=impl)


@pytest.mark.parametrize("G", [37, 51, 1, 63])
def test_cgconv_odd_edge_width_c64(dev, impl, G):
    """odd edge widths on the C = 64 tensor-core kernels: edge rows start at any 4-byte offset of the landing
    zone (4-byte loads in the split), the last k-chunk is partly padding, the tail of ea is patched from global"""
    _cgconv_case(dev, n=300, e=2500, C=64, G=G, aggr="mean", tag=impl + " oddG")


def test_cgconv_add_aggr(dev, impl):
    _cgconv_case(dev, n=200, e=1500, C=64, G=50, aggr="add", tag=impl)


def test_cgconv_hub_and_isolated(dev, impl):
    # in-degree 400 (> several rounds of 128 slots) and 7 nodes with no edges at all
    _cgconv_case(dev, n=600, e=4000, C=64, G=50, aggr="mean", hub=(11, 400), iso=7, tag=impl)


@pytest.mark.parametrize("sizes,k", [([30] * 40, 12), ([4, 60, 9, 33, 58, 17, 41] * 9, 12), ([5] * 300, 3),
                                     ([100, 20, 130, 7, 64, 64], 12), ([1] * 200 + [25] * 8, 6)])
def test_cgconv_crystal_batches(dev, impl, sizes, k):
    """Block-diagonal batches: the rounds' source / destination rows sit in short contiguous node
    ranges, which is the case the shared-memory window serves (and where it must hand over to the
    per-slot rows: blocks above the 128-row capacity, rounds that straddle several blocks)."""
    ei = block_diagonal_graph(sizes, k, seed=len(sizes))
    got = _cgconv_case(dev, n=sum(sizes), e=0, C=64, G=50, aggr="mean", seed=3, tag=impl, ei=ei)
    if impl in ("tc", "tc_nowin", "pipe", "pipe_nowin", "ws", "ws_nowin"):
        import os
        os.environ["MDL_CGCONV_WINDOW"] = "1" if impl.endswith("_nowin") else "0"
        other = _cgconv_case(dev, n=sum(sizes), e=0, C=64, G=50, aggr="mean", seed=3, ei=ei)
        os.environ["MDL_CGCONV_WINDOW"] = "0" if impl.endswith("_nowin") else "1"
        assert torch.equal(got, other), "window and per-slot staging must give bit-identical forwards"


def test_cgconv_tiny(dev, impl):
    _cgconv_case(dev, n=2, e=1, C=64, G=50, aggr="mean", tag=impl)


def test_cgconv_large_multi_tile(dev, impl):
    # > 148 tiles: every persistent CTA walks several tiles
    _cgconv_case(dev, n=6000, e=70000, C=64, G=50, aggr="mean", tag=impl)


def test_cgconv_deterministic(dev, impl):
    a = _cgconv_case(dev, n=400, e=5000, C=64, G=50, aggr="mean", seed=5)
    b = _cgconv_case(dev, n=400, e=5000, C=64, G=50, aggr="mean", seed=5)
    assert torch.equal(a, b), "CSR reduction must be bitwise reproducible"


def test_cgconv_backward_bitwise_reproducible_in_deterministic_mode(dev, monkeypatch):
    import matdeeplearn_b200.nn as mnn
    monkeypatch.setenv("MDL_CGCONV_IMPL", "tc")
    monkeypatch.setenv("MDL_CGCONV_DETERMINISTIC", "1")
    torch.manual_seed(9)
    ei = random_graph(500, 6000, 9).to(dev)
    x0 = torch.randn(500, 64, device=dev)
    ea = torch.rand(ei.shape[1], 50, device=dev)
    conv = mnn.CGConv(64, 50, aggr="mean").to(dev)
    w = torch.randn(500, 64, device=dev)
    grads = []
    for _ in range(2):
        x = x0.clone().requires_grad_(True)
        conv.zero_grad()
        conv(x, ei, ea).backward(w)
        grads.append([x.grad.clone()] + [p.grad.clone() for p in conv.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)


def test_cgconv_rejects_cpu_tensors():
    import matdeeplearn_b200.nn as mnn
    conv = mnn.CGConv(8, 4, aggr="mean")
    with pytest.raises(RuntimeError):
        conv(torch.randn(3, 8), torch.zeros(2, 1, dtype=torch.long), torch.randn(1, 4))


def test_cgcnn_model_matches_oracle(dev):
    from matdeeplearn_b200 import models as M, process as pr
    from oracle import models as OM
    ds = pr.synthetic_dataset("bulk", 24, seed=7)
    batch = ds.batch()
    torch.manual_seed(0)
    cfg = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=2)
    ref_model = OM.CGCNN(ds, **cfg).double()
    model = M.CGCNN(ds, **cfg)
    model.load_state_dict({k: v.float() if v.is_floating_point() else v
                           for k, v in ref_model.state_dict().items()})
    model = model.to(dev)
    ref = ref_model(batch.double())
    got = model(batch.to(dev))
    assert_close(got, ref, rtol=1e-4, atol_rel=1e-5, what="CGCNN forward")
    ref_loss = torch.nn.functional.l1_loss(ref, batch.y.double())
    got_loss = torch.nn.functional.l1_loss(got, batch.y.to(dev))
    ref_loss.backward()
    got_loss.backward()
    for name, pr_ in ref_model.named_parameters():
        pg = dict(model.named_parameters())[name]
        assert_close(pg.grad, pr_.grad, rtol=1e-3, atol_rel=5e-4, what=f"grad {name}")

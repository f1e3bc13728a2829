"""GPU: the CSR graph primitives behind SchNet / MPNN / MEGNet on irregular graphs -- a hub with
in-degree far above a warp, isolated nodes (no edges, no self-loop), a single-edge graph -- against
the oracle (fp64), values and gradients."""
import pytest
import torch

from tests.util import assert_close, random_graph, contiguous_batch_vector

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = [dict(n=400, e=3000, hub=(5, 350), iso=6), dict(n=3, e=1, hub=None, iso=1), dict(n=64, e=400, hub=None, iso=0)]


@pytest.mark.parametrize("case", CASES)
def test_cfconv_irregular(case):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(0)
    ei = random_graph(case["n"], case["e"], 7, case["hub"], case["iso"])
    E, n, C, G, Fi = ei.shape[1], case["n"], 32, 11, 48
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(E, G, dtype=torch.float64)
    ew = torch.rand(E, dtype=torch.float64) * 8
    ref_m = O.InteractionBlock(C, G, Fi, 8.0).double()
    m = mnn.InteractionBlock(C, G, Fi, 8.0)
    m.load_state_dict({k: v.float() for k, v in ref_m.state_dict().items()})
    m = m.to(DEV)
    xr = x.clone().requires_grad_(True)
    ref = ref_m(xr, ei, ew, ea)
    xg = x.float().to(DEV).requires_grad_(True)
    got = m(xg, ei.to(DEV), ew.float().to(DEV), ea.float().to(DEV))
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what="cfconv fwd")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(DEV))
    assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what="cfconv dx")
    for name, pr in ref_m.named_parameters():
        assert_close(dict(m.named_parameters())[name].grad, pr.grad, rtol=1e-4, atol_rel=2e-5, what=name)


@pytest.mark.parametrize("case", CASES)
def test_nnconv_irregular(case):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(1)
    ei = random_graph(case["n"], case["e"], 8, case["hub"], case["iso"])
    E, n, C, G, K = ei.shape[1], case["n"], 24, 9, 12
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(E, G, dtype=torch.float64)

    def net():
        return torch.nn.Sequential(torch.nn.Linear(G, K), torch.nn.ReLU(), torch.nn.Linear(K, C * C))

    for aggr in ("mean", "add"):
        ref_m = O.NNConv(C, C, net(), aggr=aggr).double()
        m = mnn.NNConv(C, C, net(), aggr=aggr)
        m.load_state_dict({k: v.float() for k, v in ref_m.state_dict().items()})
        m = m.to(DEV)
        xr = x.clone().requires_grad_(True)
        ref = ref_m(xr, ei, ea)
        xg = x.float().to(DEV).requires_grad_(True)
        got = m(xg, ei.to(DEV), ea.float().to(DEV))
        assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what=f"nnconv fwd {aggr}")
        w = torch.randn_like(ref)
        ref.backward(w)
        got.backward(w.float().to(DEV))
        assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what="nnconv dx")
        for name, pr in ref_m.named_parameters():
            assert_close(dict(m.named_parameters())[name].grad, pr.grad, rtol=1e-4, atol_rel=2e-5, what=name)


def test_scatter_fast_paths_agree_with_generic_path():
    """scatter(e, edge_index[0]) / scatter(x, batch) resolve their segments from the cached
    GraphCSR; a fresh copy of the same index goes through the generic sort.  Same numbers."""
    import matdeeplearn_b200.nn as mnn
    from matdeeplearn_b200.csr import csr_for
    torch.manual_seed(2)
    n, B = 300, 7
    ei = random_graph(n, 2500, 9, (3, 200), 4).to(DEV)
    batch = contiguous_batch_vector(n, B, 10).to(DEV)
    e = torch.randn(ei.shape[1], 20, device=DEV)
    x = torch.randn(n, 20, device=DEV)
    csr_for(ei, batch, num_nodes=n, num_graphs=B)
    for red in ("mean", "sum", "max"):
        fast = mnn.scatter(e, ei[0, :], dim=0, reduce=red)
        slow = mnn.scatter(e, ei[0, :].clone(), dim=0, dim_size=n, reduce=red)
        assert_close(fast, slow, rtol=1e-6, atol_rel=1e-6, what=f"by-source {red}")
        fast = mnn.scatter(e, ei[1, :], dim=0, reduce=red)
        slow = mnn.scatter(e, ei[1, :].clone(), dim=0, dim_size=n, reduce=red)
        assert_close(fast, slow, rtol=1e-6, atol_rel=1e-6, what=f"by-destination {red}")
        fast = mnn.scatter(x, batch, dim=0, reduce=red)
        slow = mnn.scatter(x, batch.clone(), dim=0, dim_size=B, reduce=red)
        assert_close(fast, slow, rtol=1e-6, atol_rel=1e-6, what=f"by-graph {red}")


def test_megnet_model_on_irregular_batch_matches_oracle():
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.data import Batch
    from oracle import models as OM
    from tests.test_oracle_golden import _DS
    torch.manual_seed(3)
    n, B, G = 120, 5, 8
    ei = random_graph(n, 700, 11, (9, 60), 0)
    E = ei.shape[1]
    b = Batch(x=torch.randn(n, 114), edge_index=ei, edge_attr=torch.rand(E, G), edge_weight=torch.rand(E) * 8,
              batch=contiguous_batch_vector(n, B, 12), u=torch.zeros(B, 3), y=torch.randn(B))
    b.num_graphs = B
    cfg = dict(dim1=32, dim2=16, dim3=32, gc_count=2, gc_fc_count=1, post_fc_count=1)
    ref_m = OM.MEGNet(_DS(b), **cfg).double()
    m = M.MEGNet(_DS(b), **cfg)
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref_m.state_dict().items()})
    m = m.to(DEV).train()
    ref = ref_m(b.double())
    got = m(b.to(DEV))
    assert_close(got, ref, rtol=1e-3, atol_rel=1e-3, what="MEGNet irregular fwd")


def test_spmm_scalar_and_edge_dot_match_dense_expression():
    """out[i] = sum_{e -> i} coef[e] h[src(e)] (GCNConv's weighted neighbour sum, reference gcn.py:141) and its
    gradients wrt h (by-source view) and wrt the per-edge coefficient (row dot product, mdl_edge_dot)."""
    from matdeeplearn_b200 import functional as MF
    from matdeeplearn_b200.csr import GraphCSR
    from tests.util import random_graph, assert_close
    torch.manual_seed(5)
    n, F_ = 257, 48
    ei = random_graph(n, 2100, 3, hub=(9, 150), isolated=4)
    E = ei.shape[1]
    h = torch.randn(n, F_, dtype=torch.float64, requires_grad=True)
    c = torch.rand(E, dtype=torch.float64, requires_grad=True)
    ref = torch.zeros(n, F_, dtype=torch.float64).index_add(0, ei[1], c[:, None] * h[ei[0]])
    w = torch.randn_like(ref)
    ref.backward(w)
    csr = GraphCSR.from_coo(ei.to("cuda:0"), num_nodes=n)
    hg = h.detach().float().to("cuda:0").requires_grad_(True)
    cg = c.detach().float().to("cuda:0").requires_grad_(True)
    got = MF.spmm_scalar(hg, cg, csr)
    got.backward(w.float().to("cuda:0"))
    assert_close(got, ref, rtol=1e-5, atol_rel=2e-6, what="spmm_scalar")
    assert_close(hg.grad, h.grad, rtol=1e-4, atol_rel=2e-5, what="spmm_scalar dh")
    assert_close(cg.grad, c.grad, rtol=1e-4, atol_rel=2e-5, what="edge_dot")

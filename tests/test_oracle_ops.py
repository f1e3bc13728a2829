"""CPU: the oracle's operator restatements on hand-worked graphs and through the invariants
the formulas imply (there are no reference tests to lean on)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pyg_ops as O


def test_scatter_semantics_match_torch_scatter_documentation():
    src = torch.tensor([[1.0, -2.0], [3.0, 4.0], [5.0, -6.0]])
    idx = torch.tensor([2, 0, 2])
    assert O.scatter(src, idx, 0, 4, "sum").tolist() == [[3, 4], [0, 0], [6, -8], [0, 0]]
    assert O.scatter(src, idx, 0, 4, "mean").tolist() == [[3, 4], [0, 0], [3, -4], [0, 0]]   # count clamps at 1
    assert O.scatter(src, idx, 0, 4, "max").tolist() == [[3, 4], [0, 0], [5, -2], [0, 0]]    # empty -> 0
    assert O.scatter(src, idx, 0, None, "sum").shape[0] == 3                                # dim_size = max+1
    assert torch.equal(O.global_mean_pool(src, torch.tensor([0, 0, 1])),
                       torch.tensor([[2.0, 1.0], [5.0, -6.0]]))


def test_cgconv_hand_worked_two_node_graph():
    # nodes 0,1 ; edges 0->1, 1->1 (loop) ; C=1, G=1 ; weights chosen by hand
    conv = O.CGConv(1, 1, aggr="mean").double()
    with torch.no_grad():
        conv.lin_f.weight.copy_(torch.tensor([[0.5, -1.0, 2.0]]))   # [x_i, x_j, e]
        conv.lin_f.bias.fill_(0.1)
        conv.lin_s.weight.copy_(torch.tensor([[1.0, 0.25, -0.5]]))
        conv.lin_s.bias.fill_(-0.2)
    x = torch.tensor([[2.0], [-1.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 1], [1, 1]])
    ea = torch.tensor([[0.3], [1.0]], dtype=torch.float64)
    out = conv(x, ei, ea)

    def msg(xi, xj, e):
        f = 0.5 * xi - 1.0 * xj + 2.0 * e + 0.1
        s = 1.0 * xi + 0.25 * xj - 0.5 * e - 0.2
        return 1 / (1 + math.exp(-f)) * math.log1p(math.exp(s))

    m01 = msg(-1.0, 2.0, 0.3)   # target i=1, source j=0
    m11 = msg(-1.0, -1.0, 1.0)
    assert out[0].item() == pytest.approx(2.0)                       # no in-edges: mean over empty = 0, + x
    assert out[1].item() == pytest.approx((m01 + m11) / 2 - 1.0, rel=1e-12)


def _rand_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    return ei


@pytest.mark.parametrize("make", ["cgconv", "interaction", "nnconv"])
def test_edge_permutation_invariance_and_block_independence(make):
    torch.manual_seed(0)
    n, e, C, G = 12, 40, 8, 5
    ei = _rand_graph(n, e, 1)
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(e, G, dtype=torch.float64)
    ew = torch.rand(e, dtype=torch.float64) * 8
    if make == "cgconv":
        m = O.CGConv(C, G, aggr="mean").double()
        f = lambda x_, ei_, ea_, ew_: m(x_, ei_, ea_)
    elif make == "interaction":
        m = O.InteractionBlock(C, G, 6, 8.0).double()
        f = lambda x_, ei_, ea_, ew_: m(x_, ei_, ew_, ea_)
    else:
        net = torch.nn.Sequential(torch.nn.Linear(G, 7), torch.nn.ReLU(), torch.nn.Linear(7, C * C))
        m = O.NNConv(C, C, net, aggr="mean").double()
        f = lambda x_, ei_, ea_, ew_: m(x_, ei_, ea_)
    ref = f(x, ei, ea, ew)
    perm = torch.randperm(e)
    torch.testing.assert_close(f(x, ei[:, perm], ea[perm], ew[perm]), ref, rtol=1e-12, atol=1e-12)
    # block-diagonal union of two graphs == the two graphs processed separately
    x2 = torch.cat([x, x * 0.5 + 1])
    ei2 = torch.cat([ei, ei + n], 1)
    out2 = f(x2, ei2, torch.cat([ea, ea]), torch.cat([ew, ew]))
    torch.testing.assert_close(out2[:n], ref, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(out2[n:], f(x * 0.5 + 1, ei, ea, ew), rtol=1e-12, atol=1e-12)


def test_cgconv_mean_equals_add_times_inverse_degree():
    torch.manual_seed(1)
    n, e, C, G = 10, 30, 4, 3
    ei = _rand_graph(n, e, 2)
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(e, G, dtype=torch.float64)
    a = O.CGConv(C, G, aggr="add").double()
    b = O.CGConv(C, G, aggr="mean").double()
    b.load_state_dict(a.state_dict())
    deg = torch.bincount(ei[1], minlength=n).clamp(min=1).double().view(-1, 1)
    torch.testing.assert_close((a(x, ei, ea) - x) / deg + x, b(x, ei, ea), rtol=1e-12, atol=1e-12)


def test_nnconv_equals_explicit_per_edge_matrices():
    torch.manual_seed(2)
    n, e, C, G, K = 6, 15, 3, 4, 5
    ei = _rand_graph(n, e, 3)
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(e, G, dtype=torch.float64)
    net = torch.nn.Sequential(torch.nn.Linear(G, K), torch.nn.ReLU(), torch.nn.Linear(K, C * C)).double()
    m = O.NNConv(C, C, net, aggr="add").double()
    out = m(x, ei, ea)
    ref = x @ m.lin.weight.t() + m.bias
    ref = ref.clone()
    for k in range(e):
        theta = net(ea[k]).view(C, C)
        ref[ei[1, k]] += x[ei[0, k]] @ theta
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=1e-12)


def test_oracle_gradients_by_finite_differences():
    torch.manual_seed(3)
    n, e, C, G = 5, 12, 3, 2
    ei = _rand_graph(n, e, 4)
    ea = torch.rand(e, G, dtype=torch.float64)
    conv = O.CGConv(C, G, aggr="mean").double()
    x = torch.randn(n, C, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: conv(t, ei, ea), (x,), eps=1e-6, atol=1e-6)
    blk = O.InteractionBlock(C, G, 4, 8.0).double()
    ew = torch.rand(e, dtype=torch.float64) * 8
    assert torch.autograd.gradcheck(lambda t: blk(t, ei, ew, ea), (x,), eps=1e-6, atol=1e-6)


def test_interaction_block_cutoff_and_shifted_softplus():
    assert O.ShiftedSoftplus()(torch.zeros(1)).abs().item() < 1e-7           # ssp(0) = 0
    blk = O.InteractionBlock(2, 3, 4, cutoff=8.0).double()
    x = torch.randn(3, 2, dtype=torch.float64)
    ei = torch.tensor([[0, 1], [2, 2]])
    ea = torch.rand(2, 3, dtype=torch.float64)
    # an edge at exactly the cutoff distance contributes nothing (cosine cutoff = 0)
    a = blk(x, ei, torch.tensor([8.0, 3.0], dtype=torch.float64), ea)
    b = blk(x, ei[:, 1:], torch.tensor([3.0], dtype=torch.float64), ea[1:])
    torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-12)


def test_gcnconv_hand_worked_against_dense_normalised_adjacency():
    # out = D^-1/2 A D^-1/2 (x W^T) + b with A[i, j] = weight of edge j -> i, D = row sums of A
    # (PyG gcn_norm without added self-loops; a node whose in-weights sum to 0 gets coefficient 0)
    torch.manual_seed(0)
    n, C = 5, 3
    ei = torch.tensor([[0, 1, 2, 3, 1, 4, 4], [1, 0, 1, 1, 2, 2, 4]])   # node 3 has no in-edges, 4 only a loop
    ew = torch.tensor([2.0, 0.5, 1.5, 4.0, 3.0, 1.0, 0.0], dtype=torch.float64)  # loop weight 0 (raw distance)
    x = torch.randn(n, C, dtype=torch.float64)
    conv = O.GCNConv(C, C, improved=True, add_self_loops=False).double()
    with torch.no_grad():
        conv.bias.copy_(torch.tensor([0.1, -0.2, 0.3]))
    A = torch.zeros(n, n, dtype=torch.float64)
    A[ei[1], ei[0]] = ew
    deg = A.sum(1)
    dinv = torch.where(deg > 0, deg.pow(-0.5), torch.zeros_like(deg))
    ref = (dinv[:, None] * A * dinv[None, :]) @ (x @ conv.lin.weight.t()) + conv.bias
    out = conv(x, ei, ew)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    assert torch.allclose(out[3], conv.bias) and torch.allclose(out[4], conv.bias)   # degree 0 -> bias only
    assert not out.isnan().any()


def test_segment_softmax_and_set2set_hand_worked():
    src = torch.tensor([[1.0], [2.0], [0.5], [-1.0], [3.0]], dtype=torch.float64)
    batch = torch.tensor([0, 0, 1, 1, 1])
    a = O.segment_softmax(src, batch, 2)
    assert torch.allclose(a[:2, 0], torch.softmax(src[:2, 0], 0)) and torch.allclose(a[2:, 0], torch.softmax(src[2:, 0], 0))
    assert torch.allclose(O.scatter(a, batch, 0, 2, "sum"), torch.ones(2, 1, dtype=torch.float64))
    # Set2Set against an explicit per-graph loop over the same LSTM
    torch.manual_seed(1)
    C = 4
    s2s = O.Set2Set(C, processing_steps=3).double()
    x = torch.randn(5, C, dtype=torch.float64)
    got = s2s(x, batch)
    rows = []
    for g in range(2):
        xg = x[batch == g]
        h = (torch.zeros(1, 1, C, dtype=torch.float64), torch.zeros(1, 1, C, dtype=torch.float64))
        q_star = torch.zeros(1, 2 * C, dtype=torch.float64)
        for _ in range(3):
            q, h = s2s.lstm(q_star.unsqueeze(0), h)
            q = q.view(1, C)
            att = torch.softmax(xg @ q.t(), 0)
            q_star = torch.cat([q, (att * xg).sum(0, keepdim=True)], -1)
        rows.append(q_star)
    assert got.shape == (2, 2 * C)
    assert torch.allclose(got, torch.cat(rows, 0), rtol=1e-10, atol=1e-12)


def test_metalayer_call_convention_hand_worked():
    """PyG MetaLayer as the reference composes it (megnet.py:235-239, 310-333): the edge model is called with
    (x[row], x[col], edge_attr, u, batch[row]) -- row = edge_index[0] the SOURCE --, then the node model with the NEW
    edge_attr, then the global model with the NEW x and edge_attr."""
    seen = {}

    class Edge(torch.nn.Module):
        def forward(self, src, dest, e, u, b):
            seen["edge"] = (src.clone(), dest.clone(), e.clone(), u.clone(), b.clone())
            return e + src.sum(1, keepdim=True) - dest.sum(1, keepdim=True) + u[b]

    class Node(torch.nn.Module):
        def forward(self, x, ei, e, u, b):
            seen["node_e"] = e.clone()
            return x + O.scatter_mean(e, ei[0], dim=0, dim_size=x.shape[0])

    class Glob(torch.nn.Module):
        def forward(self, x, ei, e, u, b):
            seen["glob_x"] = x.clone()
            return u + O.scatter_mean(x, b, dim=0)

    x = torch.tensor([[1.0], [10.0], [100.0]])
    ei = torch.tensor([[0, 1, 2, 2], [1, 0, 0, 2]])            # 0->1, 1->0, 2->0, loop 2->2
    e = torch.tensor([[0.5], [0.25], [0.125], [0.0]])
    u = torch.tensor([[1000.0], [2000.0]])
    batch = torch.tensor([0, 0, 1])
    x2, e2, u2 = O.MetaLayer(Edge(), Node(), Glob())(x, ei, e, u, batch)
    src, dest, e_in, u_in, b_in = seen["edge"]
    assert src.view(-1).tolist() == [1.0, 10.0, 100.0, 100.0] and dest.view(-1).tolist() == [10.0, 1.0, 1.0, 100.0]
    assert b_in.tolist() == [0, 0, 1, 1]                       # graph of the SOURCE node
    want_e = torch.tensor([[0.5 + 1 - 10 + 1000], [0.25 + 10 - 1 + 1000], [0.125 + 100 - 1 + 2000], [0.0 + 2000]])
    torch.testing.assert_close(e2, want_e)
    torch.testing.assert_close(seen["node_e"], want_e)         # node model sees the updated edges
    want_x = x + torch.stack([want_e[0], want_e[1], (want_e[2] + want_e[3]) / 2])   # mean over edges LEAVING each node
    torch.testing.assert_close(x2, want_x)
    torch.testing.assert_close(seen["glob_x"], want_x)         # global model sees the updated nodes
    torch.testing.assert_close(u2, u + torch.stack([(want_x[0] + want_x[1]) / 2, want_x[2]]))


def test_megnet_node_and_global_models_aggregate_edges_at_their_source():
    """megnet.py:86 and :130: scatter_mean(edge_attr, edge_index[0]) -- edges are averaged at their SOURCE node (unlike
    the convs, which aggregate at edge_index[1]); the global model then averages nodes and node-averaged edges per graph
    (megnet.py:131-132)."""
    from oracle import models as OM
    D = 2
    node = OM.Megnet_NodeModel(D, "relu", "False", "True", 0.0, fc_layers=0).double()
    glob = OM.Megnet_GlobalModel(D, "relu", "False", "True", 0.0, fc_layers=0).double()
    for m in (node, glob):                                     # identity-like first layer: output = relu(sum of the 3 blocks)
        lin = getattr(m, m._list_name)[0]
        with torch.no_grad():
            lin.weight.copy_(torch.cat([torch.eye(D)] * 3, 1).double())
            lin.bias.zero_()
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0], [2.0, 2.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 0, 1, 2], [1, 2, 0, 2]])
    e = torch.tensor([[1.0, 0.0], [3.0, 0.0], [0.0, 5.0], [7.0, 7.0]], dtype=torch.float64)
    u = torch.tensor([[10.0, 10.0], [20.0, 20.0]], dtype=torch.float64)
    batch = torch.tensor([0, 0, 1])
    v_e = torch.tensor([[2.0, 0.0], [0.0, 5.0], [7.0, 7.0]], dtype=torch.float64)      # by source: node 0 <- edges 0,1
    torch.testing.assert_close(node(x, ei, e, u, batch), x + v_e + u[batch])
    u_e = torch.stack([(v_e[0] + v_e[1]) / 2, v_e[2]])
    u_v = torch.stack([(x[0] + x[1]) / 2, x[2]])
    torch.testing.assert_close(glob(x, ei, e, u, batch), u_e + u_v + u)


def test_cfconv_equals_explicit_edge_loop():
    """PyG CFConv (schnet.py:81 -> InteractionBlock.conv): out_i = lin2( sum_{j->i} lin1(x)_j * (mlp(e_ij) * C(d_ij)) ),
    C(d) = (cos(pi d / cutoff) + 1) / 2, aggregated at edge_index[1]."""
    torch.manual_seed(3)
    n, E, Cc, G, Fw, cutoff = 5, 9, 3, 4, 6, 8.0
    mlp = torch.nn.Sequential(torch.nn.Linear(G, Fw), O.ShiftedSoftplus(), torch.nn.Linear(Fw, Fw)).double()
    conv = O.CFConv(Cc, Cc, Fw, mlp, cutoff).double()
    x = torch.randn(n, Cc, dtype=torch.float64)
    ei = torch.randint(0, n, (2, E))
    d = torch.rand(E, dtype=torch.float64) * cutoff
    ea = torch.rand(E, G, dtype=torch.float64)
    got = conv(x, ei, d, ea)
    h = x @ conv.lin1.weight.t()
    agg = torch.zeros(n, Fw, dtype=torch.float64)
    for k in range(E):
        j, i = int(ei[0, k]), int(ei[1, k])
        w = mlp(ea[k]) * 0.5 * (math.cos(math.pi * float(d[k]) / cutoff) + 1.0)
        agg[i] += h[j] * w
    torch.testing.assert_close(got, agg @ conv.lin2.weight.t() + conv.lin2.bias, rtol=1e-12, atol=1e-12)

"""Oracle restatement of the torch_geometric / torch_scatter operators the
reference's models call (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every operator is written as the op sequence PyG 2.0.1 executes
(index_select gathers -> cat -> Linear -> activation -> index_add_ scatter ->
count-divide), so that timing it on the CPU is a fair stand-in for the
reference's CPU path.  PARITY UNPINNED for these
formulas (no reference tests exist); spec = SURVEY.md Appendix A.

Direction convention (PyG flow="source_to_target"): row = edge_index[0] is the
source j (gathered), col = edge_index[1] is the destination i (aggregated).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------
# torch_scatter.scatter / scatter_mean  (call sites: reference
# matdeeplearn/models/megnet.py:13,86,130-132,342-348)
# ----------------------------------------------------------------------------
def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """torch_scatter.scatter semantics for dim=0.

    dim_size defaults to index.max()+1; empty segments give 0 for every
    reduce (torch_scatter fills max/min of empty segments with 0).
    """
    assert dim == 0, "oracle only restates dim=0 (all reference call sites)"
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    if reduce in ("sum", "add"):
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return out.index_add_(0, index, src)
    if reduce == "mean":
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        out = out.index_add_(0, index, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt = cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        return out / cnt.view((-1,) + (1,) * (src.dim() - 1))
    if reduce == "max":
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
        return out.scatter_reduce(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def scatter_mean(src, index, dim=0, dim_size=None):
    return scatter(src, index, dim, dim_size, "mean")


def scatter_add(src, index, dim=0, dim_size=None):
    return scatter(src, index, dim, dim_size, "sum")


def scatter_max(src, index, dim=0, dim_size=None):
    return scatter(src, index, dim, dim_size, "max")


# ----------------------------------------------------------------------------
# torch_geometric.nn.global_*_pool  (resolved by name at reference
# matdeeplearn/models/cgcnn.py:154,169; schnet.py:152,167; mpnn.py:168,183)
# ----------------------------------------------------------------------------
def _num_graphs(batch):
    return int(batch.max()) + 1 if batch.numel() > 0 else 0


def global_mean_pool(x, batch, size=None):
    return scatter(x, batch, 0, size if size is not None else _num_graphs(batch), "mean")


def global_add_pool(x, batch, size=None):
    return scatter(x, batch, 0, size if size is not None else _num_graphs(batch), "sum")


def global_max_pool(x, batch, size=None):
    return scatter(x, batch, 0, size if size is not None else _num_graphs(batch), "max")


# ----------------------------------------------------------------------------
# Set2Set readout (constructed at reference cgcnn.py:114-119 and twins:
# Set2Set(post_fc_dim, processing_steps=3) early, Set2Set(output_dim, 3, num_layers=1) late)
# ----------------------------------------------------------------------------
def segment_softmax(src, index, num_segments):
    """PyG 2.0.1 torch_geometric.utils.softmax: exp(src - max_seg) / (sum_seg + 1e-16)."""
    smax = scatter(src, index, 0, num_segments, "max").index_select(0, index)
    out = (src - smax).exp()
    ssum = scatter(out, index, 0, num_segments, "sum").index_select(0, index)
    return out / (ssum + 1e-16)


class Set2Set(nn.Module):
    """q_t = LSTM(q*_{t-1}); a = softmax_graph(x . q_t); r_t = sum_graph a x; q*_t = [q_t || r_t]
    (PyG 2.0.1 Set2Set; output width 2 * in_channels)."""

    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, 2 * in_channels
        self.processing_steps, self.num_layers = processing_steps, num_layers
        self.lstm = nn.LSTM(self.out_channels, in_channels, num_layers)

    def forward(self, x, batch):
        B = _num_graphs(batch)
        h = (x.new_zeros((self.num_layers, B, self.in_channels)),
             x.new_zeros((self.num_layers, B, self.in_channels)))
        q_star = x.new_zeros(B, self.out_channels)
        for _ in range(self.processing_steps):
            q, h = self.lstm(q_star.unsqueeze(0), h)
            q = q.view(B, self.in_channels)
            e = (x * q.index_select(0, batch)).sum(dim=-1, keepdim=True)
            a = segment_softmax(e, batch, B)
            r = scatter(a * x, batch, 0, B, "sum")
            q_star = torch.cat([q, r], dim=-1)
        return q_star


# ----------------------------------------------------------------------------
# CGConv  (constructed at reference matdeeplearn/models/cgcnn.py:80-82 as
# CGConv(gc_dim, num_edge_features, aggr="mean", batch_norm=False))
# ----------------------------------------------------------------------------
class CGConv(nn.Module):
    """x_i' = x_i + aggr_j sigmoid(z W_f + b_f) * softplus(z W_s + b_s),
    z = [x_i || x_j || e_ij]   (PyG 2.0.1 CGConv)."""

    def __init__(self, channels, dim=0, aggr="add", batch_norm=False, bias=True):
        super().__init__()
        assert not batch_norm, "reference always passes batch_norm=False"
        self.channels, self.dim, self.aggr = channels, dim, aggr
        self.lin_f = nn.Linear(2 * channels + dim, channels, bias=bias)
        self.lin_s = nn.Linear(2 * channels + dim, channels, bias=bias)

    def forward(self, x, edge_index, edge_attr):
        row, col = edge_index[0], edge_index[1]
        x_i = x.index_select(0, col)
        x_j = x.index_select(0, row)
        z = torch.cat([x_i, x_j, edge_attr], dim=-1)
        m = torch.sigmoid(self.lin_f(z)) * F.softplus(self.lin_s(z))
        red = {"mean": "mean", "add": "sum", "sum": "sum", "max": "max"}[self.aggr]
        out = scatter(m, col, 0, x.size(0), red)
        return out + x


# ----------------------------------------------------------------------------
# GCNConv  (constructed at reference matdeeplearn/models/gcn.py:80-82 as
# GCNConv(gc_dim, gc_dim, improved=True, add_self_loops=False); called gcn.py:141-150
# with (x, edge_index, edge_weight) -- edge_weight is the RAW distance, 0 on the loops)
# ----------------------------------------------------------------------------
class GCNConv(nn.Module):
    """x' = D^-1/2 A D^-1/2 (x W^T) + b with A_ij = edge_weight, D_i = sum_{j->i} edge_weight
    (PyG 2.0.1 GCNConv + gcn_norm; without added self-loops `improved` changes nothing)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
                 normalize=True, bias=True):
        super().__init__()
        assert not add_self_loops and normalize and not cached, "only the reference's configuration is restated"
        self.in_channels, self.out_channels, self.improved = in_channels, out_channels, improved
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.xavier_uniform_(self.lin.weight)  # PyG: glorot weight, zero bias

    def forward(self, x, edge_index, edge_weight=None):
        row, col = edge_index[0], edge_index[1]
        n = x.size(0)
        if edge_weight is None:
            edge_weight = torch.ones(row.numel(), dtype=x.dtype, device=x.device)
        deg = scatter(edge_weight, col, 0, n, "sum")
        dinv = deg.pow(-0.5)
        dinv = dinv.masked_fill(dinv == float("inf"), 0.0)
        norm = dinv.index_select(0, row) * edge_weight * dinv.index_select(0, col)
        h = self.lin(x)
        out = scatter(norm.view(-1, 1) * h.index_select(0, row), col, 0, n, "sum")
        return out + self.bias if self.bias is not None else out


# ----------------------------------------------------------------------------
# SchNet InteractionBlock / CFConv  (constructed at reference
# matdeeplearn/models/schnet.py:81; called schnet.py:134-143 with
# (x, edge_index, edge_weight, edge_attr))
# ----------------------------------------------------------------------------
class ShiftedSoftplus(nn.Module):
    def __init__(self):
        super().__init__()
        self.shift = math.log(2.0)

    def forward(self, x):
        return F.softplus(x) - self.shift


class CFConv(nn.Module):
    def __init__(self, in_channels, out_channels, num_filters, mlp, cutoff):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, num_filters, bias=False)
        self.lin2 = nn.Linear(num_filters, out_channels)
        self.nn = mlp
        self.cutoff = cutoff
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.lin1.weight)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        row, col = edge_index[0], edge_index[1]
        C = 0.5 * (torch.cos(edge_weight * math.pi / self.cutoff) + 1.0)
        W = self.nn(edge_attr) * C.view(-1, 1)
        h = self.lin1(x)
        msg = h.index_select(0, row) * W
        agg = scatter(msg, col, 0, x.size(0), "sum")
        return self.lin2(agg)


class InteractionBlock(nn.Module):
    def __init__(self, hidden_channels, num_gaussians, num_filters, cutoff):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(num_gaussians, num_filters),
            ShiftedSoftplus(),
            nn.Linear(num_filters, num_filters),
        )
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, self.mlp, cutoff)
        self.act = ShiftedSoftplus()
        self.lin = nn.Linear(hidden_channels, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.mlp[0].weight)
        self.mlp[0].bias.data.fill_(0)
        nn.init.xavier_uniform_(self.mlp[2].weight)
        self.mlp[2].bias.data.fill_(0)
        self.conv.reset_parameters()
        nn.init.xavier_uniform_(self.lin.weight)
        self.lin.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        x = self.conv(x, edge_index, edge_weight, edge_attr)
        x = self.act(x)
        x = self.lin(x)
        return x


# ----------------------------------------------------------------------------
# NNConv  (constructed at reference matdeeplearn/models/mpnn.py:83-88 as
# NNConv(gc_dim, gc_dim, nn, aggr="mean"))
# ----------------------------------------------------------------------------
class NNConv(nn.Module):
    """x_i' = x_i W_root + b + aggr_j x_j . reshape(nn(e_ij), [C_in, C_out])."""

    def __init__(self, in_channels, out_channels, nn_module, aggr="add",
                 root_weight=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.nn = nn_module
        self.aggr = aggr
        self.lin = nn.Linear(in_channels, out_channels, bias=False) if root_weight else None
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        if self.lin is not None:
            bound = 1.0 / math.sqrt(in_channels)
            nn.init.uniform_(self.lin.weight, -bound, bound)

    def forward(self, x, edge_index, edge_attr):
        row, col = edge_index[0], edge_index[1]
        theta = self.nn(edge_attr).view(-1, self.in_channels, self.out_channels)
        x_j = x.index_select(0, row)
        m = torch.matmul(x_j.unsqueeze(1), theta).squeeze(1)
        red = {"mean": "mean", "add": "sum", "sum": "sum", "max": "max"}[self.aggr]
        out = scatter(m, col, 0, x.size(0), red)
        if self.lin is not None:
            out = out + self.lin(x)
        if self.bias is not None:
            out = out + self.bias
        return out


# ----------------------------------------------------------------------------
# MetaLayer  (constructed at reference matdeeplearn/models/megnet.py:235-239)
# ----------------------------------------------------------------------------
class MetaLayer(nn.Module):
    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = global_model

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        row, col = edge_index[0], edge_index[1]
        if self.edge_model is not None:
            edge_attr = self.edge_model(
                x[row], x[col], edge_attr, u, batch if batch is None else batch[row]
            )
        if self.node_model is not None:
            x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u

"""The reference's model classes (matdeeplearn/models/*.py) on the sm_100a
engine: same constructor keywords (config.yml Models section), same forward
semantics, same parameter names -- `getattr(models, name)(data=dataset, **cfg)`
as at reference training/training.py:250 works unchanged.

Only the glue lives here; all graph arithmetic goes through matdeeplearn_b200.nn.
Booleans arrive as the strings "True"/"False" exactly as in the reference's
YAML.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn as tnn

from . import functional as MF
from . import nn as mnn
from .csr import csr_for
from .data import dense_edge_attr


def _target_dim(dataset):
    y = dataset[0].y
    return 1 if y.ndim == 0 else len(y[0])


def _prepare(data):
    """Build (or fetch the memoised) engine layout for this batch once, with the
    graph segments attached, before any operator asks for it."""
    return csr_for(data.edge_index, data.batch, num_nodes=data.x.shape[0],
                   num_graphs=getattr(data, "num_graphs", None))


class _ConvStackModel(tnn.Module):
    """Dense in -> gc_count x (conv [+BN] + dropout) -> readout -> dense out.
    Common frame of CGCNN / SchNet / MPNN (reference cgcnn.py:17-174)."""

    def __init__(self, data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                 pool_order, batch_norm, batch_track_stats, act, dropout_rate):
        super().__init__()
        if gc_count <= 0:
            raise AssertionError("Need at least 1 GC layer")
        self.batch_track_stats = not (batch_track_stats == "False")
        self.batch_norm = batch_norm
        self.pool, self.pool_order, self.act = pool, pool_order, act
        self.dropout_rate = dropout_rate
        n_in = data.num_features
        self.gc_dim = dim1 if pre_fc_count > 0 else n_in
        self.pre_lin_list = tnn.ModuleList(
            tnn.Linear(n_in if i == 0 else dim1, dim1) for i in range(pre_fc_count))
        self.conv_list = tnn.ModuleList()
        self.bn_list = tnn.ModuleList()
        # Set2Set doubles the width it pools (reference cgcnn.py:95-119)
        s2s = pool == "set2set"
        post_in = 2 * self.gc_dim if (s2s and pool_order == "early") else self.gc_dim
        out_dim = _target_dim(data)
        self.post_lin_list = tnn.ModuleList(
            tnn.Linear(post_in if i == 0 else dim2, dim2) for i in range(post_fc_count))
        self.lin_out = tnn.Linear(dim2 if post_fc_count > 0 else post_in, out_dim)
        if s2s and pool_order == "early":
            self.set2set = mnn.Set2Set(self.gc_dim, processing_steps=3)
        elif s2s and pool_order == "late":
            self.set2set = mnn.Set2Set(out_dim, processing_steps=3, num_layers=1)
            self.lin_out_2 = tnn.Linear(out_dim * 2, out_dim)

    def _add_bn(self):
        if self.batch_norm == "True":
            self.bn_list.append(tnn.BatchNorm1d(self.gc_dim, track_running_stats=self.batch_track_stats))

    def _activation(self, t):
        return getattr(F, self.act)(t)

    def _embed(self, data):
        h = data.x
        for lin in self.pre_lin_list:
            h = self._activation(MF.linear(h, lin.weight, lin.bias))
        return h

    def _bn(self, i, h, data):
        """BatchNorm after conv i.  The engine's own kernels take over when the batch is capacity-padded
        (statistics over the real rows only) or when gradients are delivered straight into the flat
        buffer (dist.FlatParameters.enable_direct); otherwise the torch module runs."""
        bn = self.bn_list[i]
        nv = getattr(data, "_n_valid", None)
        own = nv is not None or (bn.training and h.is_cuda and bn.weight is not None
                                 and getattr(bn.weight, "_mdl_grad_dest", None) is not None)
        return MF.masked_batch_norm(bn, h, nv) if own else bn(h)

    def _readout(self, h, data):
        s2s = self.pool == "set2set"
        if s2s and getattr(data, "_n_valid", None) is not None:
            raise NotImplementedError("Set2Set on capacity-padded batches (its softmax would see the padding rows)")
        pool = self.set2set if s2s else getattr(mnn, self.pool)
        if self.pool_order == "early":
            h = pool(h, data.batch)
        for lin in self.post_lin_list:
            h = self._activation(MF.linear(h, lin.weight, lin.bias))
        h = MF.linear(h, self.lin_out.weight, self.lin_out.bias)
        if self.pool_order == "late":
            h = pool(h, data.batch)
            if s2s:
                h = MF.linear(h, self.lin_out_2.weight, self.lin_out_2.bias)
        return h.view(-1) if h.shape[1] == 1 else h


class CGCNN(_ConvStackModel):
    def __init__(self, data, dim1=64, dim2=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        for _ in range(gc_count):
            self.conv_list.append(mnn.CGConv(self.gc_dim, data.num_edge_features, aggr="mean",
                                             batch_norm=False))
            self._add_bn()

    def forward(self, data):
        csr = _prepare(data)
        h = self._embed(data)
        for i, conv in enumerate(self.conv_list):
            h = conv(h, data.edge_index, data.edge_attr, csr=csr)
            if self.batch_norm == "True":
                h = self._bn(i, h, data)
            h = F.dropout(h, p=self.dropout_rate, training=self.training)
        return self._readout(h, data)


class GCN(_ConvStackModel):
    """reference matdeeplearn/models/gcn.py:17-178 (conv -> BN -> act -> dropout, gcn.py:139-152)"""

    def __init__(self, data, dim1=64, dim2=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        for _ in range(gc_count):
            self.conv_list.append(mnn.GCNConv(self.gc_dim, self.gc_dim, improved=True, add_self_loops=False))
            self._add_bn()

    def forward(self, data):
        csr = _prepare(data)
        h = self._embed(data)
        for i, conv in enumerate(self.conv_list):
            h = conv(h, data.edge_index, data.edge_weight, csr=csr)
            if self.batch_norm == "True":
                h = self._bn(i, h, data)
            h = self._activation(h)
            h = F.dropout(h, p=self.dropout_rate, training=self.training)
        return self._readout(h, data)


class SchNet(_ConvStackModel):
    """reference matdeeplearn/models/schnet.py:16-172"""

    def __init__(self, data, dim1=64, dim2=64, dim3=64, cutoff=8, pre_fc_count=1, gc_count=3,
                 post_fc_count=1, pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        for _ in range(gc_count):
            self.conv_list.append(mnn.InteractionBlock(self.gc_dim, data.num_edge_features, dim3, cutoff))
            self._add_bn()

    def forward(self, data):
        csr = _prepare(data)
        h = self._embed(data)
        for i, conv in enumerate(self.conv_list):
            h = h + conv(h, data.edge_index, data.edge_weight, data.edge_attr, csr=csr)  # residual outside the block
            if self.batch_norm == "True":
                h = self._bn(i, h, data)
            h = F.dropout(h, p=self.dropout_rate, training=self.training)
        return self._readout(h, data)


class MPNN(_ConvStackModel):
    """reference matdeeplearn/models/mpnn.py:17-188"""

    def __init__(self, data, dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        g = self.gc_dim
        self.gru_list = tnn.ModuleList()
        for _ in range(gc_count):
            edge_net = tnn.Sequential(tnn.Linear(data.num_edge_features, dim3), tnn.ReLU(),
                                      tnn.Linear(dim3, g * g))
            self.conv_list.append(mnn.NNConv(g, g, edge_net, aggr="mean"))
            self.gru_list.append(tnn.GRU(g, g))
            self._add_bn()

    def forward(self, data):
        csr = _prepare(data)
        out = self._embed(data)
        hidden = out.unsqueeze(0)          # GRU state threads through all layers (mpnn.py:141-144)
        for i, conv in enumerate(self.conv_list):
            m = conv(out, data.edge_index, data.edge_attr, csr=csr)
            if self.batch_norm == "True":
                m = self._bn(i, m, data)
            m = self._activation(m)
            m = F.dropout(m, p=self.dropout_rate, training=self.training)
            out, hidden = self.gru_list[i](m.unsqueeze(0), hidden)  # cuDNN, TF32 off (package __init__)
            out = out.squeeze(0)
        return self._readout(out, data)


# ---------------------------------------------------------------------------- MEGNet
class _MegnetStack(tnn.Module):
    """(Linear -> act -> BatchNorm -> dropout) x (fc_layers + 1), reference megnet.py:28-56."""

    def __init__(self, first_in, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, name):
        super().__init__()
        self.act, self.batch_norm, self.dropout_rate = act, batch_norm, dropout_rate
        # reference quirk: MEGNet passes a bool here and compares it with the string "False"
        track = not (batch_track_stats == "False")
        self._name = name
        setattr(self, name, tnn.ModuleList(
            tnn.Linear(first_in if i == 0 else dim, dim) for i in range(fc_layers + 1)))
        self.bn_list = tnn.ModuleList(
            tnn.BatchNorm1d(dim, track_running_stats=track) for _ in range(fc_layers + 1)
        ) if batch_norm == "True" else tnn.ModuleList()

    def _run(self, h, first_done=False, n_valid=None):
        """n_valid: device-side count of real rows of a capacity-padded batch (statistics over those rows only, the
        rest written as zero); None = every row is real and the torch module runs."""
        layers = getattr(self, self._name)
        for i, lin in enumerate(layers):
            if not (i == 0 and first_done):
                if self.act == "relu":
                    h = MF.linear(h, lin.weight, lin.bias, relu=True)
                else:
                    h = getattr(F, self.act)(MF.linear(h, lin.weight, lin.bias))
            if self.batch_norm == "True":
                h = MF.masked_batch_norm(self.bn_list[i], h, n_valid) if n_valid is not None else self.bn_list[i](h)
            h = F.dropout(h, p=self.dropout_rate, training=self.training)
        return h


class Megnet_EdgeModel(_MegnetStack):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(dim * 4, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "edge_mlp")
        self.dim = dim

    def forward(self, src, dest, edge_attr, u, batch):      # PyG MetaLayer call convention
        return self._run(torch.cat([src, dest, edge_attr, u[batch]], dim=1))

    def forward_fused(self, x, edge_index, edge_attr, u, batch):
        """Linear(cat[x[row], x[col], e, u[batch[row]]]) = x W_s^T [row] + x W_d^T [col] + e W_e^T
        + u W_u^T [batch[row]] + b : three node/graph-level GEMMs, one edge-level GEMM, one fused
        gather-add(-ReLU) kernel; the [E, 4D] concatenation never exists."""
        D = self.dim
        lin = self.edge_mlp[0]
        W = lin.weight
        A = x @ W[:, :D].t()
        B = x @ W[:, D:2 * D].t()
        base = MF.linear(edge_attr, W[:, 2 * D:3 * D].contiguous())     # the edge-level GEMM of the block
        U = u @ W[:, 3 * D:].t()
        e_valid = getattr(edge_index, "_mdl_e_valid", None)
        if e_valid is not None:
            # capacity-padded batch: padded edges point at a padding node whose graph id is B -- give them a zero row
            U = torch.cat([U, U.new_zeros(1, U.shape[1])], 0)
        csr = csr_for(edge_index, batch, num_nodes=x.shape[0], num_graphs=u.shape[0])
        relu = self.act == "relu"
        h = MF_edge_gather_add(base, A, B, U, lin.bias, edge_index, batch, csr, relu)
        if not relu:
            h = getattr(F, self.act)(h)
        return self._run(h, first_done=True, n_valid=e_valid)


class Megnet_NodeModel(_MegnetStack):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(dim * 3, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "node_mlp")

    def forward(self, x, edge_index, edge_attr, u, batch):
        v_e = _edge_mean_by_source(edge_attr, edge_index)               # by SOURCE node (megnet.py:86)
        return self._run(torch.cat([x, v_e, _expand_graph_rows(u, batch, x.shape[0])], dim=1),
                         n_valid=getattr(batch, "_mdl_n_valid", None))


class Megnet_GlobalModel(_MegnetStack):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(dim * 3, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "global_mlp")

    def forward(self, x, edge_index, edge_attr, u, batch):
        u_e = _edge_mean_by_source(edge_attr, edge_index)               # megnet.py:130 (the same tensor as :86)
        u_e = mnn.scatter_mean(u_e, batch, dim=0)
        u_v = mnn.scatter_mean(x, batch, dim=0)
        return self._run(torch.cat([u_e, u_v, u], dim=1))


def _edge_mean_by_source(edge_attr, edge_index):
    """scatter_mean(edge_attr, edge_index[0]): the reference computes it in the node model (megnet.py:86) and again
    in the global model (megnet.py:130) on the same tensors; here the second call reuses the first result (memoised
    on the tensor object, valid for its current version)."""
    hit = getattr(edge_attr, "_mdl_src_mean", None)
    if hit is not None and hit[0] == edge_attr._version and hit[1] is edge_index:
        return hit[2]
    out = mnn.scatter_mean(edge_attr, edge_index[0, :], dim=0)
    try:
        edge_attr._mdl_src_mean = (edge_attr._version, edge_index, out)
    except Exception:
        pass
    return out


def _expand_graph_rows(u, batch, n):
    """u[batch] (megnet.py:99); on CUDA through the gather kernel with the segmented sum as its backward."""
    seg = getattr(batch, "_mdl_seg", None)
    ok = (u.is_cuda and u.dtype == torch.float32 and seg is not None and seg[0] == batch._version and seg[2] is None
          and seg[1].shape[0] - 1 == u.shape[0])
    if getattr(batch, "_mdl_n_valid", None) is not None:
        # capacity-padded batch: padding nodes carry graph id B; they read a zero row (and are outside every segment,
        # so the segmented-sum backward -- deterministic, unlike index_select's atomics -- never sees them)
        u = torch.cat([u, u.new_zeros(1, u.shape[1])], 0)
        return MF.expand_by_segment(u, batch, seg[1]) if ok else u.index_select(0, batch)
    if ok:
        return MF.expand_by_segment(u, batch, seg[1])
    return u[batch]


def MF_edge_gather_add(*args):
    from . import functional as MF
    return MF.edge_gather_add(*args)


class MEGNet(tnn.Module):
    """reference matdeeplearn/models/megnet.py:150-371"""

    def __init__(self, data, dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3, gc_fc_count=2,
                 post_fc_count=1, pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__()
        if gc_count <= 0:
            raise AssertionError("Need at least 1 GC layer")
        track = not (batch_track_stats == "False")
        self.batch_norm, self.pool, self.act = batch_norm, pool, act
        self.pool_order, self.dropout_rate = pool_order, dropout_rate
        self.pool_reduce = {"global_mean_pool": "mean", "global_max_pool": "max",
                            "global_sum_pool": "sum"}.get(pool)   # reference leaves add_pool undefined
        gc_dim = dim1 if pre_fc_count > 0 else data.num_features
        self.pre_lin_list = tnn.ModuleList(
            tnn.Linear(data.num_features if i == 0 else dim1, dim1) for i in range(pre_fc_count))

        def embed(n_in):
            return tnn.Sequential(tnn.Linear(n_in, dim3), tnn.ReLU(), tnn.Linear(dim3, dim3), tnn.ReLU())

        self.e_embed_list, self.x_embed_list = tnn.ModuleList(), tnn.ModuleList()
        self.u_embed_list, self.conv_list = tnn.ModuleList(), tnn.ModuleList()
        self.bn_list = tnn.ModuleList()
        for i in range(gc_count):
            self.e_embed_list.append(embed(data.num_edge_features if i == 0 else dim3))
            self.x_embed_list.append(embed(gc_dim if i == 0 else dim3))
            self.u_embed_list.append(embed(data[0].u.shape[1] if i == 0 else dim3))
            args = (dim3, act, batch_norm, track, dropout_rate, gc_fc_count)
            self.conv_list.append(mnn.MetaLayer(Megnet_EdgeModel(*args), Megnet_NodeModel(*args),
                                                Megnet_GlobalModel(*args)))
        post_in = dim3 * 3 if pool_order == "early" else dim3
        self.post_lin_list = tnn.ModuleList(
            tnn.Linear(post_in if i == 0 else dim2, dim2) for i in range(post_fc_count))
        self.lin_out = tnn.Linear(dim2 if post_fc_count > 0 else post_in, _target_dim(data))

    def forward(self, data):
        _prepare(data)
        act = getattr(F, self.act)
        h = data.x
        for lin in self.pre_lin_list:
            h = act(lin(h))
        x, e, u = h, dense_edge_attr(data.edge_attr), data.u
        for i, block in enumerate(self.conv_list):
            e_t = MF.apply_mlp(self.e_embed_list[i], e)     # edge-level: long batch -> tensor-core weight gradients
            x_t, u_t = MF.apply_mlp(self.x_embed_list[i], x), self.u_embed_list[i](u)
            x_o, e_o, u_o = block(x_t, data.edge_index, e_t, u_t, data.batch)
            if i == 0:   # first block: residual onto the embedded inputs (megnet.py:313-315)
                x, e, u = x_o + x_t, e_o + e_t, u_o + u_t
            else:        # later blocks: onto the running state (megnet.py:334-336)
                x, e, u = x_o + x, e_o + e, u_o + u
        if self.pool_order == "early":
            x_pool = mnn.scatter(x, data.batch, dim=0, reduce=self.pool_reduce)
            e_pool = mnn.scatter(e, data.edge_index[0, :], dim=0, reduce=self.pool_reduce)
            e_pool = mnn.scatter(e_pool, data.batch, dim=0, reduce=self.pool_reduce)
            out = torch.cat([x_pool, e_pool, u], dim=1)
            for lin in self.post_lin_list:
                out = act(lin(out))
            out = self.lin_out(out)
        else:
            out = x
            for lin in self.post_lin_list:
                out = act(lin(out))
            out = self.lin_out(out)
            out = getattr(mnn, self.pool)(out, data.batch)
        return out.view(-1) if out.shape[1] == 1 else out

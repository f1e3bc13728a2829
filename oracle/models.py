"""Oracle restatement of the reference's model glue around the PyG operators
(TEST INFRASTRUCTURE -- see oracle/__init__.py).

Parameter names match the reference modules so a state_dict moves freely
between the reference glue (tests/golden/make_golden.py), this oracle and the
CUDA-backed models in matdeeplearn_b200.models.

  CGCNN   reference matdeeplearn/models/cgcnn.py:17-174
  SchNet  reference matdeeplearn/models/schnet.py:16-172
  MPNN    reference matdeeplearn/models/mpnn.py:17-188
  MEGNet  reference matdeeplearn/models/megnet.py:16-371
  GCN     reference matdeeplearn/models/gcn.py:17-178
Set2Set readout: cgcnn.py:95-119,150-168 (and the same lines of schnet / mpnn / gcn).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import pyg_ops as P


def _out_dim(data):
    y = data[0].y
    return 1 if y.ndim == 0 else len(y[0])


def _mlp_stack(n, d_in, d_hidden):
    return nn.ModuleList([nn.Linear(d_in if i == 0 else d_hidden, d_hidden) for i in range(n)])


class _Skeleton(nn.Module):
    """pre-FC -> convs (subclass) -> pool/post-FC -> lin_out, shared by
    CGCNN / SchNet / MPNN (reference cgcnn.py:121-174 and twins)."""

    def __init__(self, data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                 pool_order, batch_norm, batch_track_stats, act, dropout_rate):
        super().__init__()
        assert gc_count > 0
        self.batch_track_stats = batch_track_stats != "False"
        self.batch_norm, self.pool, self.act = batch_norm, pool, act
        self.pool_order, self.dropout_rate = pool_order, dropout_rate
        self.gc_dim = dim1 if pre_fc_count > 0 else data.num_features
        self.output_dim = _out_dim(data)
        self.pre_lin_list = _mlp_stack(pre_fc_count, data.num_features, dim1)
        # Set2Set doubles the width it pools (reference cgcnn.py:95-110)
        post_in = 2 * self.gc_dim if (pool == "set2set" and pool_order == "early") else self.gc_dim
        self.post_lin_list = _mlp_stack(post_fc_count, post_in, dim2)
        self.lin_out = nn.Linear(dim2 if post_fc_count > 0 else post_in, self.output_dim)
        if pool == "set2set" and pool_order == "early":
            self.set2set = P.Set2Set(self.gc_dim, processing_steps=3)
        elif pool == "set2set" and pool_order == "late":
            self.set2set = P.Set2Set(self.output_dim, processing_steps=3, num_layers=1)
            self.lin_out_2 = nn.Linear(self.output_dim * 2, self.output_dim)

    def _bn(self):
        return nn.BatchNorm1d(self.gc_dim, track_running_stats=self.batch_track_stats)

    def _pre(self, data):
        out = data.x
        for lin in self.pre_lin_list:
            out = getattr(F, self.act)(lin(out))
        return out

    def _post(self, out, data):
        s2s = self.pool == "set2set"
        pool = self.set2set if s2s else getattr(P, self.pool)
        if self.pool_order == "early":
            out = pool(out, data.batch)
        for lin in self.post_lin_list:
            out = getattr(F, self.act)(lin(out))
        out = self.lin_out(out)
        if self.pool_order == "late":
            out = pool(out, data.batch)
            if s2s:
                out = self.lin_out_2(out)
        return out.view(-1) if out.shape[1] == 1 else out


class CGCNN(_Skeleton):
    def __init__(self, data, dim1=64, dim2=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        self.conv_list = nn.ModuleList(
            [P.CGConv(self.gc_dim, data.num_edge_features, aggr="mean", batch_norm=False)
             for _ in range(gc_count)])
        self.bn_list = nn.ModuleList(
            [self._bn() for _ in range(gc_count)] if batch_norm == "True" else [])

    def forward(self, data):
        out = self._pre(data)
        for i, conv in enumerate(self.conv_list):
            out = conv(out, data.edge_index, data.edge_attr)
            if self.batch_norm == "True":
                out = self.bn_list[i](out)
            out = F.dropout(out, p=self.dropout_rate, training=self.training)
        return self._post(out, data)


class GCN(_Skeleton):
    """reference gcn.py: conv(out, edge_index, edge_weight) -> BN -> act -> dropout (gcn.py:139-152)"""

    def __init__(self, data, dim1=64, dim2=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0, **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        self.conv_list = nn.ModuleList(
            [P.GCNConv(self.gc_dim, self.gc_dim, improved=True, add_self_loops=False) for _ in range(gc_count)])
        self.bn_list = nn.ModuleList(
            [self._bn() for _ in range(gc_count)] if batch_norm == "True" else [])

    def forward(self, data):
        out = self._pre(data)
        for i, conv in enumerate(self.conv_list):
            out = conv(out, data.edge_index, data.edge_weight)
            if self.batch_norm == "True":
                out = self.bn_list[i](out)
            out = getattr(F, self.act)(out)
            out = F.dropout(out, p=self.dropout_rate, training=self.training)
        return self._post(out, data)


class SchNet(_Skeleton):
    def __init__(self, data, dim1=64, dim2=64, dim3=64, cutoff=8, pre_fc_count=1, gc_count=3,
                 post_fc_count=1, pool="global_mean_pool", pool_order="early",
                 batch_norm="True", batch_track_stats="True", act="relu", dropout_rate=0.0,
                 **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        self.conv_list = nn.ModuleList(
            [P.InteractionBlock(self.gc_dim, data.num_edge_features, dim3, cutoff)
             for _ in range(gc_count)])
        self.bn_list = nn.ModuleList(
            [self._bn() for _ in range(gc_count)] if batch_norm == "True" else [])

    def forward(self, data):
        out = self._pre(data)
        for i, conv in enumerate(self.conv_list):
            out = out + conv(out, data.edge_index, data.edge_weight, data.edge_attr)
            if self.batch_norm == "True":
                out = self.bn_list[i](out)
            out = F.dropout(out, p=self.dropout_rate, training=self.training)
        return self._post(out, data)


class MPNN(_Skeleton):
    def __init__(self, data, dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3,
                 post_fc_count=1, pool="global_mean_pool", pool_order="early",
                 batch_norm="True", batch_track_stats="True", act="relu", dropout_rate=0.0,
                 **kwargs):
        super().__init__(data, dim1, dim2, pre_fc_count, gc_count, post_fc_count, pool,
                         pool_order, batch_norm, batch_track_stats, act, dropout_rate)
        g = self.gc_dim
        self.conv_list = nn.ModuleList()
        self.gru_list = nn.ModuleList()
        for _ in range(gc_count):
            edge_net = nn.Sequential(nn.Linear(data.num_edge_features, dim3), nn.ReLU(),
                                     nn.Linear(dim3, g * g))
            self.conv_list.append(P.NNConv(g, g, edge_net, aggr="mean"))
            self.gru_list.append(nn.GRU(g, g))
        self.bn_list = nn.ModuleList(
            [self._bn() for _ in range(gc_count)] if batch_norm == "True" else [])

    def forward(self, data):
        out = self._pre(data)
        h = out.unsqueeze(0)
        for i, conv in enumerate(self.conv_list):
            m = conv(out, data.edge_index, data.edge_attr)
            if self.batch_norm == "True":
                m = self.bn_list[i](m)
            m = getattr(F, self.act)(m)
            m = F.dropout(m, p=self.dropout_rate, training=self.training)
            out, h = self.gru_list[i](m.unsqueeze(0), h)
            out = out.squeeze(0)
        return self._post(out, data)


class _MegnetMLP(nn.Module):
    """Linear -> act -> BN -> dropout, fc_layers+1 times (reference
    megnet.py:28-56)."""

    def __init__(self, in_mult, dim, act, batch_norm, batch_track_stats, dropout_rate,
                 fc_layers, list_name):
        super().__init__()
        self.act, self.batch_norm, self.dropout_rate = act, batch_norm, dropout_rate
        # reference quirk (megnet.py:21-24): MEGNet hands these sub-models an
        # already-converted bool, and `False == "False"` is False, so stats are
        # always tracked unless the *string* "False" arrives.
        track = not (batch_track_stats == "False")
        setattr(self, list_name, nn.ModuleList(
            [nn.Linear(dim * in_mult if i == 0 else dim, dim) for i in range(fc_layers + 1)]))
        self._list_name = list_name
        self.bn_list = nn.ModuleList(
            [nn.BatchNorm1d(dim, track_running_stats=track) for _ in range(fc_layers + 1)]
            if batch_norm == "True" else [])

    def _run(self, comb):
        out = comb
        for i, lin in enumerate(getattr(self, self._list_name)):
            out = getattr(F, self.act)(lin(out))
            if self.batch_norm == "True":
                out = self.bn_list[i](out)
            out = F.dropout(out, p=self.dropout_rate, training=self.training)
        return out


class Megnet_EdgeModel(_MegnetMLP):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(4, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "edge_mlp")

    def forward(self, src, dest, edge_attr, u, batch):
        return self._run(torch.cat([src, dest, edge_attr, u[batch]], dim=1))


class Megnet_NodeModel(_MegnetMLP):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(3, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "node_mlp")

    def forward(self, x, edge_index, edge_attr, u, batch):
        v_e = P.scatter_mean(edge_attr, edge_index[0, :], dim=0)
        return self._run(torch.cat([x, v_e, u[batch]], dim=1))


class Megnet_GlobalModel(_MegnetMLP):
    def __init__(self, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers=2):
        super().__init__(3, dim, act, batch_norm, batch_track_stats, dropout_rate, fc_layers, "global_mlp")

    def forward(self, x, edge_index, edge_attr, u, batch):
        u_e = P.scatter_mean(edge_attr, edge_index[0, :], dim=0)
        u_e = P.scatter_mean(u_e, batch, dim=0)
        u_v = P.scatter_mean(x, batch, dim=0)
        return self._run(torch.cat([u_e, u_v, u], dim=1))


class MEGNet(nn.Module):
    def __init__(self, data, dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3,
                 gc_fc_count=2, post_fc_count=1, pool="global_mean_pool", pool_order="early",
                 batch_norm="True", batch_track_stats="True", act="relu", dropout_rate=0.0,
                 **kwargs):
        super().__init__()
        assert gc_count > 0 and pool != "set2set"
        track = batch_track_stats != "False"
        self.batch_norm, self.pool, self.act = batch_norm, pool, act
        self.pool_reduce = {"global_mean_pool": "mean", "global_max_pool": "max",
                            "global_sum_pool": "sum"}.get(pool)
        self.pool_order, self.dropout_rate = pool_order, dropout_rate
        gc_dim = dim1 if pre_fc_count > 0 else data.num_features
        out_dim = _out_dim(data)
        self.pre_lin_list = _mlp_stack(pre_fc_count, data.num_features, dim1)

        def embed(d_in):
            return nn.Sequential(nn.Linear(d_in, dim3), nn.ReLU(), nn.Linear(dim3, dim3), nn.ReLU())

        self.e_embed_list, self.x_embed_list = nn.ModuleList(), nn.ModuleList()
        self.u_embed_list, self.conv_list = nn.ModuleList(), nn.ModuleList()
        self.bn_list = nn.ModuleList()
        for i in range(gc_count):
            self.e_embed_list.append(embed(data.num_edge_features if i == 0 else dim3))
            self.x_embed_list.append(embed(gc_dim if i == 0 else dim3))
            self.u_embed_list.append(embed(data[0].u.shape[1] if i == 0 else dim3))
            args = (dim3, act, batch_norm, track, dropout_rate, gc_fc_count)
            self.conv_list.append(P.MetaLayer(Megnet_EdgeModel(*args), Megnet_NodeModel(*args),
                                              Megnet_GlobalModel(*args)))
        post_in = dim3 * 3 if pool_order == "early" else dim3
        self.post_lin_list = _mlp_stack(post_fc_count, post_in, dim2)
        self.lin_out = nn.Linear(dim2 if post_fc_count > 0 else post_in, out_dim)

    def forward(self, data):
        out = data.x
        for lin in self.pre_lin_list:
            out = getattr(F, self.act)(lin(out))
        x, e, u = out, data.edge_attr, data.u
        for i, conv in enumerate(self.conv_list):
            e_t = self.e_embed_list[i](e)
            x_t = self.x_embed_list[i](x)
            u_t = self.u_embed_list[i](u)
            x_o, e_o, u_o = conv(x_t, data.edge_index, e_t, u_t, data.batch)
            if i == 0:  # first block adds to the embedded temporaries (megnet.py:313-315)
                x, e, u = x_o + x_t, e_o + e_t, u_o + u_t
            else:       # later blocks add to the previous state (megnet.py:334-336)
                x, e, u = x_o + x, e_o + e, u_o + u
        if self.pool_order == "early":
            x_pool = P.scatter(x, data.batch, dim=0, reduce=self.pool_reduce)
            e_pool = P.scatter(e, data.edge_index[0, :], dim=0, reduce=self.pool_reduce)
            e_pool = P.scatter(e_pool, data.batch, dim=0, reduce=self.pool_reduce)
            out = torch.cat([x_pool, e_pool, u], dim=1)
            for lin in self.post_lin_list:
                out = getattr(F, self.act)(lin(out))
            out = self.lin_out(out)
        else:
            out = x
            for lin in self.post_lin_list:
                out = getattr(F, self.act)(lin(out))
            out = self.lin_out(out)
            out = getattr(P, self.pool)(out, data.batch)
        return out.view(-1) if out.shape[1] == 1 else out

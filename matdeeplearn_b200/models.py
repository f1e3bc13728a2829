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

from . import nn as mnn
from .csr import csr_for


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
        if pool == "set2set":
            raise NotImplementedError("Set2Set readout is outside the engine's scope (SURVEY.md 8f)")
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
        self.post_lin_list = tnn.ModuleList(
            tnn.Linear(self.gc_dim if i == 0 else dim2, dim2) for i in range(post_fc_count))
        self.lin_out = tnn.Linear(dim2 if post_fc_count > 0 else self.gc_dim, _target_dim(data))

    def _add_bn(self):
        if self.batch_norm == "True":
            self.bn_list.append(tnn.BatchNorm1d(self.gc_dim, track_running_stats=self.batch_track_stats))

    def _activation(self, t):
        return getattr(F, self.act)(t)

    def _embed(self, data):
        h = data.x
        for lin in self.pre_lin_list:
            h = self._activation(lin(h))
        return h

    def _readout(self, h, data):
        pool = getattr(mnn, self.pool)
        if self.pool_order == "early":
            h = pool(h, data.batch)
        for lin in self.post_lin_list:
            h = self._activation(lin(h))
        h = self.lin_out(h)
        if self.pool_order == "late":
            h = pool(h, data.batch)
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
                h = self.bn_list[i](h)
            h = F.dropout(h, p=self.dropout_rate, training=self.training)
        return self._readout(h, data)

"""Minimal host-side containers standing in for torch_geometric's Data / Batch /
InMemoryDataset as far as the reference's models and training loop touch them.

The reference reads, from a batch:   x, edge_index, edge_attr, edge_weight,
batch, u, y  (matdeeplearn/models/cgcnn.py:124-154, schnet.py:134-143,
megnet.py:306-348, training/training.py:39-43) and, from the dataset handed to
a model constructor:   num_features, num_edge_features, data[0].y, data[0].u
(cgcnn.py:49-61,81; megnet.py:229).

Collation follows PyG's Batch.from_data_list as used through the DataLoader at
training.py:300-307: per-graph tensors concatenated along dim 0, edge_index
shifted by the running node count, `batch` = graph id per node.
"""
from __future__ import annotations

import torch


class Data:
    """One graph.  Attributes are plain tensors set by keyword."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.shape[0]

    @property
    def num_edges(self):
        return self.edge_index.shape[1]

    def keys(self):
        return [k for k, v in self.__dict__.items() if not k.startswith("_")]


class Batch(Data):
    """Block-diagonal union of graphs in the reference's (PyG) layout.

    Engine extras (never read by reference-style model code):
      _csr : cached matdeeplearn_b200.csr.GraphCSR built lazily from
             edge_index/batch on first use by a CUDA operator.
    """

    _TENSOR_KEYS = ("x", "edge_index", "edge_attr", "edge_weight", "batch", "u", "y")

    @classmethod
    def from_data_list(cls, graphs):
        xs, eis, eas, ews, bs, us, ys = [], [], [], [], [], [], []
        off = 0
        for g, d in enumerate(graphs):
            n = d.x.shape[0]
            xs.append(d.x)
            eis.append(d.edge_index + off)
            eas.append(d.edge_attr)
            ews.append(d.edge_weight)
            bs.append(torch.full((n,), g, dtype=torch.long))
            us.append(d.u.reshape(1, -1))
            ys.append(d.y.reshape(-1)[:1] if d.y.ndim <= 1 else d.y)
            off += n
        extras = {}
        if all(hasattr(d, "d_hat") for d in graphs):
            # normalised distances behind edge_attr (process.assemble_dataset keeps them): lets a
            # consumer ship 4 B/edge to the GPU and expand the Gaussian basis there
            extras["d_hat"] = torch.cat([d.d_hat for d in graphs], 0)
        out = cls(
            x=torch.cat(xs, 0),
            edge_index=torch.cat(eis, 1),
            edge_attr=torch.cat(eas, 0),
            edge_weight=torch.cat(ews, 0),
            batch=torch.cat(bs, 0),
            u=torch.cat(us, 0),
            y=torch.cat(ys, 0),
            **extras,
        )
        out.num_graphs = len(graphs)
        return out

    def to(self, device, non_blocking=False):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pin_memory(self):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.pin_memory() if torch.is_tensor(v) else v)
        return out

    def double(self):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.double() if torch.is_tensor(v) and v.is_floating_point() else v)
        return out


class GraphDataset:
    """List-of-Data with the attributes reference model constructors read."""

    def __init__(self, graphs):
        self.graphs = list(graphs)

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return GraphDataset(self.graphs[i])
        return self.graphs[i]

    @property
    def num_features(self):
        return self.graphs[0].x.shape[1]

    @property
    def num_edge_features(self):
        return self.graphs[0].edge_attr.shape[1]

    def batch(self, idx=None):
        gs = self.graphs if idx is None else [self.graphs[i] for i in idx]
        b = Batch.from_data_list(gs)
        smear = getattr(self, "smear", None)
        if smear is not None and hasattr(b, "d_hat"):
            b.smear = dict(smear)   # (start, stop, resolution, width) of the GaussianSmearing that made edge_attr
        return b

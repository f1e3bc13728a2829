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


class GaussianEdgeAttr:
    """`edge_attr` in its 4-bytes-per-edge form: the normalised distance d_hat [E] plus the parameters of the
    GaussianSmearing module that the reference applies once at dataset build (process/process.py:500-502,
    580-590): edge_attr[e, k] = exp(coeff * (d_hat[e] - offset[k])^2), offset = linspace(start, stop, resolution),
    coeff = -0.5 / ((stop - start) * width)^2.

    It stands where the reference's [E, G] tensor stands (Batch.edge_attr, passed through the model files into the
    conv layers untouched).  nn.CGConv consumes it directly -- the fused edge kernels expand the basis on the fly
    (mdl_cgconv_smear_fwd / _bwd), so the [E, G] tensor never exists in HBM; every other consumer calls
    `materialize()` (the GaussianSmearing kernel on CUDA, the module's formula on the host), cached per version
    of d_hat."""

    def __init__(self, d_hat, start=0.0, stop=1.0, resolution=50, width=0.2):
        self.d_hat = d_hat
        self.start, self.stop, self.resolution, self.width = float(start), float(stop), int(resolution), float(width)
        self._offset = None
        self._dense = None
        self._slots = None

    # ---- the tensor-like surface the reference's model code touches
    @property
    def shape(self):
        return torch.Size((self.d_hat.shape[0], self.resolution))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 2

    @property
    def dtype(self):
        return self.d_hat.dtype

    @property
    def device(self):
        return self.d_hat.device

    @property
    def is_cuda(self):
        return self.d_hat.is_cuda

    @property
    def requires_grad(self):
        return False

    @property
    def coeff(self):
        return -0.5 / ((self.stop - self.start) * self.width) ** 2

    @property
    def offset(self):
        if self._offset is None or self._offset.device != self.d_hat.device or self._offset.dtype != self.d_hat.dtype:
            self._offset = torch.linspace(self.start, self.stop, self.resolution, dtype=self.d_hat.dtype,
                                          device=self.d_hat.device)
        return self._offset

    def params(self):
        return dict(start=self.start, stop=self.stop, resolution=self.resolution, width=self.width)

    def _like(self, d_hat):
        return GaussianEdgeAttr(d_hat, **self.params())

    def to(self, *a, **kw):
        return self._like(self.d_hat.to(*a, **kw))

    def pin_memory(self):
        return self._like(self.d_hat.pin_memory())

    def double(self):
        return self._like(self.d_hat.double())

    def fusable(self):
        """True if the in-kernel expansion (two exponentials per 8 basis functions + a 7-step geometric
        recurrence, csrc/edge_dev.cuh) is safe for these parameters: either no basis value of a distance within
        one span of [start, stop] can underflow fp32 at all, or a value that underflows at the head of a chunk
        cannot grow back to anything significant within the chunk."""
        if self.resolution < 2 or self.d_hat.dtype != torch.float32:
            return False
        span = abs(self.stop - self.start)
        c, dmu = abs(self.coeff), span / (self.resolution - 1)
        return c * (2.0 * span) ** 2 <= 80.0 or 14.0 * c * dmu * 2.0 * span <= 50.0

    def slots(self, csr):
        """d_hat in the engine's slot (destination-major) order, memoised per (layout, version of d_hat)."""
        hit = self._slots
        if hit is not None and hit[0] is csr and hit[1] == self.d_hat._version:
            return hit[2]
        from .csr import gather_rows
        out = gather_rows(self.d_hat.view(-1, 1), csr.dst_eid).view(-1)
        self._slots = (csr, self.d_hat._version, out)
        return out

    def forget(self):
        """Drop the memoised expansions (CUDA-graph capture: they must be recomputed inside the graph)."""
        self._dense = None
        self._slots = None

    def materialize(self):
        """The [E, G] tensor the reference stores (same formula, same offset buffer)."""
        hit = self._dense
        if hit is not None and hit[0] == self.d_hat._version and hit[1].device == self.d_hat.device:
            return hit[1]
        if self.d_hat.is_cuda and self.d_hat.dtype == torch.float32:
            from . import functional as MF
            dense = MF.gaussian_smear(self.d_hat, self.offset, self.coeff)
        else:
            dense = torch.exp(self.coeff * (self.d_hat.view(-1, 1) - self.offset.view(1, -1)) ** 2)
        self._dense = (self.d_hat._version, dense)
        return dense


def dense_edge_attr(edge_attr):
    """edge_attr as a tensor (GaussianEdgeAttr -> its [E, G] expansion)."""
    return edge_attr.materialize() if isinstance(edge_attr, GaussianEdgeAttr) else edge_attr


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


def _tensor_like(v):
    return torch.is_tensor(v) or isinstance(v, GaussianEdgeAttr)


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

    def with_lazy_edge_attr(self):
        """A shallow copy whose edge_attr is the 4 B/edge GaussianEdgeAttr(d_hat, smear parameters) instead of the
        [E, G] tensor (batches made by process.assemble_dataset carry d_hat and `smear`)."""
        smear = getattr(self, "smear", None)
        if smear is None or not hasattr(self, "d_hat"):
            raise ValueError("batch carries no d_hat / smear parameters")
        out = Batch()
        out.__dict__.update({k: v for k, v in self.__dict__.items() if not k.startswith("_")})
        out.edge_attr = GaussianEdgeAttr(self.d_hat, **smear)
        return out

    def to(self, device, non_blocking=False):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.to(device, non_blocking=non_blocking) if _tensor_like(v) else v)
        return out

    def pin_memory(self):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.pin_memory() if _tensor_like(v) else v)
        return out

    def double(self):
        out = Batch()
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            setattr(out, k, v.double() if (torch.is_tensor(v) and v.is_floating_point()) or
                    isinstance(v, GaussianEdgeAttr) else v)
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

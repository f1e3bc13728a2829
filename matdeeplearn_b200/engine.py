"""Training-step driver: the body of the reference's train() loop
(matdeeplearn/training/training.py:37-50 -- data.to(device), zero_grad, forward,
loss, backward, optimizer.step) as one replayable CUDA graph over a
device-resident batch, plus the host-facing variant that starts from pinned
host buffers in the reference's layout.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import os
from collections import OrderedDict

from . import _lib, dist as mdist
from .csr import csr_for
from .data import Batch, GaussianEdgeAttr


class TrainStep:
    """zero_grad -> model(batch) -> loss -> backward -> [grad all-reduce] -> AdamW.

    `resident(batch)` captures the step for a batch already on the device and
    returns a callable replaying it; `from_host(batch)` runs the same step
    starting from a pinned host batch (H2D + layout build inside the call) and
    returns the loss as a Python float (D2H)."""

    def __init__(self, model, lr=2e-3, loss="l1_loss", weight_decay=1e-2):
        self.model = model
        self.flat = mdist.FlatParameters(model)
        self.flat.enable_direct()   # weight/bias gradients land in the flat buffer, nothing to concatenate
        self.loss_fn = getattr(F, loss)
        self.device = self.flat.param.device
        self.opt = mdist.FlatAdamW(self.flat, lr=lr, weight_decay=weight_decay)
        self.kernels_per_step = None
        self._graphs = {}
        self._host_graphs = OrderedDict()   # LRU: one static batch + graph pool per (N, E, B) shape
        self.max_host_graphs = 8
        self._store_graphs = {}
        # multi-GPU: the flat gradient all-reduce is captured INSIDE the step's CUDA graph when the backend is NCCL
        # (one replay per step; MDL_GRAPH_ALLREDUCE=0 keeps it as an eager call between two graphs)
        self.graph_allreduce = (mdist.is_distributed() and torch.distributed.get_backend() == "nccl"
                                and os.environ.get("MDL_GRAPH_ALLREDUCE", "1") != "0")
        if mdist.is_distributed():
            # what DDP's constructor does (reference training.py:262-266): every replica starts from rank 0's
            # parameters and buffers, whatever seed the caller built the model with
            mdist.broadcast_(self.flat.param)
            for b in model.buffers():
                mdist.broadcast_(b)
        # warm-up runs on a side stream (standard capture recipe); the resulting stream-mismatch
        # note from autograd's AccumulateGrad is expected and harmless here
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except Exception:
            pass

    # -- pieces ---------------------------------------------------------------
    def _fwd_bwd(self, batch):
        self.flat.release_grads()
        out = self.model(batch)
        loss = self.loss_fn(out, batch.y)
        loss.backward()
        self.flat.pack_grads()      # one concatenation into the flat gradient buffer
        return loss

    def _reduce(self):
        """DDP semantics: mean of the per-rank gradients.  The sum is ONE all-reduce of the flat
        buffer; the 1/world factor rides along in the optimizer kernel."""
        if mdist.is_distributed():
            torch.distributed.all_reduce(self.flat.grad, op=torch.distributed.ReduceOp.SUM)

    def _opt_step(self):
        world = torch.distributed.get_world_size() if mdist.is_distributed() else 1
        self.opt.step(grad_scale=1.0 / world)

    def _finish(self):
        self._reduce()
        self._opt_step()

    def eager(self, batch):
        loss = self._fwd_bwd(batch)
        self._finish()
        return loss

    def _capture(self, batch, pre=None):
        """Capture one step on `batch` (after `pre()`, e.g. batch assembly or the layout build, inside the graph).
        Single GPU: one graph.  Multi-GPU: one graph with the NCCL all-reduce of the flat gradient buffer captured
        between backward and AdamW (`graph_allreduce`), else backward | eager all-reduce | AdamW as two graphs.
        Returns (replay, loss tensor, kernels launched by libmdl_b200.so per step [+1 for the collective])."""
        distributed = mdist.is_distributed()
        fused = distributed and self.graph_allreduce
        n0 = _lib.launch_count()

        def capture_first(with_collective):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if pre is not None:
                    pre()
                loss_ = self._fwd_bwd(batch)
                if with_collective:
                    self._reduce()
                if with_collective or not distributed:
                    self._opt_step()
            return g, loss_

        if fused:
            try:
                g1, loss = capture_first(True)
            except Exception as exc:   # a communicator that cannot be captured: eager collective between two graphs
                import warnings
                warnings.warn(f"all-reduce could not be captured inside the step graph ({exc!r}); "
                              "falling back to an eager collective between two graphs")
                torch.cuda.synchronize()
                self.graph_allreduce = fused = False
                n0 = _lib.launch_count()
        if not fused:
            g1, loss = capture_first(False)
        g2 = None
        if distributed and not fused:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self._opt_step()
        kernels = _lib.launch_count() - n0 + (1 if distributed else 0)

        def replay():
            g1.replay()
            if g2 is not None:
                self._reduce()
                g2.replay()
            return loss

        return replay, loss, kernels

    # -- device-resident, graph-replayed ---------------------------------------
    def resident(self, batch: Batch, warmup=3):
        assert batch.x.is_cuda
        csr_for(batch.edge_index, batch.batch, num_nodes=batch.x.shape[0],
                num_graphs=getattr(batch, "num_graphs", None))  # layout built outside the graph
        snap = self._snapshot()      # warm-up steps are not training steps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.eager(batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        replay, loss, self.kernels_per_step = self._capture(batch)
        self._graphs[id(batch)] = (replay, loss, batch)
        return replay

    # -- from pinned host memory (the call a user of the reference makes) -------
    def from_host(self, host_batch: Batch, use_graph=True, expand_edge_attr=True):
        """One training step starting from a (pinned) host batch in the reference's layout.

        H2D copies of every tensor of the batch, engine layout build (CSR sort, slot
        permutation), forward, backward, [all-reduce], AdamW, loss read-back.  Everything
        after the copies is replayed from a CUDA graph cached per batch shape (N, E, B):
        the layout build is itself a handful of asynchronous kernels, so it sits inside
        the graph and is re-executed on the new indices at every replay."""
        if not use_graph:
            dev_batch = host_batch.to(self.device, non_blocking=True)
            return float(self.eager(dev_batch).item())
        B = int(getattr(host_batch, "num_graphs", host_batch.y.shape[0]))
        # If the batch carries the normalised distances its edge_attr was expanded from (batches made by
        # process.assemble_dataset do), ship those 4 B/edge and run GaussianSmearing (reference
        # process.py:580-590) on the device inside the graph instead of copying 4*G B/edge.
        smear = getattr(host_batch, "smear", None) if expand_edge_attr else None
        lazy = smear is not None and hasattr(host_batch, "d_hat")
        key = (host_batch.x.shape[0], host_batch.edge_index.shape[1], B, lazy)
        entry = self._host_graphs.get(key)
        if entry is None:
            entry = self._capture_host_step(host_batch, B, smear if lazy else None)
            self._host_graphs[key] = entry
            while len(self._host_graphs) > self.max_host_graphs:   # least recently used shape goes
                self._host_graphs.popitem(last=False)
        else:
            self._host_graphs.move_to_end(key)
        static, replay, loss, names = entry
        nbytes = 0
        for name in names:
            src = getattr(host_batch, name)
            getattr(static, name).copy_(src, non_blocking=True)
            nbytes += src.numel() * src.element_size()
        self.last_h2d_bytes = nbytes
        replay()
        return float(loss.item())

    # -- from a device-resident GraphStore (shuffled epochs at graph-replay speed) ----------
    def from_store(self, store, idx, sync=True):
        """One training step on graphs `idx` of a store.GraphStore.

        The batch is assembled on the GPU into capacity-padded buffers of fixed shape, so ONE captured
        CUDA graph (assembly kernel + forward + backward + AdamW) is replayed for every batch of the
        epoch whatever its node/edge count; the per-step host work is a 256-element prefix sum and a
        6 KB copy.  Padding rows are inert (empty segments, masked BatchNorm statistics): the step is
        the same function of the batch as `eager(store.batch(idx))`.  A batch larger than the
        capacity takes that exact eager path.  sync=False returns the device loss tensor (overwritten
        by the next step) instead of a float."""
        if getattr(self.model, "pool", None) == "set2set":
            raise NotImplementedError("padded replay: the Set2Set readout is not masked; use eager(store.batch(idx))")
        B = len(idx)
        key = (id(store), B)
        entry = self._store_graphs.get(key)
        if entry is None:
            entry = self._capture_store_step(store, B, idx)
            self._store_graphs[key] = entry
        static, replay, loss = entry
        if not store.load(static, idx):
            out = self.eager(store.batch(idx)).detach()
            return float(out.item()) if sync else out
        replay()
        return float(loss.item()) if sync else loss

    def _snapshot(self):
        bufs = [b for b in self.model.buffers()]
        return ([t.clone() for t in (self.flat.param, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.step_count)],
                [b.clone() for b in bufs])

    def _restore(self, snap):
        with torch.no_grad():
            for t, c in zip((self.flat.param, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.step_count), snap[0]):
                t.copy_(c)
            for b, c in zip(self.model.buffers(), snap[1]):
                b.copy_(c)

    def _capture_store_step(self, store, B, idx):
        # edge_attr in its 4 B/edge form whenever the store can provide it: CGConv expands the Gaussian basis inside
        # its fused kernels, so no [E, G] tensor is assembled, permuted or read
        static = store.static_batch(B, lazy=store.d_hat is not None and store.smear is not None)
        if not store.load(static, idx):
            raise RuntimeError("first batch exceeds the padded capacity; pass a typical batch first")
        snap = self._snapshot()      # the warm-up steps below must not count as training
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                store.assemble(static)
                self.eager(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        replay, loss, self.store_kernels_per_step = self._capture(static, pre=lambda: store.assemble(static))
        return static, replay, loss

    def _layout_inside_graph(self, static, B):
        from .csr import GraphCSR
        csr = GraphCSR.from_coo(static.edge_index, static.batch, num_nodes=static.x.shape[0], num_graphs=B)
        static.edge_index._mdl_csr = (static.edge_index._version, csr)  # what models' csr_for() will find
        if isinstance(static.edge_attr, GaussianEdgeAttr):
            static.edge_attr.forget()                                     # re-expand / re-permute inside the graph
        elif hasattr(static.edge_attr, "_mdl_slots"):
            del static.edge_attr._mdl_slots                               # re-permute inside the graph
        return csr

    def _capture_host_step(self, host_batch, B, smear=None):
        static = host_batch.to(self.device)
        static.num_graphs = B
        names = list(Batch._TENSOR_KEYS)
        if smear is not None:
            # edge_attr stays in its 4 B/edge form: CGConv expands the basis inside its fused kernels, any other
            # consumer materialises it (inside the graph: _layout_inside_graph drops the memoised copies)
            names = [n for n in names if n != "edge_attr"] + ["d_hat"]
            static.edge_attr = GaussianEdgeAttr(static.d_hat, **smear)
        snap = self._snapshot()      # the warm-up steps below must not count as training
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._layout_inside_graph(static, B)
                self.eager(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        replay, loss, _ = self._capture(static, pre=lambda: self._layout_inside_graph(static, B))
        return static, replay, loss, names

"""torch.autograd bindings of the C-ABI kernels (forward AND backward go through
libmdl_b200.so; there is no eager/PyTorch fallback for the fused ops).

Dense node-level GEMMs (the [N,C]x[C,4C] projections and their weight
gradients) are plain library GEMMs issued through torch (cuBLAS); everything of
edge size is done inside the fused kernels and never materialised.
"""
from __future__ import annotations

import torch

from . import _lib
from .csr import GraphCSR


# ----------------------------------------------------------------------------
# segmented reduce (pools, scatter_mean by source, node -> graph)
# ----------------------------------------------------------------------------
# ----------------------------------------------------------------------------
# direct gradient delivery (dist.FlatParameters.enable_direct)
# ----------------------------------------------------------------------------
def _grad_dest(p):
    """Where parameter p's gradient lives when the engine owns a flat gradient buffer and p has not
    been written this step; None otherwise (the Function then returns the gradient to autograd)."""
    if p is None:
        return None
    d = getattr(p, "_mdl_grad_dest", None)
    if d is None or getattr(p, "_mdl_written", False):
        return None
    return d


_side_streams = {}


def _side_stream(device):
    """One auxiliary stream per device for work that is independent of the main chain inside a backward (forked and
    joined with stream waits, which CUDA-graph capture records as parallel branches)."""
    key = (device.type, device.index)
    st = _side_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[key] = st
    return st


def _ptr_off(t, floats=0):
    return None if t is None else t.data_ptr() + 4 * floats


def _wgrad_map(block_rows, ldw, w_ptrs, b_ptrs):
    m = _lib.WgradOutC()
    m.block_rows, m.num_blocks, m.ldw = block_rows, len(w_ptrs), ldw
    for k, (w, b) in enumerate(zip(w_ptrs, b_ptrs)):
        m.w[k] = w
        m.b[k] = b
    return m


def linear_wgrad_into(x, g, wmap):
    """dW = g^T x, db = g.sum(0) delivered through the block map `wmap` (mdl_linear_wgrad)."""
    import ctypes
    lib = _lib.load()
    x, g = x.contiguous(), g.contiguous()
    N, I = x.shape
    O = g.shape[1]
    need = int(lib.mdl_linear_wgrad_workspace_bytes(N, I, O))
    ws = torch.empty(max(need, 4), dtype=torch.uint8, device=x.device)
    rc = lib.mdl_linear_wgrad(_lib.ptr(x), _lib.ptr(g), N, I, O, ctypes.byref(wmap), _lib.ptr(ws), ws.numel(),
                              _lib.stream())
    _lib.check(rc, "mdl_linear_wgrad")


# rows from which the tcgen05 kernels for long batches (csrc/wgrad_tc.cu, linear_tc.cu) take over from the library
# GEMMs: edge-level layers; node-level ones (a few thousand rows) stay where they were
_WGRAD_TC_MIN_ROWS = 16384


def _linear_tc(x, w, ldn, ldk, bias, K, N, act=0):
    """act(x [R,K] . B^T (+ bias)) with B[n][k] = w[n*ldn + k*ldk] on the tcgen05 kernel (mdl_linear_tc);
    act 0 = none, 1 = relu (applied in the kernel's epilogue)."""
    x = x.contiguous()
    y = torch.empty((x.shape[0], N), dtype=torch.float32, device=x.device)
    rc = _lib.load().mdl_linear_tc(_lib.ptr(x), _lib.ptr(w), ldn, ldk, _lib.ptr(bias), _lib.ptr(y), x.shape[0], K, N, act,
                                   _lib.stream())
    _lib.check(rc, "mdl_linear_tc")
    return y


def _tc_ok(R, K, N):
    return bool(_lib.load().mdl_linear_tc_supported(int(R), int(K), int(N)))


class LinearFn(torch.autograd.Function):
    """y = x W^T + b: library GEMMs for y and dx; dW and db come from mdl_linear_wgrad -- delivered straight into
    the flat gradient buffer when the engine owns one (direct delivery), and for long batches (edge-level MLPs:
    a K = E contraction + a column sum on the reference's path) also without one."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu=False):
        ctx.wb = (weight, bias)
        ctx.relu = relu
        O, I = weight.shape
        if weight.is_contiguous() and _tc_ok(x.shape[0], I, O):
            y = _linear_tc(x, weight, I, 1, bias, I, O, 1 if relu else 0)   # long batch: tcgen05 (3xTF32), ReLU in the epilogue
        else:
            y = torch.nn.functional.linear(x, weight, bias)
            if relu:
                y = torch.relu_(y)
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        wparam, bparam = ctx.wb
        g = g.contiguous()
        if ctx.relu:
            g = torch.ops.aten.threshold_backward(g, y, 0.0)         # g * [y > 0]
        O_, I_ = weight.shape
        dx = None
        if ctx.needs_input_grad[0]:
            if weight.is_contiguous() and _tc_ok(g.shape[0], O_, I_):
                dx = _linear_tc(g, weight, 1, I_, None, O_, I_)     # dx = g W on tcgen05
            else:
                dx = g.mm(weight)
        wd = _grad_dest(wparam)
        bd = _grad_dest(bparam) if bparam is not None else None
        O, I = weight.shape
        if wd is not None and (bparam is None or bd is not None) and x.shape[0] > 0:
            linear_wgrad_into(x, g, _wgrad_map(O, I, [_ptr_off(wd)], [_ptr_off(bd)]))
            wparam._mdl_written = True
            if bparam is not None:
                bparam._mdl_written = True
            return dx, None, None, None
        if x.shape[0] >= 2048 and I <= 256 and I >= 8 and O >= 8:
            dW = torch.empty_like(weight)
            db = torch.empty(O, dtype=weight.dtype, device=weight.device) if bparam is not None else None
            linear_wgrad_into(x, g, _wgrad_map(O, I, [_ptr_off(dW)], [_ptr_off(db)]))
            return dx, dW, db, None
        return dx, g.t().mm(x), (g.sum(0) if bparam is not None else None), None


def linear(x, weight, bias=None, relu=False):
    """torch.nn.functional.linear (followed by ReLU if `relu`); routed through LinearFn when the engine delivers
    gradients directly or the batch is long enough for the tensor-core kernels (the ReLU then runs in the dense
    kernel's epilogue and its mask is folded into the backward's first pass)."""
    if (x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and weight.requires_grad
            and (getattr(weight, "_mdl_grad_dest", None) is not None or x.shape[0] >= _WGRAD_TC_MIN_ROWS)):
        return LinearFn.apply(x, weight, bias, relu)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(y) if relu else y


def linear_wgrad_rs_into(x, g, rowscale, wmap):
    """dW = (diag(rowscale) g)^T x, db = sum_r rowscale[r] g[r] through the block map (mdl_linear_wgrad_rs)."""
    import ctypes
    lib = _lib.load()
    x, g = x.contiguous(), g.contiguous()
    N, I = x.shape
    O = g.shape[1]
    need = int(lib.mdl_linear_wgrad_workspace_bytes(N, I, O))
    ws = torch.empty(max(need, 4), dtype=torch.uint8, device=x.device)
    rc = lib.mdl_linear_wgrad_rs(_lib.ptr(x), _lib.ptr(g), _lib.ptr(rowscale), N, I, O, ctypes.byref(wmap), _lib.ptr(ws),
                                 ws.numel(), _lib.stream())
    _lib.check(rc, "mdl_linear_wgrad_rs")


class EdgeMLP2Fn(torch.autograd.Function):
    """Y = (act1(X W1^T + b1) W2^T + b2) * rowscale[:, None] on the fused tcgen05 kernel (mdl_edge_mlp2_fwd / _bwd);
    X (edge_attr) and rowscale (the cosine cutoff) are data: no gradient flows into them."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, rowscale, act1):
        lib = _lib.load()
        x = x.contiguous()
        E, G = x.shape
        H, O = w1.shape[0], w2.shape[0]
        y = torch.empty((E, O), dtype=torch.float32, device=x.device)
        need_bwd = any(ctx.needs_input_grad[1:5])
        t1 = torch.empty((E, H), dtype=torch.float32, device=x.device) if need_bwd else None
        rs = rowscale.contiguous() if rowscale is not None else None
        rc = lib.mdl_edge_mlp2_fwd(_lib.ptr(x), _lib.ptr(w1.contiguous()), _lib.ptr(b1), _lib.ptr(w2.contiguous()),
                                   _lib.ptr(b2), _lib.ptr(rs), _lib.ptr(y), _lib.ptr(t1), E, G, H, O, act1, 0, _lib.stream())
        _lib.check(rc, "mdl_edge_mlp2_fwd")
        ctx.save_for_backward(x, t1, rs, w2)
        ctx.act1 = act1
        ctx.params = (w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, t1, rs, w2 = ctx.saved_tensors
        w1p, b1p, w2p, b2p = ctx.params
        g = g.contiguous()
        if g.data_ptr() % 32:          # 256-bit loads in the kernel
            g = g.clone()
        E, H = t1.shape
        O = g.shape[1]
        G = x.shape[1]
        dp1 = torch.empty_like(t1)
        rc = lib.mdl_edge_mlp2_bwd(_lib.ptr(g), _lib.ptr(rs), _lib.ptr(w2.contiguous()), _lib.ptr(t1), _lib.ptr(dp1), E, H, O,
                                   ctx.act1, _lib.stream())
        _lib.check(rc, "mdl_edge_mlp2_bwd")

        def deliver(xin, gin, scale, wparam, bparam, rows_out, rows_in):
            wd = _grad_dest(wparam)
            bd = _grad_dest(bparam) if bparam is not None else None
            direct = wd is not None and (bparam is None or bd is not None)
            dW = None if direct else torch.empty((rows_out, rows_in), dtype=torch.float32, device=g.device)
            db = None if (direct or bparam is None) else torch.empty(rows_out, dtype=torch.float32, device=g.device)
            wmap = _wgrad_map(rows_out, rows_in, [_ptr_off(wd if direct else dW)], [_ptr_off(bd if direct else db)])
            if scale is not None:
                linear_wgrad_rs_into(xin, gin, scale, wmap)
            else:
                linear_wgrad_into(xin, gin, wmap)
            if direct:
                wparam._mdl_written = True
                if bparam is not None:
                    bparam._mdl_written = True
            return dW, db

        dW2, db2 = deliver(t1, g, rs, w2p, b2p, O, H)      # dW2 = (g * rs)^T T1
        dW1, db1 = deliver(x, dp1, None, w1p, b1p, H, G)   # dW1 = dPre1^T X
        return None, dW1, db1, dW2, db2, None, None


def edge_mlp2_supported(x, w1, w2):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] >= 2048
            and not x.requires_grad
            and bool(_lib.load().mdl_edge_mlp2_supported(int(x.shape[1]), int(w1.shape[0]), int(w2.shape[0])))
            and w2.shape[1] == w1.shape[0])


def edge_mlp2(x, w1, b1, w2, b2, rowscale, act1="ssp"):
    """(act1(x W1^T + b1) W2^T + b2) * rowscale[:, None]; act1 in {"ssp" (shifted softplus), "relu"}."""
    return EdgeMLP2Fn.apply(x, w1, b1, w2, b2, rowscale, {"ssp": 0, "relu": 1}[act1])


def apply_mlp(seq, x):
    """Run an nn.Sequential (or a single module) with every nn.Linear going through `linear` above; other
    layers (activations, BatchNorm) are called as they are."""
    mods = list(seq) if isinstance(seq, torch.nn.Sequential) else [seq]
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Linear):
            fuse = i + 1 < len(mods) and type(mods[i + 1]) is torch.nn.ReLU    # Linear -> ReLU: one kernel
            x = linear(x, m.weight, m.bias, relu=fuse)
            i += 2 if fuse else 1
        else:
            x = m(x)
            i += 1
    return x


class SegmentReduceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, ptr, perm, reduce, num_segments):
        lib = _lib.load()
        src = src.contiguous()
        if src.dtype != torch.float32:
            raise RuntimeError("segment_reduce: fp32 only")
        rows = src.shape[0]
        width = 1
        for d in src.shape[1:]:
            width *= int(d)
        out = torch.empty((num_segments,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        red = _lib.REDUCE[reduce]
        argmax = None
        if red == 2 and ctx.needs_input_grad[0]:
            argmax = torch.empty((num_segments, width), dtype=torch.int32, device=src.device)
        rc = lib.mdl_segment_reduce_fwd(_lib.ptr(src), _lib.ptr(ptr), _lib.ptr(perm), _lib.ptr(out),
                                        _lib.ptr(argmax), num_segments, width, red, _lib.stream())
        _lib.check(rc, "mdl_segment_reduce_fwd")
        ctx.save_for_backward(ptr, perm, argmax)
        ctx.meta = (red, rows, width, num_segments, tuple(src.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        ptr, perm, argmax = ctx.saved_tensors
        red, rows, width, S, shape = ctx.meta
        g = g.contiguous()
        # with a permutation, rows no segment refers to (capacity padding) are not written by the kernel
        gsrc = (torch.zeros if perm is not None else torch.empty)(shape, dtype=g.dtype, device=g.device)
        rc = _lib.load().mdl_segment_reduce_bwd(_lib.ptr(g), _lib.ptr(ptr), _lib.ptr(perm),
                                                _lib.ptr(argmax), _lib.ptr(gsrc), S, rows, width,
                                                red, _lib.stream())
        _lib.check(rc, "mdl_segment_reduce_bwd")
        return gsrc, None, None, None, None


def segment_reduce(src, ptr, perm=None, reduce="mean"):
    """out[s] = reduce(src[perm[r]] for r in [ptr[s], ptr[s+1]));  ptr int32 [S+1]."""
    return SegmentReduceFn.apply(src, ptr, perm, reduce, ptr.shape[0] - 1)


class ExpandBySegmentFn(torch.autograd.Function):
    """out[r] = src[s] for every row r of segment s (u[batch] in the reference: megnet.py:55, 99); the backward is
    the segmented sum kernel over the same segments instead of torch's sort-based index_put."""

    @staticmethod
    def forward(ctx, src, index32, ptr):
        from .csr import gather_rows
        ctx.save_for_backward(ptr)
        ctx.rows = src.shape[0]
        return gather_rows(src, index32)

    @staticmethod
    def backward(ctx, g):
        (ptr,) = ctx.saved_tensors
        gs = segment_reduce(g.contiguous(), ptr, None, "sum")
        if gs.shape[0] < ctx.rows:   # capacity-padded batch: src carries extra rows that only padding rows read
            gs = torch.cat([gs, gs.new_zeros((ctx.rows - gs.shape[0],) + tuple(gs.shape[1:]))], 0)
        return gs, None, None


def expand_by_segment(src, index, ptr):
    """src[index] for a sorted `index` whose segments `ptr` (int32 [S+1]) describes (S = src.shape[0])."""
    i32 = getattr(index, "_mdl_i32", None)
    if i32 is None or i32[0] != index._version:
        i32 = (index._version, index.to(torch.int32))
        try:
            index._mdl_i32 = i32
        except Exception:
            pass
    return ExpandBySegmentFn.apply(src.contiguous(), i32[1], ptr)


# ----------------------------------------------------------------------------
# BatchNorm1d (training mode) with a device-side row count
# ----------------------------------------------------------------------------
class MaskedBatchNormFn(torch.autograd.Function):
    """mdl_batchnorm_fwd / _bwd: statistics over the first n_valid[0] rows (None: all rows)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, n_valid, momentum, eps, ws):
        lib = _lib.load()
        x = x.contiguous()
        if x.dim() != 2 or x.dtype != torch.float32:
            raise RuntimeError("masked batch norm: x must be fp32 [N, C]")
        N, C = x.shape
        out = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        rc = lib.mdl_batchnorm_fwd(_lib.ptr(x), _lib.ptr(n_valid), N, C, _lib.ptr(weight), _lib.ptr(bias),
                                   _lib.ptr(running_mean), _lib.ptr(running_var), float(momentum), float(eps),
                                   _lib.ptr(out), _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(ws), ws.numel(),
                                   _lib.stream())
        _lib.check(rc, "mdl_batchnorm_fwd")
        ctx.save_for_backward(x, weight, mean, invstd, n_valid, ws)
        ctx.has_bias = bias is not None
        ctx.wb = (weight, bias)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, mean, invstd, n_valid, ws = ctx.saved_tensors
        N, C = x.shape
        g = g.contiguous()
        gx = torch.empty_like(x)
        wparam, bparam = ctx.wb
        wd = _grad_dest(wparam) if weight is not None else None
        bd = _grad_dest(bparam) if ctx.has_bias else None
        direct = (weight is None or wd is not None) and (not ctx.has_bias or bd is not None)
        if direct:
            gw, gb = wd, bd
        else:
            gw = torch.empty(C, dtype=torch.float32, device=x.device) if weight is not None else None
            gb = torch.empty(C, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        rc = _lib.load().mdl_batchnorm_bwd(_lib.ptr(g), _lib.ptr(x), _lib.ptr(n_valid), N, C, _lib.ptr(weight),
                                           _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(gx), _lib.ptr(gw),
                                           _lib.ptr(gb), _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "mdl_batchnorm_bwd")
        if direct:
            for prm in (wparam, bparam):
                if prm is not None:
                    prm._mdl_written = True
            return gx, None, None, None, None, None, None, None, None
        return gx, gw, gb, None, None, None, None, None, None


def masked_batch_norm(bn, x, n_valid=None):
    """Apply torch.nn.BatchNorm1d module `bn` to x[N,C] with batch statistics over the first
    n_valid[0] rows (device int32 tensor).  Eval mode is row-wise and goes through the module."""
    if not bn.training:
        return bn(x)
    if bn.momentum is None:
        raise NotImplementedError("masked batch norm: cumulative-average momentum is not supported")
    N, C = x.shape
    # Sized ONCE for the saturated grid (the row count only changes how many CTAs write partials), so the
    # buffer is never reallocated: CUDA graphs captured earlier keep a valid pointer whatever N comes later.
    ws = getattr(bn, "_mdl_ws", None)
    if ws is None or ws.device != x.device:
        need = int(_lib.load().mdl_batchnorm_workspace_bytes(1 << 40, C))
        ws = torch.zeros(need, dtype=torch.uint8, device=x.device)
        bn._mdl_ws = ws
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return MaskedBatchNormFn.apply(x, bn.weight, bn.bias, rm, rv, n_valid, bn.momentum, bn.eps, ws)


# ----------------------------------------------------------------------------
# GaussianSmearing
# ----------------------------------------------------------------------------
def gaussian_smear(dist, offset, coeff, out=None):
    dist = dist.contiguous()
    if out is None:
        out = torch.empty((dist.shape[0], offset.shape[0]), dtype=torch.float32, device=dist.device)
    rc = _lib.load().mdl_gaussian_smear(_lib.ptr(dist), _lib.ptr(offset.contiguous()), _lib.ptr(out),
                                        dist.shape[0], offset.shape[0], float(coeff), _lib.stream())
    _lib.check(rc, "mdl_gaussian_smear")
    return out


# ----------------------------------------------------------------------------
# CGConv
# ----------------------------------------------------------------------------
class CGConvFn(torch.autograd.Function):
    """x[N,C], lin_f/lin_s weight [C, 2C+G] + bias [C], edge_attr in slot order.

    smear = None: ea_slots is the materialised edge_attr [E,G].  smear = (offset [G], coeff): the smearing-fused
    form -- ea_slots is the normalised distance d_hat [E] in slot order and the kernels expand
    exp(coeff * (d_hat - offset)^2) themselves (mdl_cgconv_smear_fwd / _bwd; reference process.py:580-590)."""

    @staticmethod
    def forward(ctx, x, w_f, b_f, w_s, b_s, ea_slots, csr: GraphCSR, reduce, smear=None):
        lib = _lib.load()
        x = x.contiguous()
        N, C = x.shape
        G = ea_slots.shape[1] if smear is None else smear[0].shape[0]
        assert w_f.shape == (C, 2 * C + G) and w_s.shape == (C, 2 * C + G)
        # column split of lin(cat[x_i, x_j, e]) = W_i x_i + W_j x_j + W_e e + b
        Wn = torch.empty((4 * C, C), dtype=torch.float32, device=x.device)    # rows P_f | P_s | Q_f | Q_s
        bias = torch.empty(4 * C, dtype=torch.float32, device=x.device)
        WeT = torch.empty((G, 2 * C), dtype=torch.float32, device=x.device)
        w_fc, w_sc = w_f.contiguous(), w_s.contiguous()
        rc = lib.mdl_cgconv_pack_weights(_lib.ptr(w_fc), _lib.ptr(b_f), _lib.ptr(w_sc), _lib.ptr(b_s), C, G,
                                         _lib.ptr(Wn), _lib.ptr(bias), _lib.ptr(WeT), _lib.stream())
        _lib.check(rc, "mdl_cgconv_pack_weights")
        PQ = torch.addmm(bias, x, Wn.t())                                              # [N, 4C]
        out = torch.empty_like(x)
        if smear is None:
            rc = lib.mdl_cgconv_fwd(_lib.ptr(x), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(WeT),
                                    _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src), _lib.ptr(csr.dst_dst),
                                    _lib.ptr(csr.inv_deg_dst), _lib.ptr(out), N, csr.E, C, G,
                                    _lib.REDUCE[reduce], _lib.stream())
            _lib.check(rc, "mdl_cgconv_fwd")
        else:
            rc = lib.mdl_cgconv_smear_fwd(_lib.ptr(x), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(smear[0]),
                                          float(smear[1]), _lib.ptr(WeT), _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src),
                                          _lib.ptr(csr.dst_dst), _lib.ptr(csr.inv_deg_dst), _lib.ptr(out), N, csr.E,
                                          C, G, _lib.REDUCE[reduce], _lib.stream())
            _lib.check(rc, "mdl_cgconv_smear_fwd")
        ctx.save_for_backward(x, PQ, Wn, WeT, ea_slots)
        ctx.csr, ctx.reduce, ctx.smear = csr, reduce, smear
        ctx.has_bias = (b_f is not None, b_s is not None)
        ctx.params = (w_f, b_f, w_s, b_s)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, PQ, Wn, WeT, ea_slots = ctx.saved_tensors
        csr = ctx.csr
        N, C = x.shape
        smear = ctx.smear
        G = WeT.shape[0]
        g = g.contiguous()
        # mean aggregation: d(mean)/d(message) = 1/deg(dst); applied once per node here instead of
        # once per edge inside the kernel
        gk = g * csr.inv_deg_dst.unsqueeze(1) if ctx.reduce == "mean" else g
        dPQ = torch.empty_like(PQ)
        dWeT = torch.empty_like(WeT)
        ws_bytes = lib.mdl_cgconv_workspace_bytes(N, csr.E, C, G)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        if smear is None:
            rc = lib.mdl_cgconv_bwd(_lib.ptr(gk), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(WeT),
                                    _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src), _lib.ptr(csr.dst_dst),
                                    _lib.ptr(csr.src_ptr), _lib.ptr(csr.src_slot),
                                    _lib.ptr(csr.inv_deg_dst), _lib.ptr(dPQ), _lib.ptr(dWeT), N, csr.E,
                                    C, G, _lib.REDUCE[ctx.reduce], _lib.ptr(ws), ws_bytes, _lib.stream())
            _lib.check(rc, "mdl_cgconv_bwd")
        else:
            rc = lib.mdl_cgconv_smear_bwd(_lib.ptr(gk), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(smear[0]),
                                          float(smear[1]), _lib.ptr(WeT), _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src),
                                          _lib.ptr(csr.dst_dst), _lib.ptr(csr.inv_deg_dst), _lib.ptr(dPQ),
                                          _lib.ptr(dWeT), N, csr.E, C, G, _lib.REDUCE[ctx.reduce], _lib.ptr(ws),
                                          ws_bytes, _lib.stream())
            _lib.check(rc, "mdl_cgconv_smear_bwd")
        w_f, b_f, w_s, b_s = ctx.params
        dwf, dws = _grad_dest(w_f), _grad_dest(w_s)
        dbf = _grad_dest(b_f) if ctx.has_bias[0] else None
        dbs = _grad_dest(b_s) if ctx.has_bias[1] else None
        if (dwf is not None and dws is not None and (not ctx.has_bias[0] or dbf is not None)
                and (not ctx.has_bias[1] or dbs is not None)):
            # direct delivery: [P_f | P_s | Q_f | Q_s] column blocks of dPQ^T x, the bias sums and the
            # transposed edge block go straight into lin_f / lin_s .weight / .bias gradient storage.  They depend only
            # on dPQ / dWeT, not on dx: they run on a side stream next to the dx GEMM (small kernels that do not fill
            # the GPU one by one; MDL_BWD_OVERLAP=0 keeps everything on one stream)
            import ctypes
            import os
            G = dWeT.shape[0]
            ld = 2 * C + G
            overlap = os.environ.get("MDL_BWD_OVERLAP", "1") != "0" and ctx.needs_input_grad[0]
            cur = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device) if overlap else cur
            if overlap:
                side.wait_stream(cur)
            with torch.cuda.stream(side):
                linear_wgrad_into(x, dPQ, _wgrad_map(
                    C, ld, [_ptr_off(dwf), _ptr_off(dws), _ptr_off(dwf, C), _ptr_off(dws, C)],
                    [_ptr_off(dbf), _ptr_off(dbs), None, None]))
                m2 = _wgrad_map(C, ld, [_ptr_off(dwf, 2 * C), _ptr_off(dws, 2 * C)], [None, None])
                rc = _lib.load().mdl_copy_mapped(_lib.ptr(dWeT), G, 2 * C, 1, ctypes.byref(m2), _lib.stream())
                _lib.check(rc, "mdl_copy_mapped")
            dx = torch.addmm(g, dPQ, Wn) if ctx.needs_input_grad[0] else None  # residual + projections
            if overlap:
                cur.wait_stream(side)
            for prm in (w_f, w_s) + ((b_f,) if ctx.has_bias[0] else ()) + ((b_s,) if ctx.has_bias[1] else ()):
                prm._mdl_written = True
            return dx, None, None, None, None, None, None, None, None
        dx = torch.addmm(g, dPQ, Wn) if ctx.needs_input_grad[0] else None  # residual + projections
        dWn = dPQ.t().mm(x)                                                 # [4C, C]
        db = dPQ[:, :2 * C].sum(0)
        return CGConvFn._finish(ctx, dx, dWn, db, dWeT, C)

    @staticmethod
    def _finish(ctx, dx, dWn, db, dWeT, C):
        dWe = dWeT.t()                                                      # [2C, G]
        dw_f = torch.cat([dWn[0:C], dWn[2 * C:3 * C], dWe[0:C]], 1)
        dw_s = torch.cat([dWn[C:2 * C], dWn[3 * C:4 * C], dWe[C:2 * C]], 1)
        db_f = db[:C] if ctx.has_bias[0] else None
        db_s = db[C:] if ctx.has_bias[1] else None
        return dx, dw_f, db_f, dw_s, db_s, None, None, None, None


def cgconv(x, w_f, b_f, w_s, b_s, ea_slots, csr, reduce="mean", smear=None):
    return CGConvFn.apply(x, w_f, b_f, w_s, b_s, ea_slots, csr, reduce, smear)


def cgconv_smear_supported(C, G):
    return bool(_lib.load().mdl_cgconv_smear_supported(int(C), int(G)))


# ----------------------------------------------------------------------------
# CFConv aggregate:  out[i] = sum_{e: dst(e)=i} h[src(e)] * W[e]
# ----------------------------------------------------------------------------
class CFConvAggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, W, csr: GraphCSR):
        lib = _lib.load()
        h, W = h.contiguous(), W.contiguous()
        N, F_ = h.shape
        out = torch.empty_like(h)
        rc = lib.mdl_spmm_edge(_lib.ptr(h), _lib.ptr(W), _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src),
                               _lib.ptr(csr.dst_eid), _lib.ptr(out), N, F_, _lib.stream())
        _lib.check(rc, "mdl_spmm_edge")
        ctx.save_for_backward(h, W)
        ctx.csr = csr
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        h, W = ctx.saved_tensors
        csr = ctx.csr
        g = g.contiguous()
        N, F_ = h.shape
        dh = dW = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty_like(h)
            rc = lib.mdl_spmm_edge(_lib.ptr(g), _lib.ptr(W), _lib.ptr(csr.src_ptr),
                                   _lib.ptr(csr.source_order_nbr()), _lib.ptr(csr.source_order_eid()),
                                   _lib.ptr(dh), N, F_, _lib.stream())
            _lib.check(rc, "mdl_spmm_edge(bwd)")
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)
            rc = lib.mdl_edge_mul(_lib.ptr(g), _lib.ptr(h), _lib.ptr(csr.dst_dst), _lib.ptr(csr.dst_src),
                                  _lib.ptr(csr.dst_eid), _lib.ptr(dW), csr.E, F_, _lib.stream())
            _lib.check(rc, "mdl_edge_mul")
        return dh, dW, None


def cfconv_aggregate(h, W, csr):
    return CFConvAggFn.apply(h, W, csr)


class SpmmScalarFn(torch.autograd.Function):
    """out[i] = sum_{e: dst(e)=i} coef[e] * h[src(e)]  (coef in reference edge order; mdl_spmm_edge_scalar / mdl_edge_dot)."""

    @staticmethod
    def forward(ctx, h, coef, csr: GraphCSR):
        lib = _lib.load()
        h, coef = h.contiguous(), coef.contiguous()
        N, F_ = h.shape
        out = torch.empty_like(h)
        rc = lib.mdl_spmm_edge_scalar(_lib.ptr(h), _lib.ptr(coef), _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src),
                                      _lib.ptr(csr.dst_eid), _lib.ptr(out), N, F_, _lib.stream())
        _lib.check(rc, "mdl_spmm_edge_scalar")
        ctx.save_for_backward(h, coef)
        ctx.csr = csr
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        h, coef = ctx.saved_tensors
        csr = ctx.csr
        g = g.contiguous()
        N, F_ = h.shape
        dh = dc = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty_like(h)
            rc = lib.mdl_spmm_edge_scalar(_lib.ptr(g), _lib.ptr(coef), _lib.ptr(csr.src_ptr),
                                          _lib.ptr(csr.source_order_nbr()), _lib.ptr(csr.source_order_eid()),
                                          _lib.ptr(dh), N, F_, _lib.stream())
            _lib.check(rc, "mdl_spmm_edge_scalar(bwd)")
        if ctx.needs_input_grad[1]:
            dc = torch.empty_like(coef)
            rc = lib.mdl_edge_dot(_lib.ptr(g), _lib.ptr(h), _lib.ptr(csr.dst_dst), _lib.ptr(csr.dst_src),
                                  _lib.ptr(csr.dst_eid), _lib.ptr(dc), csr.E, F_, _lib.stream())
            _lib.check(rc, "mdl_edge_dot")
        return dh, dc, None


def spmm_scalar(h, coef, csr):
    return SpmmScalarFn.apply(h, coef, csr)


# ----------------------------------------------------------------------------
# NNConv message:  m[e] = sum_k hid[e,k] * XT[src(e),k,:] + XB[src(e),:]
# ----------------------------------------------------------------------------
class NNConvMsgFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hid, XT, XB, csr: GraphCSR):
        lib = _lib.load()
        hid, XT, XB = hid.contiguous(), XT.contiguous(), XB.contiguous()
        E, K = hid.shape
        N, O = XB.shape
        assert XT.shape == (N, K * O)
        m = torch.empty((E, O), dtype=hid.dtype, device=hid.device)
        rc = lib.mdl_nnconv_msg_fwd(_lib.ptr(hid), _lib.ptr(XT), _lib.ptr(XB), _lib.ptr(csr.src_ptr),
                                    _lib.ptr(csr.source_order_eid()), _lib.ptr(m), N, K, O, _lib.stream())
        _lib.check(rc, "mdl_nnconv_msg_fwd")
        ctx.save_for_backward(hid, XT)
        ctx.csr, ctx.dims = csr, (N, K, O)
        return m

    @staticmethod
    def backward(ctx, dm):
        lib = _lib.load()
        hid, XT = ctx.saved_tensors
        csr = ctx.csr
        N, K, O = ctx.dims
        dm = dm.contiguous()
        dhid = torch.zeros_like(hid)     # rows no segment refers to (capacity padding) must carry a zero gradient
        dXT = torch.empty_like(XT)
        dXB = torch.empty((N, O), dtype=hid.dtype, device=hid.device)
        rc = lib.mdl_nnconv_msg_bwd(_lib.ptr(hid), _lib.ptr(XT), _lib.ptr(dm), _lib.ptr(csr.src_ptr),
                                    _lib.ptr(csr.source_order_eid()), _lib.ptr(dhid), _lib.ptr(dXT),
                                    _lib.ptr(dXB), N, K, O, _lib.stream())
        _lib.check(rc, "mdl_nnconv_msg_bwd")
        return dhid, dXT, dXB, None


def nnconv_message(hid, XT, XB, csr):
    return NNConvMsgFn.apply(hid, XT, XB, csr)


# ----------------------------------------------------------------------------
# MEGNet edge update, first Linear on cat[x[row], x[col], e, u[batch[row]]] with split weights
# ----------------------------------------------------------------------------
class EdgeGatherAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, A, B, U, bias, edge_index, batch, csr: GraphCSR, relu):
        lib = _lib.load()
        base, A, B = base.contiguous(), A.contiguous(), B.contiguous()
        U = U.contiguous() if U is not None else None
        E, D = base.shape
        out = torch.empty_like(base)
        row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
        rc = lib.mdl_edge_gather_add(_lib.ptr(base), _lib.ptr(A), _lib.ptr(B), _lib.ptr(U), _lib.ptr(row),
                                     _lib.ptr(col), _lib.ptr(batch), _lib.ptr(bias), _lib.ptr(out), E, D,
                                     1 if relu else 0, _lib.stream())
        _lib.check(rc, "mdl_edge_gather_add")
        ctx.save_for_backward(out if relu else None)
        ctx.csr, ctx.relu = csr, relu
        ctx.has = (U is not None, bias is not None)
        ctx.u_rows = U.shape[0] if U is not None else 0
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        csr = ctx.csr
        dpre = torch.ops.aten.threshold_backward(g.contiguous(), out, 0.0) if ctx.relu else g   # g * [out > 0], one pass
        dpre = dpre.contiguous()
        dA = segment_reduce(dpre, csr.src_ptr, csr.source_order_eid(), "sum")   # edges leaving each node
        dB = segment_reduce(dpre, csr.dst_ptr, csr.dst_eid, "sum")              # edges entering each node
        dU = segment_reduce(dA, csr.graph_ptr, None, "sum") if (ctx.has[0] or ctx.has[1]) else None
        # column sums over E rows = column sums of the per-graph sums (B rows): no pass over the edge tensor
        dbias = (dU.sum(0) if dU is not None else dpre.sum(0)) if ctx.has[1] else None
        if ctx.has[0] and dU.shape[0] < ctx.u_rows:   # capacity-padded batch: U carries a zero row for the padded edges
            dU_full = torch.cat([dU, dU.new_zeros(ctx.u_rows - dU.shape[0], dU.shape[1])], 0)
        else:
            dU_full = dU
        return dpre, dA, dB, (dU_full if ctx.has[0] else None), dbias, None, None, None, None


def edge_gather_add(base, A, B, U, bias, edge_index, batch, csr, relu):
    return EdgeGatherAddFn.apply(base, A, B, U, bias, edge_index, batch, csr, relu)

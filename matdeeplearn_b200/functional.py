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
        if red == 2 and src.requires_grad:
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
        gsrc = torch.empty(shape, dtype=g.dtype, device=g.device)
        rc = _lib.load().mdl_segment_reduce_bwd(_lib.ptr(g), _lib.ptr(ptr), _lib.ptr(perm),
                                                _lib.ptr(argmax), _lib.ptr(gsrc), S, rows, width,
                                                red, _lib.stream())
        _lib.check(rc, "mdl_segment_reduce_bwd")
        return gsrc, None, None, None, None


def segment_reduce(src, ptr, perm=None, reduce="mean"):
    """out[s] = reduce(src[perm[r]] for r in [ptr[s], ptr[s+1]));  ptr int32 [S+1]."""
    return SegmentReduceFn.apply(src, ptr, perm, reduce, ptr.shape[0] - 1)


# ----------------------------------------------------------------------------
# GaussianSmearing
# ----------------------------------------------------------------------------
def gaussian_smear(dist, offset, coeff):
    dist = dist.contiguous()
    out = torch.empty((dist.shape[0], offset.shape[0]), dtype=torch.float32, device=dist.device)
    rc = _lib.load().mdl_gaussian_smear(_lib.ptr(dist), _lib.ptr(offset.contiguous()), _lib.ptr(out),
                                        dist.shape[0], offset.shape[0], float(coeff), _lib.stream())
    _lib.check(rc, "mdl_gaussian_smear")
    return out


# ----------------------------------------------------------------------------
# CGConv
# ----------------------------------------------------------------------------
class CGConvFn(torch.autograd.Function):
    """x[N,C], lin_f/lin_s weight [C, 2C+G] + bias [C], edge_attr in slot order."""

    @staticmethod
    def forward(ctx, x, w_f, b_f, w_s, b_s, ea_slots, csr: GraphCSR, reduce):
        lib = _lib.load()
        x = x.contiguous()
        N, C = x.shape
        G = ea_slots.shape[1]
        assert w_f.shape == (C, 2 * C + G) and w_s.shape == (C, 2 * C + G)
        # column split of lin(cat[x_i, x_j, e]) = W_i x_i + W_j x_j + W_e e + b
        Wn = torch.cat([w_f[:, :C], w_s[:, :C], w_f[:, C:2 * C], w_s[:, C:2 * C]], 0)  # [4C, C]
        zeros = x.new_zeros(C)
        bias = torch.cat([b_f if b_f is not None else zeros, b_s if b_s is not None else zeros,
                          zeros, zeros])
        PQ = torch.addmm(bias, x, Wn.t())                                              # [N, 4C]
        WeT = torch.cat([w_f[:, 2 * C:], w_s[:, 2 * C:]], 0).t().contiguous()          # [G, 2C]
        out = torch.empty_like(x)
        rc = lib.mdl_cgconv_fwd(_lib.ptr(x), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(WeT),
                                _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src), _lib.ptr(csr.dst_dst),
                                _lib.ptr(csr.inv_deg_dst), _lib.ptr(out), N, csr.E, C, G,
                                _lib.REDUCE[reduce], _lib.stream())
        _lib.check(rc, "mdl_cgconv_fwd")
        ctx.save_for_backward(x, PQ, Wn, WeT, ea_slots)
        ctx.csr, ctx.reduce = csr, reduce
        ctx.has_bias = (b_f is not None, b_s is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, PQ, Wn, WeT, ea_slots = ctx.saved_tensors
        csr = ctx.csr
        N, C = x.shape
        G = ea_slots.shape[1]
        g = g.contiguous()
        dPQ = torch.empty_like(PQ)
        dWeT = torch.empty_like(WeT)
        ws_bytes = lib.mdl_cgconv_workspace_bytes(N, csr.E, C, G)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mdl_cgconv_bwd(_lib.ptr(g), _lib.ptr(PQ), _lib.ptr(ea_slots), _lib.ptr(WeT),
                                _lib.ptr(csr.dst_ptr), _lib.ptr(csr.dst_src), _lib.ptr(csr.dst_dst),
                                _lib.ptr(csr.src_ptr), _lib.ptr(csr.src_slot),
                                _lib.ptr(csr.inv_deg_dst), _lib.ptr(dPQ), _lib.ptr(dWeT), N, csr.E,
                                C, G, _lib.REDUCE[ctx.reduce], _lib.ptr(ws), ws_bytes, _lib.stream())
        _lib.check(rc, "mdl_cgconv_bwd")
        dx = torch.addmm(g, dPQ, Wn) if ctx.needs_input_grad[0] else None  # residual + projections
        dWn = dPQ.t().mm(x)                                                 # [4C, C]
        db = dPQ[:, :2 * C].sum(0)
        dWe = dWeT.t()                                                      # [2C, G]
        dw_f = torch.cat([dWn[0:C], dWn[2 * C:3 * C], dWe[0:C]], 1)
        dw_s = torch.cat([dWn[C:2 * C], dWn[3 * C:4 * C], dWe[C:2 * C]], 1)
        db_f = db[:C] if ctx.has_bias[0] else None
        db_s = db[C:] if ctx.has_bias[1] else None
        return dx, dw_f, db_f, dw_s, db_s, None, None, None


def cgconv(x, w_f, b_f, w_s, b_s, ea_slots, csr, reduce="mean"):
    return CGConvFn.apply(x, w_f, b_f, w_s, b_s, ea_slots, csr, reduce)

"""ctypes binding of libmdl_b200.so (the C ABI declared in include/mdl_b200.h).

There is no CPU fallback: importing an operator without the built library, or
calling one on a non-CUDA tensor, raises.  Build with
`python -c "import __graft_entry__ as g; g.build()"` or `make -C matdeeplearn_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdl_b200.so")

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float
_sz = C.c_size_t



class GraphStoreC(C.Structure):
    """mdl_graph_store (include/mdl_b200.h); field order and types must match the header
    (tests/test_cabi.py compiles the header with gcc and compares every offset)."""
    _fields_ = ([("num_graphs", _i64), ("num_nodes", _i64), ("num_edges", _i64),
                 ("F", _i32), ("G", _i32), ("U", _i32), ("Y", _i32)] +
                [(n, _p) for n in ("node_ptr", "edge_ptr", "x", "src", "dst", "d_hat", "edge_weight", "edge_attr",
                                   "u", "y", "dst_ptr", "dst_src", "dst_dst", "dst_eid", "src_ptr", "src_slot",
                                   "inv_deg_dst", "inv_deg_src")])


class BatchOutC(C.Structure):
    """mdl_batch_out (include/mdl_b200.h)."""
    _fields_ = ([("B", _i64), ("N", _i64), ("E", _i64)] +
                [(n, _p) for n in ("graph_ids", "node_off", "edge_off", "x", "edge_index", "d_hat", "edge_weight",
                                   "edge_attr", "edge_attr_slots", "batch", "u", "y", "dst_ptr", "dst_src",
                                   "dst_dst", "dst_eid", "src_ptr", "src_slot", "inv_deg_dst", "inv_deg_src",
                                   "graph_ptr", "smear_offset")] +
                [("smear_coeff", _f32)])


class WgradOutC(C.Structure):
    """mdl_wgrad_out (include/mdl_b200.h)."""
    _fields_ = [("block_rows", _i32), ("num_blocks", _i32), ("ldw", _i64), ("w", _p * 8), ("b", _p * 8)]


# name -> (restype, argtypes); must list every symbol of include/mdl_b200.h
SIGNATURES = {
    "mdl_version": (C.c_int, []),
    "mdl_last_error": (C.c_int, [C.c_char_p, _sz]),
    "mdl_launch_count": (_i64, []),
    "mdl_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "mdl_csr_from_coo": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mdl_gather_rows": (C.c_int, [_p, _p, _p, _i64, _i64, _p]),
    "mdl_scatter_rows": (C.c_int, [_p, _p, _p, _i64, _i64, _p]),
    "mdl_gaussian_smear": (C.c_int, [_p, _p, _p, _i64, _i32, _f32, _p]),
    "mdl_segment_reduce_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p]),
    "mdl_segment_reduce_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _p]),
    "mdl_cgconv_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "mdl_cgconv_pack_weights": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p]),
    "mdl_cgconv_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "mdl_cgconv_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p, _sz, _p]),
    "mdl_cgconv_tc_supported": (C.c_int, [_i32, _i32]),
    "mdl_cgconv_smear_supported": (C.c_int, [_i32, _i32]),
    "mdl_cgconv_smear_fwd": (C.c_int, [_p, _p, _p, _p, _f32, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p]),
    "mdl_cgconv_smear_bwd": (C.c_int, [_p, _p, _p, _p, _f32, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p,
                                       _sz, _p]),
    "mdl_spmm_edge": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "mdl_spmm_edge_scalar": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "mdl_edge_dot": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "mdl_edge_mul": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p]),
    "mdl_edge_gather_add": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _p]),
    "mdl_nnconv_msg_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "mdl_nnconv_msg_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _p]),
    "mdl_build_neighbors": (C.c_int, [_p, _p, _p, _i64, _i32, C.c_double, _i32, _p, _p, _p, _p]),
    "mdl_build_neighbors_lattice": (C.c_int, [_p, _p, _p, _p, _i64, _i32, C.c_double, _i32, _p, _p, _p, _p]),
    "mdl_build_emit": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "mdl_assemble_batch": (C.c_int, [C.POINTER(GraphStoreC), C.POINTER(BatchOutC), _p]),
    "mdl_batchnorm_workspace_bytes": (_sz, [_i64, _i32]),
    "mdl_batchnorm_fwd": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _f32, _f32, _p, _p, _p, _p, _sz, _p]),
    "mdl_batchnorm_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mdl_linear_wgrad_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "mdl_linear_wgrad": (C.c_int, [_p, _p, _i64, _i32, _i32, C.POINTER(WgradOutC), _p, _sz, _p]),
    "mdl_copy_mapped": (C.c_int, [_p, _i32, _i32, _i32, C.POINTER(WgradOutC), _p]),
    "mdl_adamw_step": (C.c_int, [_p, _p, _p, _p, _p, _p, _f32, _i64, _p]),
    "mdl_debug_set_phase_buffer": (C.c_int, [_p]),
    "mdl_linear_tc_supported": (C.c_int, [_i64, _i32, _i32]),
    "mdl_linear_tc": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "mdl_linear_wgrad_rs": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p, _sz, _p]),
    "mdl_edge_mlp2_supported": (C.c_int, [_i32, _i32, _i32]),
    "mdl_edge_mlp2_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p]),
    "mdl_edge_mlp2_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
}

# libmdl_b200_selftest.so (include/mdl_b200_selftest.h): tensor-core self-tests / probes, test infrastructure only
SELFTEST_LIB_PATH = os.path.join(_HERE, "libmdl_b200_selftest.so")
SELFTEST_SIGNATURES = {
    "mdl_selftest_umma": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p]),
    "mdl_selftest_umma_ts": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p]),
    "mdl_selftest_tmem_st_bench": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p]),
    "mdl_selftest_umma_probe": (C.c_int, [_p, _i32, _p, _i32, _p] + [_i32] * 12 + [_p]),
    "mdl_selftest_umma_ex": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "mdl_last_error": (C.c_int, [C.c_char_p, _sz]),
}

REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2}

_lib = None


def load():
    """Return the loaded library, binding prototypes on first use."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
            "(run __graft_entry__.build()).  matdeeplearn_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_selftest = None


def load_selftest():
    """The self-test library (tests / profiles scripts only)."""
    global _selftest
    if _selftest is None:
        if not os.path.exists(SELFTEST_LIB_PATH):
            raise ImportError(f"{SELFTEST_LIB_PATH} is missing (run __graft_entry__.build())")
        lib = C.CDLL(SELFTEST_LIB_PATH)
        for name, (res, args) in SELFTEST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _selftest = lib
    return _selftest


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().mdl_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses host tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("matdeeplearn_b200 operators need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("matdeeplearn_b200 operators need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().mdl_launch_count())

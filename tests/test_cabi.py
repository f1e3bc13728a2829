"""CPU: libmdl_b200.so loads and exports every symbol include/mdl_b200.h declares, the ctypes
table covers exactly those symbols, and the few host-only entry points behave (no kernel is
launched: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mdl_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MDL_API\s+[\w\s\*]+?\b(mdl_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from matdeeplearn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ("mdl_csr_from_coo", "mdl_cgconv_fwd", "mdl_cgconv_bwd", "mdl_segment_reduce_fwd",
                 "mdl_segment_reduce_bwd", "mdl_spmm_edge", "mdl_nnconv_msg_fwd", "mdl_edge_gather_add",
                 "mdl_gaussian_smear", "mdl_last_error", "mdl_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(cdll, name), f"{name} declared in mdl_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    assert sorted(lib.SIGNATURES) == declared_symbols()
    lib.load()  # binds every prototype; raises AttributeError on a missing symbol


def test_struct_layouts_match_the_header(lib, tmp_path):
    """The header is plain C (gcc compiles it) and the ctypes mirrors of its two descriptor
    structs agree with the compiler on size and on every field offset."""
    def fields(cls):
        return [n for n, _ in cls._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    structs = (("mdl_graph_store", lib.GraphStoreC), ("mdl_batch_out", lib.BatchOutC),
               ("mdl_wgrad_out", lib.WgradOutC))
    for cname, cls in structs:
        prog.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields(cls):
            prog.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    prog.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True,
                                                       check=True).stdout.splitlines())
    for cname, cls in structs:
        assert int(got[cname]) == ctypes.sizeof(cls)
        for f in fields(cls):
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_selftest_code_is_not_in_the_product_library(lib):
    """tensor-core self-tests / probes live in libmdl_b200_selftest.so (include/mdl_b200_selftest.h)"""
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "selftest" not in out
    text = open(os.path.join(ROOT, "include", "mdl_b200_selftest.h")).read()
    names = sorted(set(re.findall(r"MDL_API\s+[\w\s\*]+?\b(mdl_\w+)\s*\(", text)))
    assert names and sorted(n for n in lib.SELFTEST_SIGNATURES if n != "mdl_last_error") == names
    lib.load_selftest()


def test_no_unexpected_dynamic_dependencies(lib):
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libpython" not in out, "the C ABI must not depend on torch/python"


def test_host_only_entry_points(lib):
    L = lib.load()
    assert L.mdl_version() >= 100
    assert L.mdl_csr_workspace_bytes(1000, 13000) > 3 * 13000 * 4
    assert L.mdl_cgconv_workspace_bytes(1000, 13000, 64, 50) >= 148 * 50 * 128 * 4
    # argument validation happens before any CUDA call and reports through mdl_last_error
    rc = L.mdl_cgconv_fwd(None, None, None, None, None, None, None, None, None, 10, 10, 6, 50, 1, None)
    assert rc != 0 and "multiple of 4" in lib.last_error()
    rc = L.mdl_segment_reduce_fwd(None, None, None, None, None, 3, 0, 0, None)
    assert rc != 0 and "bad shape" in lib.last_error()
    rc = L.mdl_assemble_batch(None, None, None)
    assert rc != 0 and "null descriptor" in lib.last_error()
    store, out = lib.GraphStoreC(), lib.BatchOutC()
    out.B, out.N, out.E = 2, 10, 20
    store.F, store.G = 4, 8
    rc = L.mdl_assemble_batch(ctypes.byref(store), ctypes.byref(out), None)
    assert rc != 0 and "null arrays" in lib.last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "matdeeplearn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn

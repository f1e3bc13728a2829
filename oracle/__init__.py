"""CPU oracle for the message-passing hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, fp32/fp64) restatement of the arithmetic
that Fung-Lab/MatDeepLearn executes on its hot path.  That arithmetic lives in
third-party dependencies which are NOT vendored in /root/reference and are not
installed in this image:

  * torch_geometric  ("tested on" 2.0.1, reference README.md:30; unpinned in
    requirements.txt)  -> CGConv, InteractionBlock/CFConv, NNConv, MetaLayer,
    global_*_pool
  * torch_scatter    (reference README.md:32; unpinned) -> scatter, scatter_mean

PARITY UNPINNED for the PyG / torch_scatter arithmetic: the reference ships no
tests, golden vectors or saved outputs for this path, and
neither library can be imported here.  The restatement follows the published
PyG 2.0.1 formulas and the reference's own call sites.

What IS pinned (tests/golden/, produced by tests/golden/make_golden.py by
importing the reference's own Python sources from /root/reference with the
absent third-party modules stubbed):
  * process.py: threshold_sort, GaussianSmearing, OneHotDegree, NormalizeEdge
    (reference code executed as-is)
  * models/*.py glue (pre/post FC, BN placement, residuals, GRU threading,
    MEGNet block wiring) executed as-is, with the PyG ops bound to this
    oracle's restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (matdeeplearn_b200) never
does: it fails loudly if the CUDA library is missing.
"""
from . import pyg_ops, process, models  # noqa: F401

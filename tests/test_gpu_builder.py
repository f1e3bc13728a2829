"""GPU: the graph builder kernels (csrc/builder.cu, GraphStore.from_structures) reproduce the host
builder (matdeeplearn_b200.process.assemble_dataset, itself checked against the reference's
threshold_sort / dense_to_sparse / OneHotDegree / NormalizeEdge through the oracle and the golden
fixtures) bit for bit: same edges in the same order, same float32 weights, same node features."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LAYOUT = ("dst_ptr", "dst_src", "dst_dst", "dst_eid", "src_ptr", "src_slot", "inv_deg_dst", "inv_deg_src")


def _compare(structs, ys, **kw):
    from matdeeplearn_b200 import process as pr
    from matdeeplearn_b200.store import GraphStore
    ds = pr.assemble_dataset(structs, ys, **kw)
    host = GraphStore.from_dataset(ds, DEV)
    gpu = GraphStore.from_structures(structs, ys, DEV, **kw)
    assert np.array_equal(host.n_nodes, gpu.n_nodes) and np.array_equal(host.n_edges, gpu.n_edges)
    for name in ("node_ptr", "edge_ptr", "x", "src", "dst", "edge_weight", "d_hat", "u", "y"):
        a, b = getattr(host, name), getattr(gpu, name)
        assert a.shape == b.shape and a.dtype == b.dtype, name
        assert torch.equal(a, b), name
    for name in LAYOUT:
        assert torch.equal(getattr(host.layout, name), getattr(gpu.layout, name)), name
    assert abs(gpu.edge_range[1] - ds.edge_range[1]) == 0 and gpu.edge_range[0] == ds.edge_range[0]
    # and the batches they hand out are the same tensors
    idx = list(range(len(structs)))[::-1]
    a, b = host.batch(idx), gpu.batch(idx)
    for k in ("x", "edge_index", "edge_attr", "edge_weight", "batch", "u", "y"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    return gpu


@pytest.mark.parametrize("kind,n", [("bulk", 48), ("mof", 5)])
def test_builder_matches_host_builder_on_the_synthetic_workloads(kind, n):
    from matdeeplearn_b200 import process as pr
    structs, ys = pr.synthetic_structures(kind, n, seed=17)
    _compare(structs, ys)


def test_builder_edge_cases():
    rng = np.random.default_rng(0)
    structs, ys = [], []
    # simple cubic lattice, periodic: every row is full of exact distance ties (lower column wins)
    a, m = 2.5, 3
    grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) * a
    structs.append((rng.integers(1, 90, size=len(grid)), grid.astype(np.float64), np.array([a * m] * 3)))
    # non-periodic molecule-like cluster, fewer atoms than neighbours + 1
    structs.append((np.array([8, 1, 1]), np.array([[0.0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]]), None))
    # a single atom (only its loop), and two coincident atoms (zero distance is dropped)
    structs.append((np.array([26]), np.zeros((1, 3)), None))
    structs.append((np.array([6, 6, 7]), np.array([[1.0, 1, 1], [1.0, 1, 1], [2.0, 1, 1]]), None))
    # atoms further apart than the radius: no edges but the loops
    structs.append((np.array([2, 2]), np.array([[0.0, 0, 0], [50.0, 0, 0]]), None))
    # a dense box where the radius, not the neighbour count, is the active limit
    structs.append((rng.integers(1, 90, size=30), rng.uniform(0, 6.0, size=(30, 3)), np.array([6.0, 6.0, 6.0])))
    ys = list(rng.normal(size=len(structs)))
    gpu = _compare(structs, ys)
    assert gpu.n_edges[2] == 1 and gpu.n_edges[4] == 2
    _compare(structs, ys, radius=3.0, neighbors=4)
    _compare(structs, ys, radius=20.0, neighbors=30)


def test_builder_rejects_oversized_structures_and_cpu():
    from matdeeplearn_b200.store import GraphStore
    big = [(np.ones(5000, dtype=np.int64), np.random.default_rng(1).uniform(0, 50, size=(5000, 3)), None)]
    with pytest.raises(RuntimeError, match="tiled builder"):
        GraphStore.from_structures(big, [0.0], DEV)
    with pytest.raises(RuntimeError):
        GraphStore.from_structures(big, [0.0], "cpu")


def test_builder_triclinic_cells_match_host_builder():
    """general (triclinic) cells, what the reference gets from ASE's get_all_distances(mic=True) (process.py:284-287):
    the GPU builder evaluates the host builder's minimum-image search expression for expression -- same edges, same
    order, bit-identical float32 weights -- incl. a skewed cell that needs image shifts beyond +-1, a slab (free z
    axis), mixed with orthorhombic and non-periodic structures in one call."""
    rng = np.random.default_rng(21)
    cells = [np.array([[6.0, 0, 0], [2.5, 5.0, 0], [0.5, 0.8, 7.0]]),          # generic triclinic
             np.array([[4.0, 0, 0], [2.6, 3.2, 0], [0.3, 0.2, 4.5]]),          # skewed: image shifts up to +-5
             np.array([[5.0, 0, 0], [-2.5, 4.33, 0], [0, 0, 6.0]]),            # hexagonal
             np.array([[7.0, 0.5, 0.2], [0.1, 6.5, 0.4], [0.3, 0.2, 8.0]])]    # nearly orthogonal, fully dense matrix
    structs = []
    for c in cells:
        n = int(rng.integers(12, 40))
        structs.append((rng.integers(1, 90, size=n), rng.uniform(0, 1, (n, 3)) @ c, c))
    c = cells[0]
    structs.append((rng.integers(1, 90, size=20), rng.uniform(0, 1, (20, 3)) @ c, c, (True, True, False)))   # slab
    structs.append((rng.integers(1, 90, size=25), rng.uniform(0, 6.0, size=(25, 3)), np.array([6.0, 6.0, 6.0])))
    structs.append((np.array([8, 1, 1]), np.array([[0.0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]]), None))
    ys = list(rng.normal(size=len(structs)))
    _compare(structs, ys)
    _compare(structs, ys, radius=5.0, neighbors=6)

"""CPU: periodic cells in the host graph builder -- general (triclinic) minimum-image distances against a
brute-force search over lattice translates (what the reference obtains from ASE's
`get_all_distances(mic=True)`, process.py:284-287), and the orthorhombic fast path the GPU builder mirrors."""
import numpy as np
import pytest

from matdeeplearn_b200 import process as pr
from matdeeplearn_b200.store import _box_lengths


def _brute_force(pos, cell, span=3):
    d = pos[:, None, :] - pos[None, :, :]
    best = np.full(d.shape[:2], np.inf)
    r = range(-span, span + 1)
    for i in r:
        for j in r:
            for k in r:
                c = d + (i * cell[0] + j * cell[1] + k * cell[2])
                best = np.minimum(best, np.sqrt((c * c).sum(-1)))
    return best


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_triclinic_minimum_image_matches_brute_force(seed):
    rng = np.random.default_rng(seed)
    # skewed but physical cells: lengths 4..9 A, angles 60..120 degrees
    a, b, c = rng.uniform(4, 9, 3)
    al, be, ga = np.deg2rad(rng.uniform(60, 120, 3))
    cx = c * np.cos(be)
    cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    cell = np.array([[a, 0, 0], [b * np.cos(ga), b * np.sin(ga), 0],
                     [cx, cy, np.sqrt(max(c * c - cx * cx - cy * cy, 1e-6))]])
    pos = rng.uniform(-1, 2, (17, 3)) @ cell          # also atoms outside the home cell
    got = pr.pairwise_distances(pos, cell)
    ref = _brute_force(pos, cell)
    assert np.allclose(got, ref, rtol=0, atol=1e-12)
    assert np.allclose(got, got.T) and np.all(np.diag(got) == 0)


def test_hexagonal_cell_hand_worked():
    # 2D-hexagonal lattice (a = 3, 120 degrees) stacked with c = 10.  The hollow site (1/3, 2/3, 0) is a/sqrt(3)
    # from the origin; the site (2/3, 2/3, 0) = 2 from the origin in the home cell, but its image at
    # (-1/3, -1/3, 0) is |a1 + a2| / 3 = a/3 away; an atom shifted by whole lattice vectors changes nothing
    a = 3.0
    cell = np.array([[a, 0, 0], [-a / 2, a * np.sqrt(3) / 2, 0], [0, 0, 10.0]])
    pos = np.array([[0.0, 0, 0], [1 / 3, 2 / 3, 0], [2 / 3, 2 / 3, 0], [1 / 3 + 2, 2 / 3 - 1, 3.0]]) @ cell
    D = pr.pairwise_distances(pos, cell)
    assert D[0, 1] == pytest.approx(a / np.sqrt(3), rel=1e-12)
    assert D[0, 2] == pytest.approx(a / 3, rel=1e-12)
    assert D[1, 3] == pytest.approx(0.0, abs=1e-12) and D[0, 3] == pytest.approx(a / np.sqrt(3), rel=1e-12)


def test_orthorhombic_matrix_takes_the_bit_exact_vector_path():
    rng = np.random.default_rng(5)
    L = np.array([7.0, 8.5, 6.25])
    pos = rng.uniform(0, 1, (30, 3)) * L
    assert np.array_equal(pr.pairwise_distances(pos, L), pr.pairwise_distances(pos, np.diag(L)))
    assert np.allclose(pr.pairwise_distances(pos, L), _brute_force(pos, np.diag(L), span=2), atol=1e-12)


def test_graphs_from_a_triclinic_structure_are_well_formed():
    rng = np.random.default_rng(7)
    cell = np.array([[6.0, 0, 0], [2.0, 5.5, 0], [1.0, 1.5, 7.0]])
    n = 12
    numbers = rng.integers(1, 20, n)
    pos = rng.uniform(0, 1, (n, 3)) @ cell
    ds = pr.assemble_dataset([(numbers, pos, cell)], [0.5])
    g = ds[0]
    E = g.edge_index.shape[1]
    assert g.edge_attr.shape == (E, pr.DEFAULT_EDGE_LENGTH) and g.x.shape[0] == n
    loops = g.edge_index[0] == g.edge_index[1]
    assert int(loops.sum()) == n and float(g.edge_weight[loops].abs().max()) == 0.0
    assert float(g.edge_weight.max()) <= pr.DEFAULT_RADIUS + 1e-6
    assert 0.0 <= float(g.d_hat.min()) and float(g.d_hat.max()) <= 1.0


def test_gpu_builder_cell_argument():
    s_ortho = (np.array([1, 2]), np.zeros((2, 3)), np.array([5.0, 6.0, 7.0]))
    s_diag = (np.array([1, 2]), np.zeros((2, 3)), np.diag([5.0, 6.0, 7.0]))
    s_free = (np.array([1, 2]), np.zeros((2, 3)), None)
    got, lat = _box_lengths([s_ortho, s_diag, s_free])
    assert got.dtype == np.float64 and got.flags["C_CONTIGUOUS"] and lat is None
    assert np.array_equal(got, np.array([[5.0, 6.0, 7.0], [5.0, 6.0, 7.0], [0, 0, 0]]))
    # a general (triclinic) cell travels as a 28-double lattice record: vectors, inverse, shift ranges, pbc, flag
    tri = np.array([[5.0, 0, 0], [1.0, 6.0, 0], [0, 0, 7.0]])
    s_tri = (np.array([1, 2]), np.zeros((2, 3)), tri)
    got, lat = _box_lengths([s_ortho, s_tri])
    assert np.array_equal(got, np.array([[5.0, 6.0, 7.0], [0, 0, 0]])) and lat.shape == (2, 28)
    assert lat[0, 24] == 0.0 and lat[1, 24] == 1.0
    assert np.array_equal(lat[1, :9].reshape(3, 3), tri) and np.allclose(lat[1, 9:18].reshape(3, 3) @ tri, np.eye(3))
    assert np.array_equal(lat[1, 21:24], np.ones(3)) and (lat[1, 18:21] >= 1).all()


def test_slab_periodicity_is_per_axis():
    # periodic in x, y; free along z (pbc = T, T, F): two atoms 9 A apart in z in a 10 A box are 9 A apart, not 1 A
    L = np.array([5.0, 5.0, 10.0])
    pos = np.array([[0.5, 0.5, 0.5], [4.8, 0.5, 9.5]])
    full = pr.pairwise_distances(pos, L)
    slab = pr.pairwise_distances(pos, L, pbc=(True, True, False))
    assert full[0, 1] == pytest.approx(np.sqrt(0.7 ** 2 + 1.0 ** 2))
    assert slab[0, 1] == pytest.approx(np.sqrt(0.7 ** 2 + 9.0 ** 2))
    assert np.array_equal(pr.pairwise_distances(pos, np.array([5.0, 5.0, 0.0])), slab)   # length 0 = free axis
    assert np.array_equal(pr.pairwise_distances(pos, L, pbc=(False, False, False)), pr.pairwise_distances(pos))
    got, lat = _box_lengths([(np.array([1, 2]), pos, L, (True, True, False))])
    assert np.array_equal(got, np.array([[5.0, 5.0, 0.0]])) and lat is None


def test_triclinic_slab_matches_restricted_brute_force():
    rng = np.random.default_rng(11)
    cell = np.array([[6.0, 0, 0], [2.5, 5.0, 0], [0.5, 0.8, 14.0]])
    pos = rng.uniform(0, 1, (15, 3)) @ cell
    got = pr.pairwise_distances(pos, cell, pbc=(True, True, False))
    d = pos[:, None, :] - pos[None, :, :]
    best = np.full(d.shape[:2], np.inf)
    for i in range(-3, 4):
        for j in range(-3, 4):
            c = d + i * cell[0] + j * cell[1]
            best = np.minimum(best, np.sqrt((c * c).sum(-1)))
    assert np.allclose(got, best, atol=1e-12)


def test_ase_json_cells_and_flags():
    import json
    def rec(cell, pbc):
        return json.dumps({"1": {"numbers": {"__ndarray__": [[2], "int64", [78, 78]]},
                                 "positions": {"__ndarray__": [[2, 3], "float64", [0, 0, 0, 1.0, 1.0, 1.0]]},
                                 "cell": {"__ndarray__": [[3, 3], "float64", cell]}, "pbc": pbc}, "ids": [1], "nextid": 2})
    ortho = [5.0, 0, 0, 0, 6.0, 0, 0, 0, 7.0]
    tri = [5.0, 0, 0, 1.0, 6.0, 0, 0, 0, 7.0]
    assert pr.parse_ase_json(rec(ortho, [False, False, False]))[2] is None
    n, p, c = pr.parse_ase_json(rec(ortho, [True, True, True]))
    assert np.array_equal(c, [5.0, 6.0, 7.0])
    out = pr.parse_ase_json(rec(tri, [True, True, False]))
    assert len(out) == 4 and out[2].shape == (3, 3) and out[3] == (True, True, False)
    ds = pr.assemble_dataset([out], [1.0])
    assert ds[0].edge_index.shape[1] >= 2   # two loops at least; the graph builds

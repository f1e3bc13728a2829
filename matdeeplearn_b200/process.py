"""Host-side graph builder producing the operands of the message-passing engine
in the reference's layout, plus the synthetic workloads BASELINE.json names.

Mirrors the behaviour of reference matdeeplearn/process/process.py for the
pieces on the path:
  radius + k-nearest selection      process.py:287-292, 540-560
  row-major sparse edges + loops    process.py:294-305
  node features  one-hot Z ++ degree process.py:365-388, 594-605
  global min-max edge normalisation process.py:504, 626-653
  Gaussian edge expansion           process.py:500-509, 580-590
  u = zeros[1,3], y scalar          process.py:322-332, 695-703
The GPU rebuild of this stage is a "next" row; what
matters now is that its OUTPUT FORMAT is the engine's input contract.
"""
from __future__ import annotations

import io
import json
import os
import tarfile

import numpy as np
import torch

from .data import Data, GraphDataset

DEFAULT_RADIUS = 8.0
DEFAULT_NEIGHBORS = 12
DEFAULT_EDGE_LENGTH = 50
BENCH_SEED = 20260925


def knn_radius_edges(dist, radius=DEFAULT_RADIUS, neighbors=DEFAULT_NEIGHBORS):
    """Directed edges (row -> col) of one structure from its n x n distance matrix.

    Row i keeps its `neighbors + 1` closest columns (itself at distance 0 counts
    as one of them) that lie within `radius`; ties resolve to the lower column.
    Entries at distance exactly 0 (the diagonal, coincident atoms) are then
    dropped, edges are emitted row-major, and one loop (i, i) of weight 0 per
    node is appended last.  Returns (edge_index int64 [2,E], edge_weight f32 [E]).
    """
    dist = np.asarray(dist, dtype=np.float64)
    n = dist.shape[0]
    k = min(neighbors + 1, n)
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]           # [n,k] closest first
    dsel = np.take_along_axis(dist, order, axis=1)
    keep = (dsel <= radius) & (dsel != 0.0)
    rows = np.repeat(np.arange(n), k).reshape(n, k)[keep]
    cols = order[keep]
    # row-major emission: within a row, ascending column
    perm = np.lexsort((cols, rows))
    rows, cols = rows[perm], cols[perm]
    w = dist[rows, cols].astype(np.float32)
    loops = np.arange(n)
    ei = np.stack([np.concatenate([rows, loops]), np.concatenate([cols, loops])]).astype(np.int64)
    ew = np.concatenate([w, np.zeros(n, np.float32)])
    return torch.from_numpy(ei), torch.from_numpy(ew)


def node_features(numbers, edge_index, neighbors=DEFAULT_NEIGHBORS, width=100):
    """[n, width + neighbors + 2]: one-hot atomic number (Z at column Z-1, the
    layout of the reference's dictionary_default.json) ++ one-hot of the number
    of edges leaving each node, loops included."""
    z = torch.as_tensor(np.asarray(numbers), dtype=torch.long)
    n = z.numel()
    x = torch.zeros(n, width + neighbors + 2, dtype=torch.float32)
    x[torch.arange(n), z - 1] = 1.0
    deg = torch.bincount(edge_index[0], minlength=n)
    x[torch.arange(n), width + deg] = 1.0
    return x


def gaussian_expand(d_hat, resolution=DEFAULT_EDGE_LENGTH, start=0.0, stop=1.0, width=0.2):
    mu = torch.linspace(start, stop, resolution, dtype=d_hat.dtype)
    coeff = -0.5 / ((stop - start) * width) ** 2
    diff = d_hat[:, None] - mu[None, :]
    return torch.exp(coeff * (diff * diff))


def orthorhombic_lengths(cell, pbc=None):
    """[3] box lengths if `cell` is a length vector or a diagonal 3x3 matrix, else None.  A length of 0
    means "not periodic along this axis" (what the GPU builder tests per axis); `pbc` (3 booleans, ASE's
    per-axis flags) zeroes the lengths of the free axes."""
    c = np.asarray(cell, dtype=np.float64)
    if c.shape == (3, 3):
        if not np.array_equal(c, np.diag(np.diag(c))):
            return None
        c = np.diag(c)
    elif c.shape != (3,):
        return None
    c = c.copy()
    if pbc is not None:
        c[~np.asarray(pbc, dtype=bool)] = 0.0
    return c


def lattice_record(cell, pbc=None):
    """What the minimum-image search of a general cell needs, computed once per structure on the host (and handed to
    the GPU builder as 28 doubles): the lattice vectors (rows), their inverse, the image-shift range per axis, the
    periodicity flags.  The wrapped vector is no longer than R0 = (|a1|+|a2|+|a3|)/2, so a better image needs a
    lattice vector shorter than 2 R0, i.e. |shift_i| <= ceil(2 R0 / h_i) with h_i the cell's height along axis i."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    per = np.ones(3, dtype=bool) if pbc is None else np.asarray(pbc, dtype=bool)
    inv = np.linalg.inv(cell)
    vol = abs(np.linalg.det(cell))
    r0 = 0.5 * np.linalg.norm(cell, axis=1).sum()
    heights = np.array([vol / np.linalg.norm(np.cross(cell[(i + 1) % 3], cell[(i + 2) % 3])) for i in range(3)])
    n = np.ceil(2.0 * r0 / heights).astype(int)
    n[~per] = 0
    if n.max() > 8:   # (2n+1)^3 image shifts: a needle-shaped / badly reduced cell; refuse rather than truncate the search
        raise ValueError(f"minimum image: cell too skewed (needs image shifts up to {n.tolist()}); reduce the cell first")
    return cell, inv, n, per


def _general_minimum_image(d, cell, pbc=None):
    """Minimum-image vectors for a general (triclinic) cell, rows of `cell` = lattice vectors.
    What the reference gets from ASE (`get_all_distances(mic=True)`, process.py:284-287): the shortest
    of all lattice translates of each difference vector.  Wrap the fractional coordinates into
    [-0.5, 0.5], then search every integer shift that can still shorten a wrapped vector (lattice_record).
    Every product and sum is written out elementwise in a fixed order (no BLAS, no FMA): the GPU builder
    (csrc/builder.cu) evaluates the same expressions and reproduces the result bit for bit."""
    cell, inv, n, per = lattice_record(cell, pbc)
    frac = [(d[..., 0] * inv[0, k] + d[..., 1] * inv[1, k]) + d[..., 2] * inv[2, k] for k in range(3)]
    for k in range(3):
        if per[k]:   # free axes (slabs, wires) keep their coordinate
            frac[k] = frac[k] - np.round(frac[k])
    w = np.stack([(frac[0] * cell[0, c] + frac[1] * cell[1, c]) + frac[2] * cell[2, c] for c in range(3)], -1)
    best = w.copy()
    best_d2 = (w[..., 0] * w[..., 0] + w[..., 1] * w[..., 1]) + w[..., 2] * w[..., 2]
    for i in range(-n[0], n[0] + 1):
        for j in range(-n[1], n[1] + 1):
            for k in range(-n[2], n[2] + 1):
                if i == j == k == 0:
                    continue
                shift = (float(i) * cell[0] + float(j) * cell[1]) + float(k) * cell[2]
                c = w + shift
                d2 = (c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2]
                m = d2 < best_d2
                best[m] = c[m]
                best_d2[m] = d2[m]
    return best


def pairwise_distances(pos, cell=None, pbc=None):
    """Euclidean, or minimum-image for a periodic cell: `cell` = [3] box lengths / diagonal 3x3
    (orthorhombic; the expression the GPU builder reproduces bit for bit; length 0 = free axis) or a
    general 3x3 matrix whose rows are the lattice vectors (host builder only).  `pbc`: ASE's per-axis
    periodicity flags (default: periodic along every axis of a given cell)."""
    pos = np.asarray(pos, dtype=np.float64)
    d = pos[:, None, :] - pos[None, :, :]
    if cell is not None and (pbc is None or np.any(pbc)):
        L = orthorhombic_lengths(cell, pbc)
        if L is not None:
            for ax in range(3):  # per axis, as the GPU builder: wrap only where the box is periodic
                if L[ax] > 0.0:
                    d[..., ax] -= np.round(d[..., ax] / L[ax]) * L[ax]
        else:
            d = _general_minimum_image(d, np.asarray(cell, dtype=np.float64).reshape(3, 3), pbc)
    # (dx^2 + dy^2) + dz^2 in that order, plain multiplies and adds: einsum would pick an FMA kernel whose
    # rounding depends on the host CPU; the GPU builder (csrc/builder.cu) reproduces THIS expression bit for bit
    return np.sqrt((d * d).sum(-1))


def assemble_dataset(structures, targets, radius=DEFAULT_RADIUS, neighbors=DEFAULT_NEIGHBORS,
                     edge_length=DEFAULT_EDGE_LENGTH):
    """structures: iterable of (numbers, positions, cell-or-None[, pbc]); cell = [3] box lengths or a
    3x3 matrix of lattice vectors (rows), pbc = optional per-axis flags, see pairwise_distances.

    Returns GraphDataset whose graphs carry x, edge_index, edge_weight (raw
    Angstrom), edge_attr (Gaussian basis of the globally min-max normalised
    distance), d_hat (that normalised distance; engine extra), u, y.
    """
    graphs = []
    for st, y in zip(structures, targets):
        numbers, pos, cell = st[0], st[1], st[2]
        D = pairwise_distances(pos, cell, st[3] if len(st) > 3 else None)
        ei, ew = knn_radius_edges(D, radius, neighbors)
        x = node_features(numbers, ei, neighbors)
        graphs.append(Data(x=x, edge_index=ei, edge_weight=ew,
                           u=torch.zeros(1, 3), y=torch.tensor(float(y), dtype=torch.float32)))
    lo = min(float(g.edge_weight.min()) for g in graphs)
    hi = max(float(g.edge_weight.max()) for g in graphs)
    # tensor / tensor: an IEEE division on every backend (tensor / python-scalar becomes a multiplication
    # by the reciprocal in some of torch's kernels, e.g. on CUDA, which is 1 ulp off for some elements)
    span = torch.tensor(hi - lo, dtype=torch.float32)
    for g in graphs:
        g.d_hat = (g.edge_weight - lo) / span
        g.edge_attr = gaussian_expand(g.d_hat, edge_length)
    ds = GraphDataset(graphs)
    ds.edge_range = (lo, hi)
    ds.smear = dict(start=0.0, stop=1.0, resolution=edge_length, width=0.2)  # reference process.py:500-502
    return ds


# ----------------------------------------------------------------------------
# synthetic workloads
# ----------------------------------------------------------------------------
def _random_structure(rng, n, density, min_sep=1.6, n_species=20):
    L = (n / density) ** (1.0 / 3.0)
    pos = np.empty((n, 3))
    count = 0
    while count < n:
        cand = rng.uniform(0.0, L, size=(4 * (n - count) + 8, 3))
        for c in cand:
            if count == n:
                break
            if count:
                d = pos[:count] - c
                d -= np.round(d / L) * L
                if (np.einsum("ij,ij->i", d, d) < min_sep * min_sep).any():
                    continue
            pos[count] = c
            count += 1
    numbers = rng.integers(1, n_species + 1, size=n)
    return numbers, pos, np.array([L, L, L])


def synthetic_structures(kind="bulk", num_graphs=256, seed=BENCH_SEED):
    """(structures, targets) of the synthetic workloads: structures = [(numbers, positions, box lengths)].
       kind="bulk": n ~ clip(round(N(30,8)),4,60), density 0.06 A^-3
       kind="mof" : n ~ clip(round(N(200,40)),80,400), density 0.025 A^-3"""
    rng = np.random.default_rng(seed)
    if kind == "bulk":
        mean, std, lo, hi, rho = 30, 8, 4, 60, 0.06
    elif kind == "mof":
        mean, std, lo, hi, rho = 200, 40, 80, 400, 0.025
    else:
        raise ValueError(kind)
    structs, ys = [], []
    for _ in range(num_graphs):
        n = int(np.clip(np.rint(rng.normal(mean, std)), lo, hi))
        structs.append(_random_structure(rng, n, rho))
        ys.append(rng.normal())
    return structs, ys


def synthetic_dataset(kind="bulk", num_graphs=256, seed=BENCH_SEED, edge_length=DEFAULT_EDGE_LENGTH):
    """The synthetic workload run through the host builder (see synthetic_structures)."""
    structs, ys = synthetic_structures(kind, num_graphs, seed)
    return assemble_dataset(structs, ys, edge_length=edge_length)


# ----------------------------------------------------------------------------
# the reference's bundled fixture: data/test_data (ASE db-json, non-periodic)
# ----------------------------------------------------------------------------
def _decode_ndarray(obj):
    shape, dtype, flat = obj["__ndarray__"]
    return np.asarray(flat, dtype=dtype).reshape(shape)


def parse_ase_json(text):
    rec = json.loads(text)
    rec = rec[next(k for k in rec if k.isdigit())]
    numbers = _decode_ndarray(rec["numbers"])
    pos = _decode_ndarray(rec["positions"])
    pbc = rec.get("pbc", [False, False, False])
    if isinstance(pbc, dict):
        pbc = _decode_ndarray(pbc).tolist()
    cell = None
    if any(pbc):
        c = rec["cell"]
        c = _decode_ndarray(c) if isinstance(c, dict) else np.asarray(c, dtype=float)
        cell = np.diag(c).copy() if np.allclose(c, np.diag(np.diag(c))) else c  # lengths, or the full matrix
        if not all(pbc):
            return numbers, pos, cell, tuple(bool(b) for b in pbc)   # slab / wire: free axes
    return numbers, pos, cell


def load_ase_json_tar(tar_path, limit=None):
    """Read <id>.json structures + targets.csv straight out of the reference's
    test_data tarball (no ASE needed: ASE db-json is plain JSON)."""
    structs, ys, ids = [], [], []
    with tarfile.open(tar_path) as tf:
        members = {os.path.basename(m.name): m for m in tf.getmembers() if m.isfile()}
        rows = io.TextIOWrapper(tf.extractfile(members["targets.csv"])).read().strip().splitlines()
        for row in rows[:limit]:
            sid, y = row.split(",")[:2]
            structs.append(parse_ase_json(tf.extractfile(members[sid + ".json"]).read().decode()))
            ys.append(float(y))
            ids.append(sid)
    ds = assemble_dataset(structs, ys)
    ds.ids = ids
    return ds

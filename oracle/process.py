"""Oracle restatement of the reference graph builder / edge featuriser
(TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows reference matdeeplearn/process/process.py:
  threshold_sort      :540-560   (radius mask + ordinal rank, keep rank <= k+1)
  dense_to_sparse +
  add_self_loops      :294-303   (row-major nonzero scan, N loops appended)
  OneHotDegree        :594-605   (degree of edge_index[0], incl. the loop)
  GaussianSmearing    :580-590
  GetRanges/Normalize :626-653   (dataset-global min/max)
These functions ARE pinned: tests/golden/process_*.npz hold outputs of the
reference's own code for them (tests/golden/make_golden.py).
"""
import numpy as np
import torch


def threshold_sort(dist, radius, neighbors):
    """Return the trimmed distance matrix (reference threshold_sort, adj=False).

    Entry (i, j) keeps dist[i, j] iff dist[i, j] <= radius and its 1-based
    ordinal rank inside row i (ties broken by column order, i.e. a stable
    sort) is <= neighbors + 1; everything else becomes 0.
    """
    dist = np.asarray(dist, dtype=np.float64)
    order = np.argsort(dist, axis=1, kind="stable")
    rank = np.empty_like(order)
    n = dist.shape[0]
    rows = np.arange(n)[:, None]
    rank[rows, order] = np.arange(1, dist.shape[1] + 1)[None, :]
    keep = (dist <= radius) & (rank <= neighbors + 1)
    return np.where(keep, dist, 0.0)


def dense_to_sparse_with_loops(trimmed):
    """edge_index [2, E] int64 + edge_weight [E] float32.

    Non-zero entries in row-major order (so exact-zero distances, including the
    diagonal, drop out), then one (i, i) loop per node with weight 0, appended
    last -- reference process.py:294-303.
    """
    t = torch.as_tensor(np.asarray(trimmed), dtype=torch.float32)
    n = t.shape[0]
    idx = t.nonzero(as_tuple=False).t().contiguous()
    w = t[idx[0], idx[1]]
    loops = torch.arange(n, dtype=torch.long)
    edge_index = torch.cat([idx, torch.stack([loops, loops])], dim=1)
    edge_weight = torch.cat([w, torch.zeros(n, dtype=torch.float32)])
    return edge_index, edge_weight


def one_hot_degree(edge_index, num_nodes, max_degree):
    """[N, max_degree + 1] one-hot of the out-degree (occurrences in
    edge_index[0], loops included) -- reference process.py:594-597."""
    deg = torch.bincount(edge_index[0], minlength=num_nodes)
    return torch.nn.functional.one_hot(deg, num_classes=max_degree + 1).to(torch.float32)


def gaussian_smearing(dist, start=0.0, stop=1.0, resolution=50, width=0.2):
    """exp(coeff * (d - mu_k)^2), mu = linspace(start, stop, resolution),
    coeff = -0.5 / ((stop - start) * width)^2 -- reference process.py:580-590."""
    offset = torch.linspace(start, stop, resolution, dtype=dist.dtype)
    coeff = -0.5 / ((stop - start) * width) ** 2
    d = dist.unsqueeze(-1) - offset.view(1, -1)
    return torch.exp(coeff * (d * d))


def normalize_edges(weights):
    """Dataset-global min-max normalisation of a list of per-graph distance
    tensors -- reference process.py:626-653."""
    lo = min(float(w.min()) for w in weights if w.numel() > 0)
    hi = max(float(w.max()) for w in weights if w.numel() > 0)
    return [(w - lo) / (hi - lo) for w in weights], lo, hi


def mic_distance_matrix(pos, cell=None, pbc=False):
    """All-pairs distances; minimum-image for an orthorhombic periodic cell.

    The reference calls ASE get_all_distances(mic=True) (process.py:284); ASE
    is absent here.  For non-periodic input this is the Euclidean matrix; for
    the synthetic orthorhombic cells of the benchmark the per-axis wrap below
    is the exact minimum image.
    """
    pos = np.asarray(pos, dtype=np.float64)
    d = pos[:, None, :] - pos[None, :, :]
    if pbc:
        L = np.asarray(cell, dtype=np.float64).reshape(-1)[:3] if np.ndim(cell) == 1 else np.diag(np.asarray(cell, dtype=np.float64))
        d = d - np.round(d / L) * L
    return np.sqrt((d * d).sum(-1))


def atom_one_hot(numbers, width=100):
    """dictionary_default.json maps Z -> one-hot at index Z-1 (100 wide)."""
    z = torch.as_tensor(np.asarray(numbers), dtype=torch.long)
    return torch.nn.functional.one_hot(z - 1, num_classes=width).to(torch.float32)


def build_graph(numbers, pos, cell=None, pbc=False, radius=8.0, neighbors=12):
    """One structure -> dict(x_onehot_z, edge_index, edge_weight, num_nodes)
    following reference process.py:284-305 (+ :365-388 for x)."""
    D = mic_distance_matrix(pos, cell, pbc)
    trimmed = threshold_sort(D, radius, neighbors)
    edge_index, edge_weight = dense_to_sparse_with_loops(trimmed)
    n = len(numbers)
    x = torch.cat([atom_one_hot(numbers), one_hot_degree(edge_index, n, neighbors + 1)], dim=1)
    return dict(x=x, edge_index=edge_index, edge_weight=edge_weight, num_nodes=n)

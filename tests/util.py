"""Shared helpers for the parity tests."""
import numpy as np
import torch


def assert_close(got, ref, rtol, atol_rel, what=""):
    """|got - ref| <= atol_rel * max|ref| + rtol * |ref|  elementwise."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = float(ref.abs().max()) if ref.numel() else 0.0
    err = (got - ref).abs()
    bound = atol_rel * scale + rtol * ref.abs()
    bad = err > bound
    if bad.any():
        i = int(torch.argmax(err - bound))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; worst err "
            f"{float(err.flatten()[i]):.3e} at ref {float(ref.flatten()[i]):.3e} "
            f"(scale {scale:.3e}, rtol {rtol}, atol_rel {atol_rel})")


def random_graph(n_nodes, n_edges, seed, hub=None, isolated=0, loops=True):
    """Random directed multigraph-free COO with optional hub destination and
    isolated (no in-edges, no loop) trailing nodes.  Returns int64 [2,E]."""
    rng = np.random.default_rng(seed)
    live = n_nodes - isolated
    pairs = set()
    while len(pairs) < n_edges:
        a, b = int(rng.integers(live)), int(rng.integers(live))
        if a != b:
            pairs.add((a, b))
    if hub is not None:
        hub_node, deg = hub
        for a in rng.permutation(live)[:deg]:
            if int(a) != hub_node:
                pairs.add((int(a), hub_node))
    edges = sorted(pairs)  # row-major like the reference builder
    if loops:
        edges += [(i, i) for i in range(live)]
    return torch.tensor(edges, dtype=torch.int64).t().contiguous()


def contiguous_batch_vector(n_nodes, n_graphs, seed):
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.choice(np.arange(1, n_nodes), size=n_graphs - 1, replace=False))
    sizes = np.diff(np.concatenate([[0], cuts, [n_nodes]]))
    return torch.repeat_interleave(torch.arange(n_graphs), torch.tensor(sizes))


def block_diagonal_graph(sizes, edges_per_node, seed):
    """Crystal-batch-shaped COO: independent blocks of the given node counts, every edge inside its
    block, one self-loop per node appended per block (as the reference builder does).  Returns
    int64 [2,E] in the reference's per-graph row-major order."""
    rng = np.random.default_rng(seed)
    out, base = [], 0
    for n in sizes:
        pairs = set()
        want = min(n * edges_per_node, n * (n - 1))
        while len(pairs) < want:
            a, b = int(rng.integers(n)), int(rng.integers(n))
            if a != b:
                pairs.add((a, b))
        out += [(base + a, base + b) for a, b in sorted(pairs)] + [(base + i, base + i) for i in range(n)]
        base += n
    return torch.tensor(out, dtype=torch.int64).t().contiguous()

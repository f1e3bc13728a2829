"""CPU: data.GaussianEdgeAttr -- the 4-bytes-per-edge form of edge_attr -- expands to exactly the tensor the reference's
GaussianSmearing produces (process/process.py:580-590 applied at :500-502), travels with a Batch, and keeps its memoised
expansions consistent with the version of d_hat."""
import torch

from matdeeplearn_b200 import process as pr
from matdeeplearn_b200.data import GaussianEdgeAttr, dense_edge_attr


def test_materialize_equals_the_dataset_builders_edge_attr():
    ds = pr.synthetic_dataset("bulk", 6, seed=3)
    b = ds.batch()
    lazy = b.with_lazy_edge_attr()
    assert isinstance(lazy.edge_attr, GaussianEdgeAttr) and lazy.edge_attr.shape == b.edge_attr.shape
    assert lazy.edge_attr.dim() == 2 and lazy.edge_attr.size(1) == b.edge_attr.shape[1] and not lazy.edge_attr.requires_grad
    assert torch.equal(lazy.edge_attr.materialize(), b.edge_attr)          # same formula, same linspace offsets
    assert dense_edge_attr(lazy.edge_attr) is lazy.edge_attr.materialize()  # memoised
    assert dense_edge_attr(b.edge_attr) is b.edge_attr
    # every other tensor of the batch is shared, not copied
    assert lazy.x is b.x and lazy.edge_index is b.edge_index


def test_parameters_follow_the_reference_module():
    g = GaussianEdgeAttr(torch.rand(7), start=0.0, stop=1.0, resolution=50, width=0.2)
    assert abs(g.coeff - (-0.5 / (1.0 * 0.2) ** 2)) < 1e-12                 # process.py:585
    assert torch.equal(g.offset, torch.linspace(0.0, 1.0, 50))              # process.py:584
    assert g.fusable()
    assert not GaussianEdgeAttr(torch.rand(3), resolution=50, width=0.01).fusable()   # exp(-5000 d^2): would underflow
    assert not GaussianEdgeAttr(torch.rand(3), resolution=1).fusable()


def test_memo_tracks_the_version_of_d_hat_and_moves_with_the_batch():
    d = torch.rand(9)
    g = GaussianEdgeAttr(d, resolution=8)
    a = g.materialize()
    d.mul_(0.5)                      # in-place edit bumps the version: the memo must not be served
    b = g.materialize()
    assert not torch.equal(a, b) and torch.allclose(b, torch.exp(g.coeff * (d[:, None] - g.offset[None, :]) ** 2))
    g.forget()
    assert g._dense is None and g._slots is None
    h = g.double()
    assert h.dtype == torch.float64 and h.params() == g.params()
    ds = pr.synthetic_dataset("bulk", 3, seed=1)
    lb = ds.batch().with_lazy_edge_attr()
    moved = lb.to("cpu")
    assert isinstance(moved.edge_attr, GaussianEdgeAttr) and moved.edge_attr.d_hat is not None
    assert isinstance(lb.double().edge_attr, GaussianEdgeAttr) and lb.double().edge_attr.dtype == torch.float64


def test_chunked_recurrence_of_the_fused_kernels_stays_within_the_stated_bound():
    """csrc/edge_dev.cuh: smear_chunk8 -- eight consecutive basis values from two exponentials
    (t_{k+1} = t_k rho_k, rho_{k+1} = rho_k q, restarted every 8 columns from offset[k]) -- restated in float32 numpy
    and compared with the reference's formula exp(coeff (d - mu_k)^2) (process/process.py:588-590) in float64:
    within 2.5e-6 of the basis' scale (1) for the reference's parameters (G = 50; G >= 37 in general), within 6e-6
    for coarse (G = 8: the per-step ratio is large) or twice narrower bases (the GPU's ex2.approx adds about one more
    ulp per exponential)."""
    import numpy as np
    f = np.float32
    log2e = f(1.4426950408889634)
    for G, width in ((50, 0.2), (64, 0.2), (37, 0.2), (50, 0.1), (8, 0.2), (2, 0.2)):
        mu = torch.linspace(0.0, 1.0, G).numpy().astype(f)
        coeff = f(-0.5 / (1.0 * width) ** 2)
        c2 = f(coeff * log2e)
        dmu = f((mu[G - 1] - mu[0]) / f(G - 1)) if G > 1 else f(0)
        q2 = np.exp2(f(f(2.0) * c2 * dmu * dmu)).astype(f)
        d = np.concatenate([np.linspace(0, 1, 4001), [0.0, 1.0, 0.5]]).astype(f)
        got = np.zeros((d.size, G), dtype=f)
        KP = (G + 7) // 8 * 8
        for k0 in range(0, KP, 8):
            diff = (d - mu[min(k0, G - 1)]).astype(f)
            t = np.exp2((c2 * (diff * diff).astype(f)).astype(f)).astype(f)
            rho = np.exp2((c2 * dmu * (dmu - f(2.0) * diff).astype(f)).astype(f)).astype(f)
            for j in range(8):
                if k0 + j < G:
                    got[:, k0 + j] = t
                t = (t * rho).astype(f)
                rho = (rho * q2).astype(f)
        ref = np.exp(float(coeff) * (d.astype(np.float64)[:, None] - mu.astype(np.float64)[None, :]) ** 2)
        err = np.abs(got.astype(np.float64) - ref).max()
        assert err <= (2.5e-6 if (G >= 37 and width == 0.2) else 6e-6), (G, width, err)

"""CPU: pin the oracle (and the product's host-side builder) against fixtures
produced by running the reference's own sources (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from tests.golden.make_golden import MODEL_CFGS

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pg():
    return np.load(os.path.join(G, "process_golden.npz"))


@pytest.mark.parametrize("i", [0, 1, 2, 3])
def test_threshold_sort_oracle_and_builder(pg, i):
    from oracle import process as OP
    from matdeeplearn_b200 import process as PR
    D = pg[f"dist_{i}"]
    radius, k = pg[f"args_{i}"]
    ref = pg[f"trimmed_{i}"]
    got = OP.threshold_sort(D, float(radius), int(k))
    assert np.array_equal(got, ref)
    # product builder: same edge set/order as nonzero-scan of the reference matrix + loops
    ei, ew = PR.knn_radius_edges(D, float(radius), int(k))
    r, c = np.nonzero(ref)
    n = D.shape[0]
    assert np.array_equal(ei[0].numpy(), np.concatenate([r, np.arange(n)]))
    assert np.array_equal(ei[1].numpy(), np.concatenate([c, np.arange(n)]))
    assert np.array_equal(ew.numpy(), np.concatenate([ref[r, c], np.zeros(n)]).astype(np.float32))


@pytest.mark.parametrize("i", [0, 1, 2])
def test_sparse_edges_and_degree_onehot(pg, i):
    from oracle import process as OP
    ref_t = pg[f"trimmed_{i}"] if pg[f"args_{i}"][0] == 8.0 and pg[f"args_{i}"][1] == 12 else None
    D = pg[f"dist_{i}"]
    t = OP.threshold_sort(D, 8.0, 12)
    ei, ew = OP.dense_to_sparse_with_loops(t)
    assert np.array_equal(ei.numpy(), pg[f"ei_{i}"])
    assert np.array_equal(ew.numpy(), pg[f"ew_{i}"])
    deg = OP.one_hot_degree(ei, D.shape[0], 13)
    assert np.array_equal(deg.numpy(), pg[f"onehotdeg_{i}"][:, 3:])  # fixture x had 3 leading zero columns
    if ref_t is not None:
        assert np.array_equal(t, ref_t)


def test_normalize_edges(pg):
    from oracle import process as OP
    ws = [torch.from_numpy(pg[f"ew_{i}"]) for i in range(3)]
    normed, lo, hi = OP.normalize_edges(ws)
    for i in range(3):
        assert np.array_equal(normed[i].numpy(), pg[f"norm_{i}"])


@pytest.mark.parametrize("Gw", [50, 100, 200])
def test_gaussian_smearing(pg, Gw):
    from oracle import process as OP
    from matdeeplearn_b200 import process as PR
    d = torch.from_numpy(pg["smear_in"])
    assert np.array_equal(OP.gaussian_smearing(d, 0.0, 1.0, Gw, 0.2).numpy(), pg[f"smear_{Gw}"])
    assert np.array_equal(PR.gaussian_expand(d, Gw).numpy(), pg[f"smear_{Gw}"])
    assert abs(float(pg[f"smear_coeff_{Gw}"]) + 12.5) < 1e-12


def load_batch():
    from matdeeplearn_b200.data import Batch
    z = np.load(os.path.join(G, "batch_inputs.npz"))
    b = Batch(**{k: torch.from_numpy(z[k]) for k in z.files})
    b.num_graphs = int(b.y.shape[0])
    return b


class _DS:
    """what a model constructor reads from the dataset"""

    def __init__(self, b):
        self.num_features = b.x.shape[1]
        self.num_edge_features = b.edge_attr.shape[1]
        from matdeeplearn_b200.data import Data
        self._g = Data(y=b.y[0], u=b.u[:1])

    def __getitem__(self, i):
        return self._g


def load_model_fixture(tag):
    z = np.load(os.path.join(G, f"model_{tag}.npz"))
    sd = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return z, sd, grads


@pytest.mark.parametrize("tag", list(MODEL_CFGS))
def test_oracle_model_glue_matches_reference_glue(tag):
    """oracle/models.py (own restatement of the glue) == the reference's model
    file executed as-is, both on the oracle's operator restatement, fp64."""
    from oracle import models as OM
    b = load_batch()
    z, sd, grads = load_model_fixture(tag)
    model = getattr(OM, tag.split("_")[0])(_DS(b), **MODEL_CFGS[tag]).double()
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    out = model(b.double())
    np.testing.assert_allclose(out.detach().numpy(), z["out_train"], rtol=1e-10, atol=1e-12)
    loss = torch.nn.functional.l1_loss(out, b.y.double())
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-12
    for name, p in model.named_parameters():
        ref = grads[name]
        if ref.numel() == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        np.testing.assert_allclose(p.grad.numpy(), ref.numpy(), rtol=1e-8, atol=1e-12, err_msg=name)
    model.eval()
    with torch.no_grad():
        np.testing.assert_allclose(model(b.double()).numpy(), z["out_eval"], rtol=1e-10, atol=1e-12)

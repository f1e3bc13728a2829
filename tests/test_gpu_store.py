"""GPU: a batch assembled from the device-resident GraphStore (mdl_assemble_batch) is, bit for bit,
the batch the reference's collate would produce (Batch.from_data_list, restating PyG's collate used at
training.py:300-307) copied to the device, and the layout it carries equals mdl_csr_from_coo's."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LAYOUT = ("dst_ptr", "dst_src", "dst_dst", "dst_eid", "src_ptr", "src_slot", "inv_deg_dst", "inv_deg_src",
          "graph_ptr")


def _dataset(kind="bulk", n=40, seed=5):
    from matdeeplearn_b200 import process as pr
    return pr.synthetic_dataset(kind, n, seed=seed)


@pytest.mark.parametrize("keep", [True, False])
@pytest.mark.parametrize("kind,n,B", [("bulk", 40, 17), ("bulk", 40, 1), ("mof", 6, 4)])
def test_assembled_batch_equals_host_collate(kind, n, B, keep):
    from matdeeplearn_b200.csr import GraphCSR, gather_rows
    from matdeeplearn_b200.data import Batch
    from matdeeplearn_b200.store import GraphStore
    from matdeeplearn_b200 import functional as MF
    ds = _dataset(kind, n)
    store = GraphStore.from_dataset(ds, DEV, keep_edge_attr=keep)
    rng = np.random.default_rng(B)
    for trial in range(3):
        idx = rng.integers(0, n, size=B)          # repeats allowed, arbitrary order
        got = store.batch(idx, d_hat=True)
        ref = Batch.from_data_list([ds[int(i)] for i in idx])
        for k in ("x", "edge_index", "edge_weight", "batch", "u", "y", "d_hat"):
            a, b = getattr(got, k).cpu(), getattr(ref, k)
            assert a.dtype == b.dtype and a.shape == b.shape, k
            assert torch.equal(a, b), k
        if keep:
            assert torch.equal(got.edge_attr.cpu(), ref.edge_attr)
        else:
            # on-the-fly expansion: identical to the smearing kernel, within an ulp or two of the host's exp
            s = store.smear
            dev_ea = MF.gaussian_smear(got.d_hat, store.smear_offset, store.smear_coeff)
            assert torch.equal(got.edge_attr, dev_ea)
            assert (got.edge_attr.cpu() - ref.edge_attr).abs().max().item() < 2e-6
            assert s["resolution"] == ref.edge_attr.shape[1]
        # the carried layout == a fresh sort of the assembled COO
        csr = got.edge_index._mdl_csr[1]
        fresh = GraphCSR.from_coo(got.edge_index.clone(), got.batch.clone(), num_nodes=got.x.shape[0], num_graphs=B)
        for name in LAYOUT:
            assert torch.equal(getattr(csr, name), getattr(fresh, name)), name
        hit = got.edge_attr._mdl_slots
        assert hit[0] is csr
        assert torch.equal(hit[2], gather_rows(got.edge_attr, fresh.dst_eid))
        assert csr.to_slots(got.edge_attr) is hit[2]


def test_training_step_on_assembled_batch_equals_host_path():
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 24)
    store = GraphStore.from_dataset(ds, DEV, keep_edge_attr=True)
    idx = [3, 7, 1, 20, 11, 5, 9, 2]
    torch.manual_seed(0)
    model = M.CGCNN(ds, dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    b1 = store.batch(idx)
    b2 = ds.batch(idx).to(DEV)
    l1 = torch.nn.functional.l1_loss(m1(b1), b1.y)
    l2 = torch.nn.functional.l1_loss(m2(b2), b2.y)
    l1.backward()
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * max(1.0, abs(l2.item()))
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        scale = max(p2.grad.abs().max().item(), 1e-6)
        assert (p1.grad - p2.grad).abs().max().item() <= 2e-5 * scale, n1


def test_store_argument_checks():
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 5)
    store = GraphStore.from_dataset(ds, DEV)
    with pytest.raises(IndexError):
        store.batch([0, 5])
    with pytest.raises(ValueError):
        store.batch([])
    with pytest.raises(RuntimeError):
        GraphStore.from_dataset(ds, "cpu")
    b = store.batch([4, 4, 0], layout=False)
    assert not hasattr(b.edge_index, "_mdl_csr") and b.x.shape[0] == 2 * ds[4].x.shape[0] + ds[0].x.shape[0]

"""The reference's processed-dataset files (`processed/data.pt` = (data, slices) of InMemoryDataset.collate,
process.py:521-523; `data_{i}.pt`, process.py:525-532) load without torch_geometric.

torch_geometric is not installable here, so the files are produced with stand-in classes registered under
PyG's module paths: what reaches the disk is exactly what PyG writes -- a pickle that names
`torch_geometric.data.data.Data` (and, for PyG 2.x, `torch_geometric.data.storage.GlobalStorage`) and carries
the attribute dict -- in both internal layouts (1.x: attributes in __dict__; 2.x: _store._mapping)."""
import sys
import types

import pytest
import torch

from matdeeplearn_b200 import compat, process as pr


def _fake_pyg(layout):
    data_mod = types.ModuleType("torch_geometric.data.data")
    storage_mod = types.ModuleType("torch_geometric.data.storage")

    class GlobalStorage:
        def __init__(self, mapping):
            self._mapping = dict(mapping)

    GlobalStorage.__module__ = "torch_geometric.data.storage"
    GlobalStorage.__qualname__ = "GlobalStorage"

    class Data:
        def __init__(self, **kw):
            if layout == 2:
                self._store = GlobalStorage(kw)
            else:
                self.__dict__.update(kw)

    Data.__module__ = "torch_geometric.data.data"
    Data.__qualname__ = "Data"
    data_mod.Data = Data
    storage_mod.GlobalStorage = GlobalStorage
    pkg = types.ModuleType("torch_geometric")
    pkg_data = types.ModuleType("torch_geometric.data")
    mods = {"torch_geometric": pkg, "torch_geometric.data": pkg_data,
            "torch_geometric.data.data": data_mod, "torch_geometric.data.storage": storage_mod}
    return Data, mods


def _collate(graphs, keys):
    """InMemoryDataset.collate: concatenate along the attribute's cat dim, NO index offsets, slices per key."""
    data, slices = {}, {}
    for k in keys:
        vals = [getattr(g, k) for g in graphs]
        dim = 1 if k == "edge_index" else 0
        vals = [v.reshape(1, -1) if (k in ("y", "u") and v.dim() <= 1) else v for v in vals]
        data[k] = torch.cat(vals, dim)
        sizes = torch.tensor([0] + [v.shape[dim] for v in vals])
        slices[k] = torch.cumsum(sizes, 0)
    data["structure_id"] = [[f"id{i}"] for i in range(len(graphs))]
    slices["structure_id"] = torch.arange(len(graphs) + 1)
    return data, slices


@pytest.mark.parametrize("layout", [1, 2])
def test_data_pt_roundtrip(tmp_path, layout):
    ds = pr.synthetic_dataset("bulk", 7, seed=5)
    keys = ("x", "edge_index", "edge_attr", "edge_weight", "u", "y")
    Data, mods = _fake_pyg(layout)
    sys.modules.update(mods)
    try:
        d, s = _collate(ds.graphs, keys)
        torch.save((Data(**d), s), tmp_path / "data.pt")
        for i, g in enumerate(ds.graphs[:3]):
            torch.save(Data(**{k: (getattr(g, k).reshape(1, -1) if k == "y" else getattr(g, k)) for k in keys}),
                       tmp_path / f"data_{i}.pt")
    finally:
        for m in mods:
            sys.modules.pop(m, None)
    assert "torch_geometric" not in sys.modules
    got = compat.load_processed(str(tmp_path / "data.pt"))
    assert len(got) == len(ds)
    assert got.num_features == ds.num_features and got.num_edge_features == ds.num_edge_features
    for a, b in zip(got.graphs, ds.graphs):
        for k in keys:
            va, vb = getattr(a, k), getattr(b, k)
            assert torch.equal(va.reshape(vb.shape) if k in ("u", "y") else va, vb), k
    assert got[2].structure_id == ["id2"] or got[2].structure_id == "id2"
    # collated batch == the batch of the original dataset (what the models and the GraphStore consume)
    ba, bb = got.batch(), ds.batch()
    for k in ("x", "edge_index", "edge_attr", "edge_weight", "batch", "u", "y"):
        assert torch.equal(getattr(ba, k), getattr(bb, k)), k
    # directory form: data.pt wins; without it the data_{i}.pt files are read in index order
    assert len(compat.load_processed(str(tmp_path))) == len(ds)
    (tmp_path / "data.pt").unlink()
    large = compat.load_processed(str(tmp_path))
    assert len(large) == 3
    assert torch.equal(large[1].edge_index, ds[1].edge_index) and torch.equal(large[1].y, ds[1].y)
    # target_index = -1 keeps the stored [1, T] targets
    assert compat.load_processed(str(tmp_path), target_index=-1)[0].y.shape == (1, 1)


def test_load_processed_rejects_foreign_files(tmp_path):
    torch.save(({"a": torch.zeros(3)}, {"a": torch.tensor([0, 3])}), tmp_path / "data.pt")
    with pytest.raises(ValueError):
        compat.load_processed(str(tmp_path / "data.pt"))

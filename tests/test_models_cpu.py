"""CPU: the engine's model classes accept the reference's constructor keywords and
state_dicts (parameter names/shapes identical to the reference modules whose
fixtures were produced by tests/golden/make_golden.py), and refuse to run
without CUDA (no silent fallback)."""
import pytest
import torch

from tests.golden.make_golden import MODEL_CFGS
from tests.test_oracle_golden import _DS, load_batch, load_model_fixture


@pytest.mark.parametrize("tag", list(MODEL_CFGS))
def test_state_dict_compatible_with_reference(tag):
    from matdeeplearn_b200 import models as M
    b = load_batch()
    _, sd, _ = load_model_fixture(tag)
    model = getattr(M, tag.split("_")[0])(_DS(b), **MODEL_CFGS[tag])
    res = model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()},
                                strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_models_refuse_cpu_batches():
    from matdeeplearn_b200 import models as M
    b = load_batch()
    model = M.CGCNN(_DS(b), **MODEL_CFGS["CGCNN"])
    with pytest.raises(RuntimeError):
        model(b)


def test_model_lookup_by_name_like_reference():
    # reference training/training.py:250: getattr(models, model_name)(data=dataset, **params)
    from matdeeplearn_b200 import models as M
    for name in ("CGCNN", "SchNet", "MPNN", "MEGNet", "GCN"):
        assert callable(getattr(M, name))

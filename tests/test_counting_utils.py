"""CPU: known answers for host-side helpers of the counting path -- the row-wise activation choices of --rep_act_func /
--pred_act_func that torch does not ship (subgraph_isomorphism/utils/act.py:210-455), expand_dimensions, the dummy column
of subisomorphisms, the lazy batch dict.  The comparisons with the reference's own functions are the live tests in
tests/test_oracle_vs_reference.py."""
import pytest
import torch

from dummynode4graphlearning_b200.subgraph_isomorphism.utils import (Maximum, Minimum, Sparsemax, map_activation_str_to_layer,
                                                                     supported_act_funcs)


def test_sparsemax_known_answers():
    sp = Sparsemax(dim=-1)
    x = torch.tensor([[0.1, 1.1, 0.2], [5.0, 0.0, 0.0], [0.5, 0.5, 0.5], [1.0, 1.0, 0.5]])
    y = sp(x)
    # (0.1, 1.1, 0.2): support {1.1, 0.2}, tau = 0.15 -> (0, 0.95, 0.05); a dominant logit takes everything; ties split
    assert torch.allclose(y, torch.tensor([[0.0, 0.95, 0.05], [1.0, 0.0, 0.0], [1 / 3, 1 / 3, 1 / 3], [0.5, 0.5, 0.0]]), atol=1e-6)
    assert torch.allclose(y.sum(-1), torch.ones(4), atol=1e-6) and bool((y >= 0).all())
    # shift invariance, and other axes / ranks
    assert torch.allclose(sp(x + 3.0), y, atol=1e-6)
    z = torch.randn(3, 4, 5, generator=torch.Generator().manual_seed(0))
    for dim in (0, 1, 2):
        out = Sparsemax(dim=dim)(z)
        assert out.shape == z.shape and torch.allclose(out.sum(dim), torch.ones_like(out.sum(dim)), atol=1e-5)


def test_maximum_minimum_known_answers():
    x = torch.tensor([[1.0, 3.0, 3.0, -2.0], [-1.0, -4.0, 0.5, -4.0]])
    assert Maximum()(x).tolist() == [[0.0, 3.0, 3.0, 0.0], [0.0, 0.0, 0.5, 0.0]]
    assert Minimum()(x).tolist() == [[0.0, 0.0, 0.0, -2.0], [0.0, -4.0, 0.0, -4.0]]
    up = Maximum(scale_up=True)(x)                      # kept entries rescaled to the row's sum: 5 / 6 and -8.5 / 0.5
    assert torch.allclose(up, torch.tensor([[0.0, 2.5, 2.5, 0.0], [0.0, 0.0, -8.5, 0.0]]))
    assert Maximum(scale_up=True)(torch.zeros(1, 3)).tolist() == [[0.0, 0.0, 0.0]]      # 0 / 0 -> 0, not nan
    y = x.clone()
    assert Minimum(inplace=True)(y) is y and y.tolist() == Minimum()(x).tolist()


def test_registry_is_complete_and_shared():
    names = {"none", "softmax", "sparsemax", "gumbel_softmax", "sigmoid", "tanh", "relu", "relu6", "leaky_relu", "prelu", "elu",
             "celu", "selu", "gelu", "maximum", "minimum"}
    assert set(supported_act_funcs) == names                                             # utils/act.py:457-473
    assert map_activation_str_to_layer("relu") is map_activation_str_to_layer("relu")   # shared singletons (App. A-3)
    assert abs(map_activation_str_to_layer("leaky_relu").negative_slope - 1 / 5.5) < 1e-12
    g = map_activation_str_to_layer("gumbel_softmax")(torch.zeros(5, 7))
    assert torch.allclose(g.sum(-1), torch.ones(5), atol=1e-6)


def test_expand_dimensions_known_answer():
    """utils/dl.py:157-195: old values land in the trailing corner (pre_pad) or the leading one, zeros elsewhere;
    parameters that only the new module has keep their initialisation."""
    import torch.nn as nn
    from dummynode4graphlearning_b200.subgraph_isomorphism.utils import expand_dimensions
    old, new = torch.arange(1.0, 7.0).view(2, 3), torch.full((3, 5), 9.0)
    expand_dimensions(old, new, pre_pad=True)
    assert new.tolist() == [[0, 0, 0, 0, 0], [0, 0, 1, 2, 3], [0, 0, 4, 5, 6]]
    new = torch.full((3, 5), 9.0)
    expand_dimensions(old, new, pre_pad=False)
    assert new.tolist() == [[1, 2, 3, 0, 0], [4, 5, 6, 0, 0], [0, 0, 0, 0, 0]]
    a, b = nn.ModuleDict({"x": nn.Linear(2, 2)}), nn.ModuleDict({"x": nn.Linear(4, 2), "y": nn.Linear(3, 1)})
    keep = b["y"].weight.clone()
    expand_dimensions(a, b)
    assert torch.equal(b["x"].weight[:, 2:], a["x"].weight) and float(b["x"].weight.detach()[:, :2].abs().sum()) == 0.0
    assert torch.equal(b["x"].bias, a["x"].bias) and torch.equal(b["y"].weight, keep)


def test_add_dummy_to_subisomorphisms():
    """train.py:437-443: the dummy column is the graph's ORIGINAL node count; samples without matches stay empty."""
    import numpy as np
    from dummynode4graphlearning_b200.subgraph_isomorphism.matching import add_dummy_to_subisomorphisms, pack_subisomorphisms
    g = dict(node_ptr=np.array([0, 5, 12, 15], dtype=np.int32))
    mats = [np.array([[0, 1], [3, 4]]), np.zeros((0, 3), dtype=np.int64), np.array([[2, 1, 0]])]
    out = add_dummy_to_subisomorphisms(mats, g)
    assert out[0].tolist() == [[0, 1, 5], [3, 4, 5]] and out[1].shape == (0, 4) and out[2].tolist() == [[2, 1, 0, 3]]
    packed = pack_subisomorphisms(out)
    assert packed["val_ptr"].tolist() == [0, 6, 6, 10] and packed["rows"].tolist() == [2, 0, 1]


def test_lazy_batch_dict_semantics():
    """transforms.LazyDict: derived columns are computed once, on first access; membership sees them before that;
    pop() discards a pending column WITHOUT computing it (the pipeline drops columns the model does not read);
    iteration / dict(...) materialise everything."""
    from dummynode4graphlearning_b200.transforms import LazyDict
    calls = []

    def make(name, value):
        def fn():
            calls.append(name)
            return value
        return fn

    d = LazyDict(a=1)
    d.lazy("b", make("b", 2))
    d.lazy("c", make("c", 3))
    d.lazy("e", make("e", 5))
    assert "b" in d and "zzz" not in d and calls == []
    assert d["b"] == 2 and d["b"] == 2 and calls == ["b"]            # computed once, then stored
    assert d.get("c") == 3 and d.get("zzz", 7) == 7 and calls == ["b", "c"]
    assert d.pop("e", None) is None and "e" not in d and calls == ["b", "c"]   # discarded, never computed
    with pytest.raises(KeyError):
        d["e"]
    d.lazy("f", make("f", 6))
    assert dict(d) == {"a": 1, "b": 2, "c": 3, "f": 6} and calls == ["b", "c", "f"]
    assert sorted(d.keys()) == ["a", "b", "c", "f"] and len(list(d.items())) == 4

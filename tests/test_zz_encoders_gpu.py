"""GPU: the counting models with ``--enc_net Position`` (sinusoid tables, subgraph_isomorphism/models/embed.py:211-222,
basemodel.py:642-646) against the oracle -- forward tensors, loss and every parameter gradient, same procedure and
tolerance as the Multihot cases of test_models_gpu.py.  (The oracle itself is pinned to the unmodified reference with
this encoder in tests/test_oracle_vs_reference.py::test_counting_models_live[live/*_position].)"""
import pytest

from test_models_gpu import test_counting_models_match_oracle_live as _run_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,shape,bs,over", [
    ("RGIN", "small", 48, dict(enc_net="Position")),
    ("DMPNN", "small", 48, dict(enc_net="Position", node_pred=True, edge_pred=True, pred_return_weights="node,edge")),
])
def test_position_encoder_models_match_oracle(device, name, shape, bs, over):
    _run_against_oracle(device, name, shape, bs, over)

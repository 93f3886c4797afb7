"""GPU parity of the tier-2 token transformer drop-in (models/transformer.py) against fixtures from the REAL reference:
logits, cross-entropy loss and every parameter gradient (tiny: fp32 SIMT path; wide: tensor-core Linear layers)."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,tol_f,tol_g", [("tiny", 1e-3, 5e-3), ("wide", 1e-3, 5e-3)])
def test_make_a_scene_forward_backward_vs_reference(tag, tol_f, tol_g):
    from models.transformer import MakeAScene
    g = torch.load(os.path.join(GOLDEN, f"transformer_{tag}.pt"), weights_only=False)
    dev = torch.device("cuda:0")
    m = MakeAScene(**g["cfg"])
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
    m.load_state_dict(g["state_dict"])
    m.to(dev)
    m.device = dev
    logits = m(g["text"].to(dev), g["seg"].to(dev), g["img"].to(dev))
    assert logits.shape == g["logits"].shape
    assert rel_err(logits, g["logits"]) < tol_f
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), g["img"].to(dev).reshape(-1))
    assert abs(float(loss) - float(g["loss"])) < tol_f * max(1.0, float(g["loss"]))
    loss.backward()
    named = dict(m.named_parameters())
    for k, gv in g["grads"].items():
        assert named[k].grad is not None, k
        assert rel_err(named[k].grad, gv) < tol_g, (tag, k)
    for k, v in g["grad_norms"].items():
        got = float(named[k].grad.double().norm())
        assert abs(got - v) <= 2e-2 * max(v, 1e-4 * named[k].numel() ** 0.5), (tag, k, got, v)

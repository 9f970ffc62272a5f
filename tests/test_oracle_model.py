"""CPU: oracle/model.py (torch restatement) against golden vectors produced by the executed
reference Model (tests/golden/make_golden_model.py): outputs, loss, gradient fingerprints,
BatchNorm running statistics, the in-place input shift, and the no-grad position table."""
import os
import random

import numpy as np
import pytest
import torch

from make_golden_model import CASES, grad_fingerprint, make_input, scalar_loss
from oracle import model as om


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "model_golden.npz"))


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_oracle_model_matches_reference(golden, ci):
    name, D, NL, B, L, pyseed, inp = CASES[ci]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
              else v.clone()) for k, v in om.formula_state_dict(D, NL).items()}
    x = make_input(B, L, inp)
    with torch.no_grad():
        pred, aux = om.model_forward(sd, x.clone(), training=False)
    assert pred.shape == (B, (L + 7) // 8, 80) and aux.shape == (B, (L + 7) // 8, 48)
    np.testing.assert_allclose(pred.numpy(), golden[f"{name}_eval_pred"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(aux.numpy(), golden[f"{name}_eval_aux"], rtol=0, atol=2e-6)

    random.seed(pyseed)
    xt = x.clone()
    pred, aux = om.model_forward(sd, xt, training=True, dropout_p=0.0)
    loss = scalar_loss(pred, aux)
    loss.backward()
    np.testing.assert_allclose(pred.detach().numpy(), golden[f"{name}_train_pred"], atol=2e-6)
    np.testing.assert_array_equal(xt[:, -9:, :].numpy(), golden[f"{name}_train_x_after"])
    assert abs(loss.item() - float(golden[f"{name}_train_loss"])) < 1e-6
    grads = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    fp = grad_fingerprint(grads)
    n_checked = 0
    for key in golden.files:
        if key.startswith(f"{name}_grad::"):
            k = key.split("::", 1)[1]
            np.testing.assert_allclose(fp[k], golden[key], rtol=1e-4, atol=1e-7, err_msg=k)
            n_checked += 1
    assert n_checked >= 40
    for k, v in sd.items():
        if k.endswith("relative_positional.embeddings"):
            assert v.grad is None
        if "running_" in k:
            np.testing.assert_allclose(v.numpy(), golden[f"{name}_buf::{k}"], rtol=1e-6, atol=1e-7)

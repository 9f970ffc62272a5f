"""CPU: the numpy CTC oracle (oracle/ctc.py) against golden vectors produced by executing the
reference's formulation F.ctc_loss(F.log_softmax(pred, 2), ...) (recognition_model.py:96-101) in
float64 (tests/golden/make_golden_ctc.py): per-utterance nll, the 'mean'-reduced loss, and both
gradients w.r.t. the logits, with ragged lengths, empty targets and repeated labels."""
import os

import numpy as np

from oracle import ctc as octc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_golden.npz"))


def test_oracle_ctc_matches_executed_reference():
    for i in range(int(G["n_cases"])):
        g = {k: G[f"c{i}_{k}"] for k in ("logits", "targets", "il", "tl", "nll", "loss",
                                         "grad_mean", "grad_nll")}
        blank = g["logits"].shape[2] - 1
        nll, grad = octc.ctc_nll_and_grad(g["logits"], g["targets"], g["il"], g["tl"], blank)
        assert np.allclose(nll, g["nll"], rtol=1e-10, atol=1e-10), i
        assert np.abs(grad - g["grad_nll"]).max() < 1e-9, i
        loss, gm = octc.ctc_loss_mean(g["logits"], g["targets"], g["il"], g["tl"], blank)
        assert abs(loss - float(g["loss"])) < 1e-10 * max(1.0, abs(loss)), i
        assert np.abs(gm - g["grad_mean"]).max() < 1e-10, i
        for n in range(len(nll)):                       # padding frames carry no gradient
            assert (grad[n, int(g["il"][n]):] == 0).all()


def test_oracle_ctc_closed_forms():
    # one frame, one symbol: p(l) = softmax(x)[l]
    x = np.array([[[0.3, -1.2, 0.5]]])
    nll, grad = octc.ctc_nll_and_grad(x, [[1]], [1], [1], blank=2)
    p = np.exp(x[0, 0]) / np.exp(x[0, 0]).sum()
    assert abs(nll[0] + np.log(p[1])) < 1e-12
    assert np.allclose(grad[0, 0], p - np.array([0, 1, 0]))
    # empty target: only blanks
    x = np.random.RandomState(0).randn(1, 5, 4)
    nll, _ = octc.ctc_nll_and_grad(x, [[0]], [5], [0], blank=3)
    lp = x[0] - np.log(np.exp(x[0]).sum(1, keepdims=True))
    assert abs(nll[0] + lp[:, 3].sum()) < 1e-12
    # infeasible: "aa" needs 3 frames
    nll, grad = octc.ctc_nll_and_grad(np.zeros((1, 2, 3)), [[0, 0]], [2], [2], blank=2)
    assert np.isinf(nll[0]) and (grad == 0).all()

"""GPU parity of the loss + training step (silent_speech_b200.losses / training) against the
CPU oracle step (oracle/step.py, which follows transduction_model.py:98-157,196-212):
same synthetic collate_raw batch, same formula weights, dropout 0."""
import random

import numpy as np
import pytest
import torch
from absl import flags

from oracle import model as om
from oracle import step as ostep
from silent_speech_b200.read_emg import synthetic_batch

pytestmark = pytest.mark.gpu


def build(D, NL):
    from silent_speech_b200 import architecture as A
    F = flags.FLAGS
    if not F.is_parsed():
        F(["test"])
    F.model_size, F.num_layers, F.dropout = D, NL, 0.0
    m = A.Model(112, 80, 48)
    m.load_state_dict(om.formula_state_dict(D, NL), strict=True)
    return m.cuda().train()


def clone_batch(b):
    return {k: ([t.clone() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else v)
            for k, v in b.items()}


@pytest.mark.parametrize("frames", [40, 130])
def test_dtw_loss_matches_oracle(frames):
    from silent_speech_b200.losses import dtw_loss
    batch = synthetic_batch(6, frames, seed=3)
    g = torch.Generator().manual_seed(0)
    pred = torch.randn(6, frames, 80, generator=g)
    phon = torch.randn(6, frames, 48, generator=g)
    want = ostep.dtw_loss(pred.clone().requires_grad_(True), phon, batch)
    p2 = pred.clone().cuda().requires_grad_(True)
    got, acc = dtw_loss(p2, phon.cuda(), batch, phoneme_loss_weight=0.5)
    assert abs(got.item() - want.item()) < 2e-5 * abs(want.item())
    got.backward()
    assert torch.isfinite(p2.grad).all()
    got2, acc2 = dtw_loss(p2.detach(), phon.cuda(), batch, phoneme_eval=True,
                          phoneme_confusion=np.zeros((48, 48)), phoneme_loss_weight=0.5)
    assert 0.0 <= acc2 <= 1.0 and abs(got2.item() - got.item()) < 1e-6


def test_train_step_matches_oracle(monkeypatch):
    monkeypatch.setenv("SSB_GEMM", "simt")   # semantics test: exact-fp32 engine (AdamW divides by
    #                                           sqrt(v): kink-induced gradient noise would dominate)
    from silent_speech_b200.training import GradientBucket, train_step
    D, NL, frames, n = 32, 1, 130, 4
    batch = synthetic_batch(n, frames, seed=11)
    m = build(D, NL)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-7)
    bucket = GradientBucket(m)
    params = ostep.make_params(om.formula_state_dict(D, NL))
    oopt = ostep.make_optimizer(params)
    for it in range(2):
        random.seed(100 + it)
        l_gpu = train_step(m, opt, clone_batch(batch), "cuda", frames, bucket)
        random.seed(100 + it)
        l_cpu = ostep.train_step(params, oopt, clone_batch(batch), frames)
        assert abs(l_gpu - l_cpu) < 1e-4 * abs(l_cpu), (it, l_gpu, l_cpu)
    worst = 0.0
    for k, p in m.named_parameters():
        if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias",
                                                       "residual_path.bias")):
            continue   # zero-gradient parameters: Adam turns their rounding noise into +-lr steps
        a, b = p.detach().cpu().double(), params[k].detach().double()
        worst = max(worst, ((a - b).norm() / (b.norm() + 1e-30)).item())
    assert worst < 1e-4, worst

"""Fused log_softmax + CTC (csrc/ctc.cu) against the reference's formulation
F.ctc_loss(F.log_softmax(pred, 2)...) (recognition_model.py:96-101): loss and gradient w.r.t. the
logits, ragged input / target lengths, repeated labels, empty and infeasible targets."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def reference(logits, targets, il, tl, blank, reduction):
    x = logits.detach().double().requires_grad_(True)        # fp64 reference of the same formula
    lp = F.log_softmax(x, 2).transpose(0, 1)
    loss = F.ctc_loss(lp, targets, il, tl, blank=blank, reduction=reduction)
    (loss.sum() if reduction == 'none' else loss).backward()
    return loss.detach(), x.grad


def case(N, T, C, Lmax, seed, repeat=False):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(N, T, C, generator=g) * 2).cuda()
    hi = 3 if repeat else C - 1                              # tiny alphabet -> many repeated labels
    targets = torch.randint(0, hi, (N, Lmax), generator=g).cuda()
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g)
    il[0] = T
    tl = torch.randint(1, Lmax + 1, (N,), generator=g)
    tl = torch.minimum(tl, il // 2)                          # feasible even with repeats
    return logits, targets, il, tl


@pytest.mark.parametrize("N,T,C,Lmax,repeat", [(3, 50, 38, 12, False), (4, 163, 38, 40, True),
                                               (2, 750, 38, 200, False), (5, 33, 5, 16, True),
                                               (1, 7, 3, 1, False), (32, 750, 38, 150, False)])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_ctc_matches_torch(N, T, C, Lmax, repeat, reduction):
    from silent_speech_b200.losses import ctc_loss
    logits, targets, il, tl = case(N, T, C, Lmax, seed=N * 1000 + T, repeat=repeat)
    blank = C - 1
    x = logits.clone().requires_grad_(True)
    loss = ctc_loss(x, targets, il, tl, blank=blank, reduction=reduction)
    (loss.sum() if reduction == 'none' else loss).backward()
    want, gwant = reference(logits, targets, il, tl, blank, reduction)
    assert torch.allclose(loss.double(), want, rtol=2e-5, atol=1e-5), (loss, want)
    # gradient = softmax - occupancy cancels, and the fp32 log-space recursion carries ~1e-5
    # absolute noise after hundreds of steps (torch's own fp32 ctc_loss sits at the same level
    # against this fp64 reference); the bar of the north star is 1e-3
    rel = ((x.grad.double() - gwant).norm() / gwant.norm()).item()
    assert rel < 3e-4, rel
    for n in range(N):                                        # padding frames get exactly zero
        assert (x.grad[n, int(il[n]):] == 0).all()


def test_ctc_empty_and_infeasible_targets():
    from silent_speech_b200.losses import ctc_loss
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(3, 20, 6, generator=g).cuda()
    targets = torch.tensor([[0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1], [1] * 12, [2] * 12]).cuda()
    il = torch.tensor([20, 20, 5])
    tl = torch.tensor([0, 12, 4])            # empty; 12 repeats need 23 frames > 20; 4 repeats need 7 > 5
    x = logits.clone().requires_grad_(True)
    nll = ctc_loss(x, targets, il, tl, blank=5, reduction='none')
    want, gwant = reference(logits[:1], targets[:1], il[:1], tl[:1], 5, 'none')
    assert torch.allclose(nll[:1].double(), want, rtol=2e-5)
    assert torch.isinf(nll[1]) and torch.isinf(nll[2])
    nll[0].backward()
    assert ((x.grad[0].double() - gwant[0]).norm() / gwant[0].norm()).item() < 3e-4
    assert (x.grad[1:] == 0).all()


def test_ctc_through_the_model_head():
    """The recognition step (recognition_model.py:89-101) with the fused loss vs stock torch."""
    import random

    from absl import flags
    from silent_speech_b200 import architecture as A
    from silent_speech_b200.losses import ctc_loss
    FL = flags.FLAGS
    if not FL.is_parsed():
        FL(["t"])
    FL.model_size, FL.num_layers, FL.dropout = 64, 1, 0.0
    torch.manual_seed(0)
    m = A.Model(112, 38).cuda().train()
    x = torch.randn(2, 1304, 8, generator=torch.Generator().manual_seed(1)).cuda()
    tgt = torch.randint(0, 37, (2, 20), generator=torch.Generator().manual_seed(2)).cuda()
    il, tl = torch.tensor([163, 150]), torch.tensor([20, 17])
    grads = []
    for fused in (True, False):
        m.zero_grad()
        random.seed(4)
        out = m(None, x.clone(), None)
        if fused:
            loss = ctc_loss(out, tgt, il, tl, blank=37)
        else:
            loss = F.ctc_loss(F.log_softmax(out, 2).transpose(0, 1), tgt, il, tl, blank=37)
        loss.backward()
        grads.append((loss.item(), m.w_out.weight.grad.clone()))
    assert abs(grads[0][0] - grads[1][0]) < 1e-4 * abs(grads[1][0])
    assert ((grads[0][1] - grads[1][1]).norm() / grads[1][1].norm()).item() < 1e-3


def test_ctc_kernel_matches_golden_fixture_and_oracle():
    """The committed fixture (tests/golden/ctc_golden.npz: the reference's formulation executed in
    fp64 by make_golden_ctc.py) and the numpy oracle on the same inputs."""
    import os

    import numpy as np
    from oracle import ctc as octc
    from silent_speech_b200.losses import ctc_loss
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_golden.npz"))
    for i in range(int(G["n_cases"])):
        logits = torch.from_numpy(G[f"c{i}_logits"]).cuda().requires_grad_(True)
        targets = torch.from_numpy(G[f"c{i}_targets"]).cuda()
        il, tl = torch.from_numpy(G[f"c{i}_il"]), torch.from_numpy(G[f"c{i}_tl"])
        blank = logits.shape[2] - 1
        loss = ctc_loss(logits, targets, il, tl, blank=blank)
        loss.backward()
        want, gwant = float(G[f"c{i}_loss"]), G[f"c{i}_grad_mean"]
        assert abs(loss.item() - want) < 2e-5 * abs(want), (i, loss.item(), want)
        got = logits.grad.double().cpu().numpy()
        assert np.linalg.norm(got - gwant) / np.linalg.norm(gwant) < 3e-4, i
        oloss, ograd = octc.ctc_loss_mean(G[f"c{i}_logits"], G[f"c{i}_targets"], G[f"c{i}_il"],
                                          G[f"c{i}_tl"], blank)
        assert abs(loss.item() - oloss) < 2e-5 * abs(oloss)
        assert np.linalg.norm(got - ograd) / np.linalg.norm(ograd) < 3e-4

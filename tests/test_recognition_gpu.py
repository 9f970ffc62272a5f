"""GPU: the recognition configuration (recognition_model.py:66,96-101): the same Model with a
single 38-way head (no aux output) under stock CTC loss, forward + backward; and eval-mode
un-chunked inference of a whole utterance (transduction_model.py:60-64)."""
import random

import pytest
import torch
import torch.nn.functional as F
from absl import flags

from oracle import model as om

pytestmark = pytest.mark.gpu


def test_ctc_head_forward_backward_matches_oracle():
    from silent_speech_b200 import architecture as A
    FL = flags.FLAGS
    if not FL.is_parsed():
        FL(["t"])
    FL.model_size, FL.num_layers, FL.dropout = 64, 2, 0.0
    torch.manual_seed(0)
    m = A.Model(112, 38)                               # num_aux_outs=None -> single tensor out
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    x = torch.randn(2, 1304, 8, generator=torch.Generator().manual_seed(1))
    tgt = torch.randint(0, 37, (2, 20), generator=torch.Generator().manual_seed(2))
    random.seed(4)
    out = m(None, x.clone().cuda(), None)
    assert isinstance(out, torch.Tensor) and out.shape == (2, 163, 38)
    lp = F.log_softmax(out, 2).transpose(0, 1)          # (T, N, C) as recognition_model.py:97-100
    loss = F.ctc_loss(lp, tgt.cuda(), torch.tensor([163, 163]), torch.tensor([20, 17]), blank=37)
    loss.backward()
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
              else v.clone()) for k, v in sd0.items()}
    random.seed(4)
    o2 = om.model_forward(sd, x.clone(), training=True, dropout_p=0.0)
    l2 = F.ctc_loss(F.log_softmax(o2, 2).transpose(0, 1), tgt, torch.tensor([163, 163]),
                    torch.tensor([20, 17]), blank=37)
    l2.backward()
    assert abs(loss.item() - l2.item()) < 1e-3 * abs(l2.item())
    g, r = m.w_out.weight.grad.cpu(), sd["w_out.weight"].grad
    assert ((g - r).norm() / r.norm()).item() < 2e-2


def test_eval_whole_utterance_inference():
    from silent_speech_b200 import architecture as A
    FL = flags.FLAGS
    if not FL.is_parsed():
        FL(["t"])
    FL.model_size, FL.num_layers, FL.dropout = 64, 1, 0.2
    torch.manual_seed(3)
    m = A.Model(112, 80, 48).cuda().eval()
    x = torch.randn(1, 2777, 8).cuda()                  # odd length, T = 348
    with torch.no_grad():
        a, b = m(None, x, None)
        a2, _ = m(None, x, None)
    assert a.shape == (1, 348, 80) and b.shape == (1, 348, 48)
    assert torch.equal(a, a2)                           # eval: no dropout, deterministic

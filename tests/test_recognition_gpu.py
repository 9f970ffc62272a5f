"""GPU: the recognition configuration (recognition_model.py:66,96-101): the same Model with a
single 38-way head (no aux output) under stock CTC loss, forward + backward; and eval-mode
un-chunked inference of a whole utterance (transduction_model.py:60-64)."""
import math
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


def test_cfg5_shape_ctc_step_engines_agree(monkeypatch):
    """BASELINE cfg-5 geometry (768-dim model, L = 6000 -> T = 750, 38-way head, CTC) at reduced
    batch / depth: one forward + fused CTC + backward on the tensor-core path (fused attention with
    a ragged last key tile: 750 = 5 * 128 + 110) against the CUDA-core attention engine on the same
    weights, and against stock F.ctc_loss on the same logits."""
    from silent_speech_b200 import architecture as A
    from silent_speech_b200.losses import ctc_loss
    FL = flags.FLAGS
    if not FL.is_parsed():
        FL(["t"])
    FL.model_size, FL.num_layers, FL.dropout = 768, 2, 0.0
    torch.manual_seed(0)
    m = A.Model(112, 38).cuda().train()
    x = torch.randn(2, 6000, 8, generator=torch.Generator().manual_seed(1)).cuda()
    tgt = torch.randint(0, 37, (2, 120), generator=torch.Generator().manual_seed(2)).cuda()
    il, tl = torch.tensor([750, 700]), torch.tensor([120, 95])
    res = {}
    for eng in ("fused", "simt"):
        monkeypatch.setenv("SSB_ATTN", eng)
        m.zero_grad()
        random.seed(4)
        out = m(None, x.clone(), None)
        assert out.shape == (2, 750, 38)
        loss = ctc_loss(out, tgt, il, tl, blank=37)
        loss.backward()
        res[eng] = (out.detach(), loss.item(), m.w_out.weight.grad.clone(),
                    m.transformer.layers[0].self_attn.w_q.grad.clone())
    stock = F.ctc_loss(F.log_softmax(res["fused"][0].double(), 2).transpose(0, 1), tgt, il, tl,
                       blank=37).item()
    assert abs(res["fused"][1] - stock) < 1e-4 * abs(stock)
    assert math.isfinite(res["fused"][1])
    a, b = res["fused"], res["simt"]
    assert ((a[0] - b[0]).norm() / b[0].norm()).item() < 2e-4
    assert abs(a[1] - b[1]) < 1e-4 * abs(b[1])
    for i in (2, 3):
        assert ((a[i] - b[i]).norm() / b[i].norm()).item() < 2e-2

"""Fused flat AdamW (csrc/optim.cu, silent_speech_b200/optim.py) against
torch.optim.AdamW with the reference's hyper-parameters and warm-up schedule
(transduction_model.py:178-189,210)."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(37, 53)            # odd sizes: exercises the n % 4 tail
        self.b = nn.Conv1d(5, 7, 3)
        self.c = nn.LayerNorm(53)
        self.d = nn.Parameter(torch.randn(3))


def _grads(step, params):
    g = torch.Generator(device="cpu").manual_seed(100 + step)
    return [torch.randn(p.shape, generator=g).cuda() * (10.0 ** (step % 3 - 1)) for p in params]


@pytest.mark.parametrize("world", [1, 4])
def test_flat_adamw_matches_torch(world):
    from silent_speech_b200.optim import FlatAdamW
    from silent_speech_b200.training import GradientBucket
    torch.manual_seed(0)
    m_ref = Toy().cuda()
    m = copy.deepcopy(m_ref)
    keys = list(m.state_dict().keys())
    o_ref = torch.optim.AdamW(m_ref.parameters(), weight_decay=1e-7)   # lr 1e-3 default
    bucket = GradientBucket(m)
    o = FlatAdamW(bucket, weight_decay=1e-7)
    assert list(m.state_dict().keys()) == keys                          # checkpoint contract
    assert all(p.data_ptr() >= o.flat_param.data_ptr() for p in m.parameters())
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(o, 'min', 0.5, patience=5)   # accepted
    o.grad_scale = 1.0 / world
    for it in range(8):
        lr = (it + 1) * 1e-3 / 5 if it < 5 else 1e-3                    # linear warm-up
        for opt in (o_ref, o):
            for group in opt.param_groups:                              # the reference's set_lr
                group['lr'] = lr
        o.zero_grad()
        gs = _grads(it, list(m_ref.parameters()))
        for p, q, g in zip(m_ref.parameters(), m.parameters(), gs):
            p.grad = g.clone()
            q.grad.add_(g * world)        # the bucket holds the SUM over ranks
        o_ref.step()
        o.step()
        for (k, p), q in zip(m_ref.named_parameters(), m.parameters()):
            err = (p - q).abs().max().item()
            assert err <= 2e-6 * max(1.0, p.abs().max().item()), (it, k, err)
    assert int(o.step_cell.item()) == 8
    sd = o.state_dict()
    o.load_state_dict(sd)
    sched.step(1.0)
    assert sd["step"] == 8 and sd["exp_avg"].shape == o.flat_param.shape


def test_flat_adamw_in_train_step_matches_torch_adamw():
    """Whole training steps: FlatAdamW vs torch AdamW on the same model, batches and RNG."""
    import random

    from absl import flags
    from oracle import model as om
    from silent_speech_b200 import architecture as A
    from silent_speech_b200.optim import FlatAdamW
    from silent_speech_b200.read_emg import synthetic_batch
    from silent_speech_b200.training import GradientBucket, train_step
    F = flags.FLAGS
    if not F.is_parsed():
        F(["test"])
    F.model_size, F.num_layers, F.dropout = 32, 1, 0.0
    m1 = A.Model(112, 80, 48)
    m1.load_state_dict(om.formula_state_dict(32, 1), strict=True)
    m1 = m1.cuda().train()
    m2 = copy.deepcopy(m1)
    b1, b2 = GradientBucket(m1), GradientBucket(m2)
    o1 = torch.optim.AdamW(m1.parameters(), lr=1e-4, weight_decay=1e-7)
    o2 = FlatAdamW(b2, lr=1e-4, weight_decay=1e-7)
    for it in range(3):
        batch = synthetic_batch(4, 130, seed=40 + it)
        random.seed(it)
        l1 = train_step(m1, o1, batch, "cuda", 130, b1)
        random.seed(it)
        l2 = train_step(m2, o2, batch, "cuda", 130, b2)
        assert abs(l1 - l2) <= 1e-4 * abs(l1), (it, l1, l2)
    # Adam's 1/sqrt(v) amplifies the run-to-run rounding noise of near-zero gradients, so the
    # parameters are compared by update direction, not bit pattern
    for (k, p), q in zip(m1.named_parameters(), m2.parameters()):
        if k.endswith("relative_positional.embeddings"):
            assert torch.equal(p, q)
            continue
        assert (p - q).abs().max().item() <= 3.5e-4, k     # <= ~1 lr-sized step of disagreement

"""GPU: the WeightPlanes arena (weights.py + csrc/prep.cu) against the per-use derivation it
replaces.  Same model, same input: the forward must be BIT-identical (the operand planes hold the
same values either way) and gradients equal up to the split-K atomics' summation order; the arena
must follow optimiser updates (torch optimisers via tensor versions, FlatAdamW via the epoch) and
checkpoint loads."""
import random

import pytest
import torch
from absl import flags

pytestmark = pytest.mark.gpu


def _model(D, NL):
    from silent_speech_b200 import architecture as A
    F = flags.FLAGS
    if not F.is_parsed():
        F(["t"])
    F.model_size, F.num_layers, F.dropout = D, NL, 0.0
    torch.manual_seed(0)
    return A.Model(112, 80, 48).cuda().train()


def _run(m, x, monkeypatch, wplanes, blocks=False):
    monkeypatch.setenv("SSB_WPLANES", "1" if wplanes else "0")
    monkeypatch.setenv("SSB_BLOCKS", "1" if blocks else "0")
    m.zero_grad(set_to_none=True)
    random.seed(3)
    pred, aux = m(None, x.clone(), None)
    (pred.square().mean() + aux.square().mean()).backward()
    return pred.detach().clone(), aux.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()
                                                          if p.grad is not None}


@pytest.mark.parametrize("D,NL,L", [(128, 2, 1200), (768, 1, 4000)])
def test_arena_equals_per_use_derivation(D, NL, L, monkeypatch):
    m = _model(D, NL)
    x = torch.randn(2, L, 8, generator=torch.Generator().manual_seed(1)).cuda()
    p0, a0, g0 = _run(m, x, monkeypatch, False)
    p1, a1, g1 = _run(m, x, monkeypatch, True)
    assert m._wp is not None and m._wp.n_entries > 10
    # ... and with each half of an encoder layer as one autograd node (_AttnBlockFn / _FFNBlockFn)
    p2, a2, g2 = _run(m, x, monkeypatch, True, blocks=True)
    for p_, a_, g_ in ((p1, a1, g1), (p2, a2, g2)):
        assert torch.equal(p0, p_) and torch.equal(a0, a_)
        assert set(g0) == set(g_)
        for k in g0:
            if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias", "residual_path.bias")):
                # analytically zero in front of a training BatchNorm: rounding noise on the per-use
                # path, an exact zero (column sum skipped) on the arena path
                assert g0[k].abs().max() < 1e-4 and g_[k].abs().max() < 1e-4
                continue
            # same arithmetic, different summation orders (split-K atomics; the stacked heads'
            # data gradient is one K = 128 GEMM instead of two CUDA-core GEMMs and an add)
            d = (g0[k] - g_[k]).norm() / (g0[k].norm() + 1e-30)
            assert d < 5e-5, (k, d.item())


def test_arena_follows_optimizer_steps_and_checkpoint_loads(monkeypatch):
    from silent_speech_b200.optim import FlatAdamW
    from silent_speech_b200.training import GradientBucket
    monkeypatch.setenv("SSB_WPLANES", "1")
    m = _model(128, 1)
    x = torch.randn(2, 800, 8, generator=torch.Generator().manual_seed(2)).cuda()

    def fwd():
        random.seed(0)
        return m(None, x.clone(), None)[0].detach().clone()

    def fwd_ref():
        monkeypatch.setenv("SSB_WPLANES", "0")
        try:
            return fwd()
        finally:
            monkeypatch.setenv("SSB_WPLANES", "1")

    y0 = fwd()
    assert torch.equal(y0, fwd())                       # no parameter changed: planes reused
    opt = torch.optim.SGD(m.parameters(), lr=0.1)       # in-place torch update
    random.seed(0)
    m(None, x.clone(), None)[0].square().mean().backward()
    opt.step()
    y1 = fwd()
    assert not torch.equal(y0, y1) and torch.equal(y1, fwd_ref())
    bucket = GradientBucket(m)                          # FlatAdamW re-points storage, then writes
    fopt = FlatAdamW(bucket, lr=1e-2)                   # parameters through raw pointers
    y2 = fwd()
    assert torch.equal(y2, y1)
    bucket.zero()
    random.seed(0)
    m(None, x.clone(), None)[0].square().mean().backward()
    fopt.step()
    y3 = fwd()
    assert not torch.equal(y3, y2) and torch.equal(y3, fwd_ref())
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    assert not torch.equal(fwd(), y3)
    m.load_state_dict(sd)                               # copy_ into the same storage
    assert torch.equal(fwd(), y3)

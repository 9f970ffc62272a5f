"""CUDA-graph training step (training.GraphedTrainStep) against the eager step on the GPU:
same weights, same batches, same host RNG for the shift augmentation -> same losses and
parameters; with dropout on, replays must draw fresh masks (device-resident seed offset,
include/ssb.h ssb_set_seed_source) and the shift must follow the host draw."""
import copy
import random

import pytest
import torch
from absl import flags

from oracle import model as om
from silent_speech_b200.read_emg import synthetic_batch

pytestmark = pytest.mark.gpu


def build(D, NL, dropout):
    from silent_speech_b200 import architecture as A
    F = flags.FLAGS
    if not F.is_parsed():
        F(["test"])
    F.model_size, F.num_layers, F.dropout = D, NL, dropout
    m = A.Model(112, 80, 48)
    m.load_state_dict(om.formula_state_dict(D, NL), strict=True)
    return m.cuda().train()


def test_graph_replay_matches_eager(monkeypatch):
    monkeypatch.setenv("SSB_GEMM", "simt")     # bit-stable engine: both arms run identical kernels
    from silent_speech_b200.training import GradientBucket, GraphedTrainStep, train_step
    D, NL, frames, n = 32, 1, 130, 4
    batches = [synthetic_batch(n, frames, seed=20 + i) for i in range(5)]
    m_e = build(D, NL, 0.0)
    m_g = copy.deepcopy(m_e)
    # SGD: Adam's 1/sqrt(v) turns the rounding noise of zero-gradient parameters (split-K atomics
    # reorder fp32 sums run to run) into +-lr steps, which would mask what is compared here
    o_e = torch.optim.SGD(m_e.parameters(), lr=1e-4)
    o_g = torch.optim.SGD(m_g.parameters(), lr=1e-4)
    b_e, b_g = GradientBucket(m_e), GradientBucket(m_g)
    step = GraphedTrainStep(m_g, o_g, "cuda", frames, b_g)
    for it, batch in enumerate(batches):       # step 0 eager, step 1 captures, 2.. replay
        random.seed(7 + it)
        le = train_step(m_e, o_e, batch, "cuda", frames, b_e)
        random.seed(7 + it)
        lg = step(batch)
        # split-K atomics reorder fp32 sums run to run (the eager arm alone moves ~5e-7 between
        # runs); the formula weights amplify that over the updates. A replay bug (stale pointer,
        # frozen seed or shift) shows up at >= 1e-3.
        assert abs(le - lg) <= 5e-5 * abs(le), (it, le, lg)
    assert len(step._graphs) == 1 and step.kernels_per_replay > 0
    for (k, a), (_, b) in zip(m_e.state_dict().items(), m_g.state_dict().items()):
        # includes BN running stats / num_batches_tracked (split-K atomics reorder fp32 sums)
        assert torch.allclose(a.double(), b.double(), rtol=1e-4, atol=1e-6), k


def test_graph_replay_draws_fresh_dropout_and_shift():
    from silent_speech_b200.training import GradientBucket, GraphedTrainStep
    D, NL, frames, n = 64, 1, 130, 4
    batch = synthetic_batch(n, frames, seed=5)
    m = build(D, NL, 0.2)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)          # weights frozen: only the RNG moves
    step = GraphedTrainStep(m, opt, "cuda", frames, GradientBucket(m))
    random.seed(0)
    step(batch), step(batch)                               # eager, capture + first replay

    class FixedShift:                                      # pin the augmentation, vary the masks
        @staticmethod
        def randrange(n):
            return 3
    import silent_speech_b200.training as T
    real = T.random
    T.random = FixedShift
    try:
        losses = [step(batch) for _ in range(4)]
    finally:
        T.random = real
    assert len(set(losses)) == 4, losses                   # fresh dropout masks on every replay
    assert max(losses) - min(losses) < 0.2 * abs(losses[0])

    # dropout off, shift r vs r': the loss must follow the host-drawn shift
    m2 = build(D, NL, 0.0)
    step2 = GraphedTrainStep(m2, torch.optim.SGD(m2.parameters(), lr=0.0), "cuda", frames,
                             GradientBucket(m2))
    step2(batch), step2(batch)
    out = {}
    for r in (0, 5, 0):
        class Shift:
            @staticmethod
            def randrange(n, r=r):
                return r
        T.random = Shift
        try:
            out.setdefault(r, []).append(step2(batch))
        finally:
            T.random = real
    assert out[0][0] == out[0][1] and out[5][0] != out[0][0]


def test_device_shift_matches_reference_semantics():
    from silent_speech_b200.architecture import _shift_rows_device
    x = torch.randn(3, 50, 8, device="cuda")
    for r in range(8):
        want = x.clone()
        if r > 0:
            want[:, :-r, :] = x[:, r:, :]
            want[:, -r:, :] = 0
        got = x.clone()
        _shift_rows_device(got, torch.tensor(r, device="cuda"))
        assert torch.equal(got, want), r

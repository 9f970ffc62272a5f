"""GPU: the on-device batched transduction loss — ragged DTW batches
(csrc/dtw.cu, ssb_dtw_align_ragged), fused cdist + phoneme-cost matrices and the per-frame loss /
sparse-gradient kernel (csrc/dtwloss.cu) — against the C DTW oracle (bit-exact paths) and the CPU
oracle of the reference's dtw_loss (oracle/step.py, transduction_model.py:98-157)."""
import numpy as np
import pytest
import torch

from oracle import dtw as odtw
from oracle import step as ostep
from silent_speech_b200 import align
from silent_speech_b200.read_emg import EMGDataset, synthetic_batch

pytestmark = pytest.mark.gpu


def test_ragged_dtw_batch_is_bit_exact_per_pair():
    """One launch over pairs of different shapes: every path equals the oracle's (= the
    reference's, align.py:16-34) on the SAME matrix.  Shapes cross the 128-row band and the
    16-step chunk boundaries, include tie-heavy integer costs, +inf entries and tiny matrices."""
    rs = np.random.RandomState(0)
    shapes = [(500, 600), (37, 41), (128, 129), (129, 127), (1, 9), (9, 1), (2, 2), (250, 333),
              (64, 16), (15, 17), (600, 500), (300, 257), (131, 640), (5, 300)]   # (T_pred, T_tgt)
    mats = []
    for i, (Tp, Tg) in enumerate(shapes):
        if i % 3 == 1:
            a = rs.randint(0, 4, size=(Tp, Tg)).astype(np.float32)       # heavy ties
        else:
            a = np.abs(rs.randn(Tp, Tg)).astype(np.float32)
        if i % 5 == 2:
            a[rs.rand(Tp, Tg) < 0.05] = np.inf
        mats.append(a)
    paths = align.align_ragged([torch.from_numpy(a).cuda() for a in mats])
    for a, p in zip(mats, paths):
        assert p.dtype == torch.int32 and p.shape == (a.shape[1],)
        assert p.cpu().tolist() == odtw.align_from_distances(a.T), a.shape


def _ragged_batch(n=16, seed=7):
    ds = EMGDataset(num_examples=n, frames=150, frames_jitter=60, seed=seed)
    for i, it in enumerate(ds._items):
        it['silent'] = (i % 2 == 0) or (i % 5 == 0)
    return EMGDataset.collate_raw([ds[i] for i in range(n)])


def _preds_for(batch, seed, chunk=200):
    rows = sum(batch['lengths'])
    B = (rows + chunk - 1) // chunk
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, chunk, 80, generator=g), torch.randn(B, chunk, 48, generator=g) * 2)


def test_ragged_16_utterance_dtw_loss_and_gradients_match_oracle():
    """Every utterance has its own (T_pred, T_target); chunks straddle utterances and the last one
    is zero-padded, exactly like combine_fixed_length / decollate_tensor make them."""
    from silent_speech_b200.losses import dtw_loss
    batch = _ragged_batch()
    assert len(set(batch['lengths'])) > 8 and 5 <= sum(batch['silent']) <= 12
    pred, phon = _preds_for(batch, 1)
    po, qo = pred.clone().requires_grad_(True), phon.clone().requires_grad_(True)
    want = ostep.dtw_loss(po, qo, batch)
    want.backward()
    pg, qg = pred.clone().cuda().requires_grad_(True), phon.clone().cuda().requires_grad_(True)
    got, acc = dtw_loss(pg, qg, batch, phoneme_loss_weight=0.5)
    assert abs(got.item() - want.item()) < 1e-5 * abs(want.item()), (got.item(), want.item())
    got.backward()
    # direct distances vs ATen's mm-based cdist differ by ~1e-6 relative: gradients agree to that,
    # unless two DTW paths tie to within that rounding (none does on this seed)
    for a, b, name in ((pg.grad, po.grad, "pred"), (qg.grad, qo.grad, "phoneme")):
        e = ((a.cpu() - b).norm() / b.norm()).item()
        assert e < 2e-5, (name, e)
    rows = sum(batch['lengths'])
    assert torch.count_nonzero(pg.grad.view(-1, 80)[rows:]) == 0       # padding rows: no gradient


def test_fused_loss_equals_grouped_formulation_and_counts_phonemes(monkeypatch):
    from silent_speech_b200 import losses
    batch = synthetic_batch(8, 120, seed=3)                # uniform shapes: both formulations apply
    pred, phon = _preds_for(batch, 2, chunk=120)
    conf_a, conf_b = np.zeros((48, 48)), np.zeros((48, 48))
    a, acc_a = losses.dtw_loss(pred.cuda(), phon.cuda(), batch, True, conf_a, 0.5)
    monkeypatch.setenv("SSB_LOSS", "grouped")
    b, acc_b = losses.dtw_loss(pred.cuda(), phon.cuda(), batch, True, conf_b, 0.5)
    assert abs(a.item() - b.item()) < 1e-5 * abs(b.item())
    assert acc_a == acc_b and 0.0 <= acc_a <= 1.0
    np.testing.assert_array_equal(conf_a, conf_b)
    assert conf_a.sum() == sum(t.shape[0] for t in batch['audio_features'])


def test_fused_loss_launches_only_library_kernels():
    """`cutlass...sgemm` (ATen cdist) and the per-group torch ops must be gone from
    the loss: what runs is dtw_cost / dtw_fill / dtw_backtrace / dtw_loss_rows plus two
    concatenations and one reduction."""
    from torch.profiler import ProfilerActivity, profile
    from silent_speech_b200.losses import dtw_loss
    batch = _ragged_batch(seed=11)
    for k in ('audio_features', 'phonemes'):               # targets already on the device
        batch[k] = [t.cuda() for t in batch[k]]
    pred, phon = _preds_for(batch, 3)
    pg, qg = pred.cuda().requires_grad_(True), phon.cuda().requires_grad_(True)
    dtw_loss(pg, qg, batch)[0].backward()                  # warm-up: plan tables, lazy init
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loss, _ = dtw_loss(pg, qg, batch)
        loss.backward()
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    assert not any("cutlass" in n or "gemm" in n.lower() or "cdist" in n for n in names), names
    ours = [n for n in names if any(k in n for k in ("dtw_cost", "dtw_fill", "dtw_backtrace",
                                                     "dtw_loss_rows"))]
    assert len(ours) == 4, names
    assert len(names) <= 16, names                          # + cat x2, sum, div, backward scaling


def test_voiced_only_and_silent_only_batches():
    from silent_speech_b200.losses import dtw_loss
    for silent in (False, True):
        ds = EMGDataset(num_examples=5, frames=90, frames_jitter=20, seed=2)
        for it in ds._items:
            it['silent'] = silent
        batch = EMGDataset.collate_raw([ds[i] for i in range(5)])
        pred, phon = _preds_for(batch, 4)
        want = ostep.dtw_loss(pred.clone(), phon.clone(), batch)
        got, _ = dtw_loss(pred.cuda(), phon.cuda(), batch, phoneme_loss_weight=0.5)
        assert abs(got.item() - want.item()) < 1e-5 * abs(want.item())

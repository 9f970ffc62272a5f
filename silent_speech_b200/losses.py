"""On-device, batched restatement of the reference's transduction loss.

`dtw_loss(predictions, phoneme_predictions, example, ...)` keeps the signature and return value
of transduction_model.py:98-157 but removes its per-utterance Python loop and the device->host
sync per silent utterance (transduction_model.py:126): all silent utterances of equal shape
are aligned in ONE ssb_dtw_align_batch call on their cost matrices, which never leave HBM.
(the first "next" row, built because a GPU training step is otherwise
serialised by the loss.)  cdist / log_softmax / cross_entropy stay stock torch ops here.
"""
from collections import defaultdict

import torch
import torch.nn.functional as F

from . import align
from .data_utils import decollate_tensor


def _phoneme_weight(explicit):
    if explicit is not None:
        return explicit
    try:
        from absl import flags
        return flags.FLAGS['phoneme_loss_weight'].value     # defined by transduction_model.py:29
    except Exception:
        return 0.5


def dtw_loss_grouped(predictions, phoneme_predictions, example, phoneme_eval=False,
                     phoneme_confusion=None, phoneme_loss_weight=None):
    """Round-1 formulation, kept for A/B tests (SSB_LOSS=grouped): stock torch cdist /
    log_softmax / gather around one batched DTW launch per group of equally-shaped utterances."""
    device = predictions.device
    w = _phoneme_weight(phoneme_loss_weight)
    preds = decollate_tensor(predictions, example['lengths'])
    phone_preds = decollate_tensor(phoneme_predictions, example['lengths'])
    audio = [t.to(device, non_blocking=True) for t in example['audio_features']]
    phone_tgts = [t.to(device, non_blocking=True) for t in example['phonemes']]

    groups = defaultdict(list)      # (silent, T_pred, T_tgt) -> utterance indices
    for i, (p, y, s) in enumerate(zip(preds, audio, example['silent'])):
        assert p.dim() == 2 and y.dim() == 2
        if not s:
            assert y.size(0) == p.size(0)
        groups[(bool(s), p.size(0), y.size(0))].append(i)

    total = predictions.new_zeros(())
    correct_phones = 0
    total_length = 0
    for (silent, _, Tg), idxs in groups.items():
        P = torch.stack([preds[i] for i in idxs])             # (G, Tp, 80)
        Y = torch.stack([audio[i] for i in idxs])             # (G, Tg, 80)
        PH = torch.stack([phone_preds[i] for i in idxs])      # (G, Tp, 48)
        YP = torch.stack([phone_tgts[i] for i in idxs])       # (G, Tg)
        total_length += Tg * len(idxs)
        if silent:
            lp = F.log_softmax(PH, -1)                                        # :121
            phone_lprobs = torch.gather(lp, 2, YP[:, None, :].expand(-1, lp.size(1), -1))  # :122
            costs = torch.cdist(P, Y) + w * -phone_lprobs                     # :116,124  (G,Tp,Tg)
            alignment = align.align_batch(costs.detach().transpose(1, 2)).long()  # :126  (G,Tg)
            total = total + torch.gather(costs, 1, alignment[:, None, :]).sum()   # :128
            if phoneme_eval:
                pp = lp.argmax(-1)                                             # (G,Tp)
                picked = torch.gather(pp, 1, alignment)
                correct_phones += (picked == YP).sum().item()
                if phoneme_confusion is not None:
                    for a, b in zip(picked.flatten().tolist(), YP.flatten().tolist()):
                        phoneme_confusion[a, b] += 1
        else:
            dists = F.pairwise_distance(Y, P)                                  # :141
            ce = F.cross_entropy(PH.reshape(-1, PH.size(-1)), YP.reshape(-1), reduction='sum')
            total = total + dists.sum() + w * ce                               # :144-145
            if phoneme_eval:
                pp = PH.argmax(-1)
                correct_phones += (pp == YP).sum().item()
                if phoneme_confusion is not None:
                    for a, b in zip(pp.flatten().tolist(), YP.flatten().tolist()):
                        phoneme_confusion[a, b] += 1
    return total / total_length, correct_phones / total_length


# ------------------------------------------------------------------------------------------
# fused, ragged, on-device loss (csrc/dtwloss.cu + the ragged DTW of csrc/dtw.cu)
# ------------------------------------------------------------------------------------------
class _LossPlan:
    """Everything about one batch SIGNATURE (utterance lengths, target lengths, silent flags)
    that the kernels need as device tables.  Built on the host once per signature and cached:
    a CUDA-graph capture of the step then only sees device-resident, address-stable tables."""

    def __init__(self, lengths, tgt_lengths, silent, device):
        import ctypes
        from . import _lib
        n = len(lengths)
        table = (_lib.Utt * max(n, 1))()
        N, M, off, pitch = [], [], [], []
        pr = tr = cur = 0
        for u, (Tp, Tg, s) in enumerate(zip(lengths, tgt_lengths, silent)):
            e = table[u]
            e.pred_row, e.tgt_row, e.Tp, e.Tg, e.silent = pr, tr, int(Tp), int(Tg), int(bool(s))
            e.pair, e.cost_off, e.pitch = -1, 0, 0
            if s:
                q = (int(Tg) + 3) // 4 * 4          # 16 B aligned rows for the DTW's vector loads
                e.pair, e.cost_off, e.pitch = len(N), cur, q
                N.append(int(Tg)), M.append(int(Tp)), off.append(cur), pitch.append(q)
                cur += int(Tp) * q
            else:
                assert Tg == Tp, "voiced utterance: target and prediction lengths differ"
            pr += int(Tp)
            tr += int(Tg)
        self.n_utt, self.rows_used, self.tgt_rows = n, pr, tr
        self.cost_floats = max(cur, 4)
        self.max_Tp, self.max_Tg = max(M, default=0), max(N, default=0)
        self.dtw = align.RaggedPlan(N, M, off, pitch, device)
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self.table_host = raw.pin_memory() if torch.cuda.is_available() else raw
        self.table = self.table_host.to(device, non_blocking=True)
        # per target frame: owning utterance and local index (validation-time phoneme accuracy)
        utt = torch.repeat_interleave(torch.arange(n), torch.tensor([int(t) for t in tgt_lengths]))
        local = torch.cat([torch.arange(int(t)) for t in tgt_lengths]) if n else torch.zeros(0, dtype=torch.int64)
        self.tgt_utt, self.tgt_local = utt.to(device), local.to(device)
        self.utt_pred_row = torch.tensor([table[u].pred_row for u in range(n)], device=device)
        self.utt_pair = torch.tensor([table[u].pair for u in range(n)], device=device)


_plans = {}
_PLAN_LRU = 64


def _plan_for(lengths, tgt_lengths, silent, device):
    key = (tuple(int(x) for x in lengths), tuple(int(x) for x in tgt_lengths),
           tuple(bool(x) for x in silent), str(device))
    plan = _plans.pop(key, None)
    if plan is None:
        plan = _LossPlan(lengths, tgt_lengths, silent, device)
    _plans[key] = plan
    while len(_plans) > _PLAN_LRU:
        _plans.pop(next(iter(_plans)))
    if torch.cuda.is_current_stream_capturing():
        from . import functional as F_
        F_._capture_keepalive.append(plan)      # the graph replays into this plan's tables
    return plan


class _FusedDtwLossFn(torch.autograd.Function):
    """Sum over the batch of the per-utterance losses of transduction_model.py:111-147 and, from
    the same launch, its gradient w.r.t. both prediction tensors (sparse along the DTW path)."""

    @staticmethod
    def forward(ctx, predictions, phoneme_predictions, tgt, tgt_phone, plan, w):
        from . import _lib
        lib = _lib.load()
        F_, NP = predictions.shape[-1], phoneme_predictions.shape[-1]
        pred = predictions.reshape(-1, F_).contiguous()
        phon = phoneme_predictions.reshape(-1, NP).contiguous()
        rows = pred.shape[0]
        assert phon.shape[0] == rows and plan.rows_used <= rows
        dev = pred.device
        st = _lib.current_stream()
        with torch.cuda.device(dev):
            path = None
            if plan.dtw.npairs:
                cost = torch.empty(plan.cost_floats, dtype=torch.float32, device=dev)
                _lib.check(lib.ssb_dtw_cost_batch(pred.data_ptr(), phon.data_ptr(), tgt.data_ptr(),
                                                  tgt_phone.data_ptr(), plan.table.data_ptr(),
                                                  plan.n_utt, plan.max_Tp, plan.max_Tg, F_, NP,
                                                  float(w), cost.data_ptr(), st))
                path = plan.dtw.run(cost)
            row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
            gpred = torch.empty_like(pred)
            gphon = torch.empty_like(phon)
            _lib.check(lib.ssb_dtw_loss_rows(
                pred.data_ptr(), phon.data_ptr(), tgt.data_ptr(), tgt_phone.data_ptr(),
                plan.table.data_ptr(), plan.n_utt, path.data_ptr() if path is not None else None,
                path.shape[1] if path is not None else 0, rows, F_, NP, float(w), 1e-6,
                row_loss.data_ptr(), gpred.data_ptr(), gphon.data_ptr(), st))
        ctx.save_for_backward(gpred, gphon)
        ctx.shapes = (predictions.shape, phoneme_predictions.shape)
        if path is None:
            path = torch.zeros((0, 1), dtype=torch.int32, device=dev)
        ctx.mark_non_differentiable(path)
        return row_loss.sum(), path

    @staticmethod
    def backward(ctx, g, _gpath):
        gpred, gphon = ctx.saved_tensors
        sp, sq = ctx.shapes
        return (gpred * g).view(sp), (gphon * g).view(sq), None, None, None, None


def _fused_ok(predictions, phoneme_predictions):
    import os
    return (os.environ.get("SSB_LOSS", "fused") == "fused" and predictions.is_cuda
            and predictions.dtype == torch.float32 and phoneme_predictions.dtype == torch.float32
            and predictions.shape[-1] <= 128 and phoneme_predictions.shape[-1] <= 64)


def dtw_loss(predictions, phoneme_predictions, example, phoneme_eval=False,
             phoneme_confusion=None, phoneme_loss_weight=None):
    """transduction_model.py:98-157, same signature and return value, for the whole batch in
    four launches (cost matrices, ragged DTW fill, backtrace, per-frame loss + gradient) with
    no host round trip; utterances may all have different lengths."""
    if not _fused_ok(predictions, phoneme_predictions):
        return dtw_loss_grouped(predictions, phoneme_predictions, example, phoneme_eval,
                                phoneme_confusion, phoneme_loss_weight)
    device = predictions.device
    w = _phoneme_weight(phoneme_loss_weight)
    audio = [t.to(device, non_blocking=True) for t in example['audio_features']]
    phone_tgts = [t.to(device, non_blocking=True) for t in example['phonemes']]
    tgt_lengths = [int(t.shape[0]) for t in audio]
    plan = _plan_for(example['lengths'], tgt_lengths, example['silent'], device)
    tgt = torch.cat(audio, 0).to(torch.float32).contiguous()
    tgt_phone = torch.cat(phone_tgts, 0).to(torch.int64).contiguous()
    total, path = _FusedDtwLossFn.apply(predictions, phoneme_predictions, tgt, tgt_phone, plan, w)
    total_length = plan.tgt_rows
    correct_phones = 0
    if phoneme_eval:                                         # :129-137, :147-152 (validation)
        NP = phoneme_predictions.shape[-1]
        pp = phoneme_predictions.reshape(-1, NP).argmax(-1)
        pair = plan.utt_pair[plan.tgt_utt]
        if path.shape[0]:
            aligned = path.long()[pair.clamp(min=0), plan.tgt_local.clamp(max=path.shape[1] - 1)]
        else:
            aligned = plan.tgt_local
        local = torch.where(pair >= 0, aligned, plan.tgt_local)
        picked = pp[plan.utt_pred_row[plan.tgt_utt] + local]
        correct_phones = (picked == tgt_phone).sum().item()
        if phoneme_confusion is not None:
            for a, b in zip(picked.tolist(), tgt_phone.tolist()):
                phoneme_confusion[a, b] += 1
    return total / total_length, correct_phones / total_length


# ------------------------------------------------------------------------------------------
# fused log_softmax + CTC (csrc/ctc.cu)                      recognition_model.py:96-101
# ------------------------------------------------------------------------------------------
def _lengths_on(x, dev):
    """int64 length vector on `dev`; a tensor already there is used as is (no copy: the caller
    may be under CUDA-graph capture, where a pageable H2D copy is illegal)."""
    if torch.is_tensor(x) and x.device == dev and x.dtype == torch.int64:
        return x.contiguous()
    return torch.as_tensor(x, dtype=torch.int64).to(dev).contiguous()


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank, mean):
        from . import _lib
        lib = _lib.load()
        _lib.require_cuda(logits, "logits")
        if logits.dtype != torch.float32 or logits.dim() != 3:
            raise TypeError("ctc_loss: logits must be float32 (N, T, C)")
        logits = logits.contiguous()
        N, T, C = logits.shape
        dev = logits.device
        targets = targets.to(device=dev, dtype=torch.int64).contiguous().view(N, -1)
        il = _lengths_on(input_lengths, dev)
        tl = _lengths_on(target_lengths, dev)
        Lmax = targets.shape[1]
        nll = torch.empty(N, dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(logits) if need_grad else None
        ws = torch.empty(max(16, lib.ssb_ctc_workspace_bytes(N, T, Lmax)), dtype=torch.uint8,
                         device=dev)
        _lib.check(lib.ssb_ctc_loss_fused(
            logits.data_ptr(), N, T, C, targets.data_ptr() if Lmax else None, Lmax, il.data_ptr(),
            tl.data_ptr(), int(blank), int(mean), nll.data_ptr(),
            grad.data_ptr() if need_grad else None, ws.data_ptr(), ws.numel(),
            _lib.current_stream()))
        ctx.save_for_backward(grad)
        ctx.mean = mean
        if mean:     # torch reduction='mean': per-utterance loss / target length, then batch mean
            return (nll / tl.clamp(min=1).to(nll.dtype)).mean()
        return nll

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        if ctx.mean:
            return grad * g, None, None, None, None, None
        return grad * g.view(-1, 1, 1), None, None, None, None, None


def ctc_loss(logits, targets, input_lengths, target_lengths, blank=0, reduction='mean'):
    """Fused replacement for recognition_model.py:96-101:

        pred = F.log_softmax(pred, 2); pred = pad_sequence(decollate_tensor(pred, lengths))
        loss = F.ctc_loss(pred, y, lengths, text_int_lengths, blank=n_chars)

    logits: (N, T, C) batch-first, UN-normalised model outputs padded in time; targets (N, Lmax)
    padded int64; lengths as lists or tensors.  reduction: 'mean' (torch default: per-utterance
    loss / target length, batch mean), 'sum', or 'none'.  One kernel computes the row
    log-softmax, alpha/beta recursions, the loss and d loss / d logits."""
    if reduction not in ('mean', 'sum', 'none'):
        raise ValueError(reduction)
    out = _CtcFn.apply(logits, targets, input_lengths, target_lengths, int(blank), reduction == 'mean')
    return out.sum() if reduction == 'sum' else out


def ctc_loss_from_chunks(pred, example, blank):
    """recognition_model.py:94-101 on the model's chunked output: `pred` (n_chunks, 200, C) as
    returned by Model on combine_fixed_length batches, `example` a collate_raw dict with
    'lengths', 'text_int', 'text_int_lengths'."""
    seqs = decollate_tensor(pred, example['lengths'])
    logits = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True)
    y = torch.nn.utils.rnn.pad_sequence(example['text_int'], batch_first=True)
    # device-resident length vectors when the caller prepared them (training.GraphedTrainStep)
    return ctc_loss(logits, y, example.get('lengths_dev', example['lengths']),
                    example.get('text_int_lengths_dev', example['text_int_lengths']), blank)

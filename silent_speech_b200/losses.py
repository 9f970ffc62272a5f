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


def dtw_loss(predictions, phoneme_predictions, example, phoneme_eval=False,
             phoneme_confusion=None, phoneme_loss_weight=None):
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

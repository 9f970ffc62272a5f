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

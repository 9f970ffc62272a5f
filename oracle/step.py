"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's loss and training step:
  dtw_loss     transduction_model.py:98-157  (per-utterance loop, DTW via oracle/dtw_oracle.c)
  train step   transduction_model.py:196-212 (forward, loss, backward, AdamW weight_decay 1e-7)
on top of oracle/model.py.  Used by tests (parity of silent_speech_b200.losses / training) and
by bench.py's cpu_baseline / `--impl reference` legs as the CPU port of the reference step.
"""
import torch
import torch.nn.functional as F

from . import dtw as odtw
from . import model as omodel


def decollate(tensor, lengths):
    b, s, d = tensor.size()
    flat = tensor.reshape(b * s, d)
    out, idx = [], 0
    for n in lengths:
        out.append(flat[idx:idx + n])
        idx += n
    return out


def dtw_loss(predictions, phoneme_predictions, example, phoneme_loss_weight=0.5):
    """transduction_model.py:98-157 without the phoneme_eval bookkeeping."""
    preds = decollate(predictions, example['lengths'])
    phs = decollate(phoneme_predictions, example['lengths'])
    losses, total_length = [], 0
    for pred, y, pp, yp, silent in zip(preds, example['audio_features'], phs, example['phonemes'],
                                       example['silent']):
        if silent:
            costs = torch.cdist(pred.unsqueeze(0), y.unsqueeze(0)).squeeze(0)          # :116-117
            lp = F.log_softmax(pp, -1)                                                  # :121
            costs = costs + phoneme_loss_weight * -lp[:, yp]                            # :122-124
            alignment = odtw.align_from_distances(costs.T.detach().numpy())             # :126
            loss = costs[alignment, range(len(alignment))].sum()                        # :128
        else:
            assert y.size(0) == pred.size(0)
            loss = F.pairwise_distance(y, pred).sum() + \
                phoneme_loss_weight * F.cross_entropy(pp, yp, reduction='sum')          # :141-145
        losses.append(loss)
        total_length += y.size(0)
    return sum(losses) / total_length


def make_params(sd):
    """Trainable leaf tensors for a reference-shaped state_dict (BN running stats stay buffers)."""
    out = {}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            out[k] = v.clone().requires_grad_(True)
        else:
            out[k] = v.clone()
    return out


def make_optimizer(params, lr=1e-3, weight_decay=1e-7):
    """transduction_model.py:178 — torch.optim.AdamW(model.parameters(), weight_decay=FLAGS.l2)."""
    leaves = [v for v in params.values() if v.requires_grad]
    return torch.optim.AdamW(leaves, lr=lr, weight_decay=weight_decay)


def train_step(params, optim, batch, seq_len_frames, dropout_p=0.0):
    """transduction_model.py:197-210 on the CPU oracle model."""
    optim.zero_grad()
    raw = torch.cat(list(batch['raw_emg']), 0)
    L = seq_len_frames * 8
    if raw.size(0) % L:
        raw = torch.cat([raw, raw.new_zeros(L - raw.size(0) % L, raw.size(1))], 0)
    X_raw = raw.view(-1, L, raw.size(1))
    pred, phon = omodel.model_forward(params, X_raw, training=True, dropout_p=dropout_p)
    loss = dtw_loss(pred, phon, batch)
    loss.backward()
    optim.step()
    return loss.item()

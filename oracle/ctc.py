"""ORACLE — test infrastructure, not product code.

CPU restatement (numpy, float64) of the recognition loss of recognition_model.py:96-101:

    pred = F.log_softmax(pred, 2)
    pred = pad_sequence(decollate_tensor(pred, lengths))            # (T, N, C)
    loss = F.ctc_loss(pred, y, lengths, text_int_lengths, blank=n_chars)   # reduction='mean'

The arithmetic lives in a third-party dependency (PyTorch, pinned 2.0 in environment.yml:10;
2.11 in this image): this file restates the published CTC forward-backward algorithm (Graves et
al. 2006) in log space and is pinned against `F.ctc_loss` executed in the build container
(tests/golden/ctc_golden.npz, made by tests/golden/make_golden_ctc.py).
"""
import numpy as np


def _lse(*xs):
    m = max(xs)
    if m == -np.inf:
        return -np.inf
    return m + np.log(sum(np.exp(x - m) for x in xs))


def ctc_nll_and_grad(logits, targets, input_lengths, target_lengths, blank):
    """logits (N, T, C) un-normalised; targets (N, Lmax) ints; returns
    nll (N,) and d nll[n] / d logits (N, T, C) (zero on padding frames)."""
    logits = np.asarray(logits, dtype=np.float64)
    N, T, C = logits.shape
    nll = np.zeros(N)
    grad = np.zeros_like(logits)
    for n in range(N):
        Tn, L = int(input_lengths[n]), int(target_lengths[n])
        ext = [blank]
        for c in np.asarray(targets[n][:L]).tolist():
            ext += [int(c), blank]
        S = len(ext)
        x = logits[n, :Tn]
        lp = x - (x.max(1, keepdims=True) + np.log(np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)))
        al = np.full((Tn, S), -np.inf)
        al[0, 0] = lp[0, ext[0]]
        if S > 1:
            al[0, 1] = lp[0, ext[1]]
        for t in range(1, Tn):
            for s in range(S):
                c = [al[t - 1, s]]
                if s >= 1:
                    c.append(al[t - 1, s - 1])
                if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                    c.append(al[t - 1, s - 2])
                al[t, s] = _lse(*c) + lp[t, ext[s]]
        ll = _lse(al[Tn - 1, S - 1], al[Tn - 1, S - 2] if S > 1 else -np.inf)
        nll[n] = -ll
        if ll == -np.inf:
            continue                      # infeasible: inf loss, zero gradient (kernel convention)
        be = np.full((Tn, S), -np.inf)
        be[Tn - 1, S - 1] = lp[Tn - 1, ext[S - 1]]
        if S > 1:
            be[Tn - 1, S - 2] = lp[Tn - 1, ext[S - 2]]
        for t in range(Tn - 2, -1, -1):
            for s in range(S):
                c = [be[t + 1, s]]
                if s + 1 < S:
                    c.append(be[t + 1, s + 1])
                if s + 2 < S and ext[s + 2] != blank and ext[s + 2] != ext[s]:
                    c.append(be[t + 1, s + 2])
                be[t, s] = _lse(*c) + lp[t, ext[s]]
        occ = np.zeros((Tn, C))
        for s in range(S):
            occ[:, ext[s]] += np.exp(al[:, s] + be[:, s] - lp[:, ext[s]] - ll)
        grad[n, :Tn] = np.exp(lp) - occ
    return nll, grad


def ctc_loss_mean(logits, targets, input_lengths, target_lengths, blank):
    """torch reduction='mean': per-utterance nll / max(target length, 1), then the batch mean;
    returns (loss, d loss / d logits)."""
    nll, grad = ctc_nll_and_grad(logits, targets, input_lengths, target_lengths, blank)
    w = 1.0 / (np.maximum(np.asarray(target_lengths, dtype=np.float64), 1.0) * len(nll))
    return float((nll * w).sum()), grad * w[:, None, None]

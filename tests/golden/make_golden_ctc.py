"""Golden vectors for the recognition loss: the reference's own formulation
(recognition_model.py:96-101) executed with PyTorch in the build container, in float64.

    python tests/golden/make_golden_ctc.py      # writes tests/golden/ctc_golden.npz
"""
import os

import numpy as np
import torch
import torch.nn.functional as F


def case(N, T, C, Lmax, seed, small_alphabet=False):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(N, T, C, generator=g) * 2).to(torch.float32)
    hi = 3 if small_alphabet else C - 1
    targets = torch.randint(0, hi, (N, Lmax), generator=g)
    il = torch.randint(max(1, T // 2), T + 1, (N,), generator=g)
    il[0] = T
    tl = torch.minimum(torch.randint(0, Lmax + 1, (N,), generator=g), il // 2)
    return logits, targets, il, tl


def main():
    out = {}
    cases = [(3, 40, 6, 8, 1, False), (2, 120, 38, 30, 2, False), (4, 33, 5, 12, 3, True)]
    for i, (N, T, C, Lmax, seed, small) in enumerate(cases):
        logits, targets, il, tl = case(N, T, C, Lmax, seed, small)
        x = logits.double().requires_grad_(True)
        lp = F.log_softmax(x, 2).transpose(0, 1)                       # (T, N, C) as the reference
        nll = F.ctc_loss(lp, targets, il, tl, blank=C - 1, reduction='none')
        loss = F.ctc_loss(lp, targets, il, tl, blank=C - 1)            # reduction='mean' (default)
        (gm,) = torch.autograd.grad(loss, x, retain_graph=True)
        (gn,) = torch.autograd.grad(nll.sum(), x)
        for k, v in dict(logits=logits, targets=targets, il=il, tl=tl, nll=nll.detach(),
                         loss=loss.detach(), grad_mean=gm, grad_nll=gn).items():
            out[f"c{i}_{k}"] = v.numpy()
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("c1_")})


if __name__ == "__main__":
    main()

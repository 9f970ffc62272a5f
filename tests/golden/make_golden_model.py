"""Generate tests/golden/model_golden.npz by running the reference's Model
(architecture.py + transformer.py, unmodified) in this container on formula-defined weights
and inputs, forward and backward, eval and train mode (dropout 0).  Also prints the deviation
of oracle/model.py from the executed reference.   Run: python tests/golden/make_golden_model.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _reference_import import fix_transformer_shim, import_reference  # noqa: E402

OUT = os.path.join(HERE, "model_golden.npz")

# (name, model_size, num_layers, B, L, python-random seed used for the train-mode shift, input id)
# Input ids were picked so that no ReLU pre-activation of the fp64 oracle lies within 1e-5 of
# zero (short 8.6e-5, band 4.2e-5, odd 1.2e-5): a gradient comparison across implementations
# is only meaningful when no mask sits on the kink (one flipped mask moved a conv-block
# gradient of the original "odd" input, margin 3e-6, by 10 %).
CASES = [("short", 32, 2, 3, 200, 1, 0), ("band", 32, 1, 2, 1000, 2, 1),
         ("odd", 32, 1, 2, 1003, 3, 11)]


def make_input(B, L, case_idx):
    i = torch.arange(B * L * 8, dtype=torch.float64)
    x = torch.sin(i * 0.731 + case_idx) + 0.5 * torch.sin(i * 0.0173 + 2.0 * case_idx)
    return x.to(torch.float32).reshape(B, L, 8).clone()


def probe_weights(shape, k):
    n = int(np.prod(shape))
    i = torch.arange(n, dtype=torch.float64)
    return torch.cos(i * 0.91 + k).to(torch.float32).reshape(shape)


def scalar_loss(pred, aux):
    return (pred * probe_weights(pred.shape, 1)).sum() / pred.numel() * 100.0 + \
           (aux * probe_weights(aux.shape, 2)).sum() / aux.numel() * 100.0


def grad_fingerprint(named_grads):
    """per-parameter: L2 norm, sum, and 4 strided samples"""
    out = {}
    for name, g in named_grads.items():
        g = g.detach().double().flatten()
        idx = torch.linspace(0, g.numel() - 1, 4).long()
        out[name] = np.array([g.norm().item(), g.sum().item()] + g[idx].tolist())
    return out


def main():
    from absl import flags
    arch, = import_reference("architecture")
    FLAGS = flags.FLAGS
    FLAGS(["make_golden_model"])
    from oracle import model as om

    store = {"meta": np.array([",".join(map(str, c)) for c in CASES])}
    for ci, (name, D, NL, B, L, pyseed, inp) in enumerate(CASES):
        FLAGS.model_size, FLAGS.num_layers, FLAGS.dropout = D, NL, 0.0
        ref = fix_transformer_shim(arch.Model(112, 80, 48))
        sd = om.formula_state_dict(D, NL)
        ref.load_state_dict(sd, strict=True)

        # ---- eval forward
        ref.eval()
        x = make_input(B, L, inp)
        with torch.no_grad():
            pred, aux = ref(None, x.clone(), None)
        store[f"{name}_eval_pred"] = pred.numpy()
        store[f"{name}_eval_aux"] = aux.numpy()

        # ---- train forward/backward (batch-stat BN, random shift, dropout 0)
        ref.train()
        random.seed(pyseed)
        xt = x.clone()
        pred, aux = ref(None, xt, None)
        loss = scalar_loss(pred, aux)
        loss.backward()
        store[f"{name}_train_pred"] = pred.detach().numpy()
        store[f"{name}_train_aux"] = aux.detach().numpy()
        store[f"{name}_train_loss"] = np.array(loss.item())
        store[f"{name}_train_x_after"] = xt[:, -9:, :].numpy()   # shows the in-place shift
        grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        nograd = [k for k, p in ref.named_parameters() if p.grad is None]
        assert all(k.endswith("relative_positional.embeddings") for k in nograd), nograd
        for k, v in grad_fingerprint(grads).items():
            store[f"{name}_grad::{k}"] = v
        rsd = ref.state_dict()
        for k in rsd:
            if "running_" in k or "num_batches" in k:
                store[f"{name}_buf::{k}"] = rsd[k].numpy()

        # ---- oracle restatement vs executed reference (printed, not stored)
        sd2 = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                   else v.clone()) for k, v in om.formula_state_dict(D, NL).items()}
        with torch.no_grad():
            op, oa = om.model_forward(sd2, x.clone(), training=False)
        e_eval = (op - torch.from_numpy(store[f"{name}_eval_pred"])).abs().max().item()
        random.seed(pyseed)
        op, oa = om.model_forward(sd2, x.clone(), training=True, dropout_p=0.0)
        ol = scalar_loss(op, oa)
        ol.backward()
        e_train = (op.detach() - pred.detach()).abs().max().item()
        worst = 0.0
        for k, g in grads.items():
            og = sd2[k].grad
            worst = max(worst, ((og - g).norm() / (g.norm() + 1e-30)).item())
        assert sd2[f"transformer.layers.0.self_attn.relative_positional.embeddings"].grad is None
        e_buf = max((sd2[k] - rsd[k]).abs().max().item() for k in rsd if "running_" in k)
        print(f"{name}: T={pred.shape[1]} oracle-vs-reference max|eval|={e_eval:.2e} "
              f"max|train|={e_train:.2e} loss {ol.item():.6f}/{loss.item():.6f} "
              f"worst grad rel-L2={worst:.2e} running-stat max|d|={e_buf:.2e}")

    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()

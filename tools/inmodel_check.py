"""In-model replay check: wrap functional._bn_bwd / conv backward pieces and verify each call
against fp64 formulas evaluated on the very tensors the kernels received."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
os.environ["SSB_GEMM"] = "simt"
from make_golden_model import make_input, scalar_loss  # noqa: E402
from oracle import model as om  # noqa: E402
from silent_speech_b200 import functional as SF  # noqa: E402

D, NL, B, L, CI = (int(v) for v in sys.argv[1:6])
orig_bwd = SF._bn_bwd
calls = []


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def checked_bn_bwd(dy2, mask_src, x2, stats, gamma, training):
    dx, dg, db = orig_bwd(dy2, mask_src, x2, stats, gamma, training)
    torch.cuda.synchronize()
    dz = dy2.double() * ((mask_src.double() > 0) if mask_src is not None else 1.0)
    mean, rstd = stats[0].double(), stats[1].double()
    xh = (x2.double() - mean) * rstd
    n = x2.shape[0]
    rdb, rdg = dz.sum(0), (dz * xh).sum(0)
    rdx = gamma.double() * rstd * (dz - rdb / n - xh * rdg / n)
    tm, tv = x2.double().mean(0), x2.double().var(0, unbiased=False)
    print(f"bn_bwd rows={n} C={x2.shape[1]}: dx {rel(dx, rdx):.1e} dgamma {rel(dg, rdg):.1e} dbeta {rel(db, rdb):.1e} "
          f"| saved mean vs true {rel(mean, tm):.1e} rstd vs true {rel(rstd, 1/torch.sqrt(tv+1e-5)):.1e} "
          f"| mask nonzero frac {(mask_src > 0).float().mean().item() if mask_src is not None else -1:.3f}")
    return dx, dg, db


SF._bn_bwd = checked_bn_bwd
from absl import flags  # noqa: E402
from silent_speech_b200 import architecture as A  # noqa: E402
F = flags.FLAGS
F(["x"])
F.model_size, F.num_layers, F.dropout = D, NL, 0.0
m = A.Model(112, 80, 48)
m.load_state_dict(om.formula_state_dict(D, NL))
m = m.cuda().train()
random.seed(3)
pred, aux = m(None, make_input(B, L, CI).cuda(), None)
scalar_loss(pred.cpu(), aux.cpu()).backward()

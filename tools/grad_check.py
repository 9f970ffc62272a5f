"""Diagnostic: forward / per-parameter gradient error of the CUDA Model vs the fp64 CPU oracle.
usage: python tools/grad_check.py D NL B L [engine]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_model import make_input, scalar_loss  # noqa: E402
from oracle import model as om  # noqa: E402

D, NL, B, L = (int(v) for v in sys.argv[1:5])
if len(sys.argv) > 5:
    os.environ["SSB_GEMM"] = sys.argv[5]
from absl import flags  # noqa: E402
from silent_speech_b200 import architecture as A  # noqa: E402
F = flags.FLAGS
F(["x"])
F.model_size, F.num_layers, F.dropout = D, NL, 0.0
m = A.Model(112, 80, 48)
sd0 = om.formula_state_dict(D, NL)
m.load_state_dict(sd0)
m = m.cuda().train()
x = make_input(B, L, 5)
random.seed(3)
pred, aux = m(None, x.clone().cuda(), None)
scalar_loss(pred.cpu(), aux.cpu()).backward()
sd = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
          else (v.double() if v.is_floating_point() else v.clone())) for k, v in sd0.items()}
random.seed(3)
op, oa = om.model_forward(sd, x.double(), training=True, dropout_p=0.0)
p64 = torch.cos(torch.arange(op.numel(), dtype=torch.float64) * 0.91 + 1).reshape(op.shape)
a64 = torch.cos(torch.arange(oa.numel(), dtype=torch.float64) * 0.91 + 2).reshape(oa.shape)
((op * p64).sum() / op.numel() * 100 + (oa * a64).sum() / oa.numel() * 100).backward()
print(f"D={D} NL={NL} B={B} L={L} engine={os.environ.get('SSB_GEMM','tc')}: fwd rel-L2 "
      f"{((pred.detach().cpu().double() - op.detach()).norm() / op.detach().norm()).item():.2e}")
errs = []
for k, p in m.named_parameters():
    if sd[k].grad is None or k.endswith(("conv1.bias", "conv2.bias", "residual_path.bias")) and k.startswith("conv"):
        continue
    g, r = p.grad.cpu().double(), sd[k].grad
    errs.append(((g - r).norm() / (r.norm() + 1e-30)).item())
    errs[-1] = (errs[-1], k)
errs.sort(reverse=True)
for e, k in errs[:6]:
    print(f"   {e:.2e}  {k}")

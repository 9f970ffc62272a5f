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
CI = int(sys.argv[6]) if len(sys.argv) > 6 else 5
x = make_input(B, L, CI)
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
for e, k in errs[:int(os.environ.get("TOPN", "6"))]:
    print(f"   {e:.2e}  {k}")

# ---- optional: compare the gradient arriving at each ResBlock output with the oracle's ----
if os.environ.get("BLOCKS"):
    import torch.nn.functional as Fnn
    from silent_speech_b200.architecture import ResBlock
    outs = []
    orig = ResBlock.forward_cl

    def wrapped(self, x):
        y = orig(self, x)
        y.retain_grad()
        outs.append(y)
        return y
    ResBlock.forward_cl = wrapped
    m.zero_grad()
    random.seed(3)
    pred, aux = m(None, make_input(B, L, CI).cuda(), None)
    scalar_loss(pred.cpu(), aux.cpu()).backward()
    ocap = []
    orb = om.res_block

    def orb_cap(*a, **k):
        y = orb(*a, **k)
        y.retain_grad()
        ocap.append(y)
        return y
    om.res_block = orb_cap
    sd2 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
               else (v.double() if v.is_floating_point() else v.clone())) for k, v in sd0.items()}
    random.seed(3)
    op, oa = om.model_forward(sd2, make_input(B, L, CI).double(), training=True, dropout_p=0.0)
    ((op * p64).sum() / op.numel() * 100 + (oa * a64).sum() / oa.numel() * 100).backward()
    errs2 = []
    for k, p_ in m.named_parameters():
        if sd[k].grad is None or (k.startswith("conv") and k.endswith(("conv1.bias", "conv2.bias", "residual_path.bias"))):
            continue
        g_, r_ = p_.grad.cpu().double(), sd[k].grad
        errs2.append((((g_ - r_).norm() / (r_.norm() + 1e-30)).item(), k))
    errs2.sort(reverse=True)
    print("SECOND run worst params:", [(f"{e:.1e}", k) for e, k in errs2[:4]])
    for i in range(3):
        a_, b_ = outs[i].grad.cpu().double(), ocap[i].grad.transpose(1, 2)
        d = (a_ - b_)
        bad = d.abs().amax(dim=(0, 2))
        print(f"block {i} output-grad rel err {(d.norm() / b_.norm()).item():.2e}; "
              f"fwd {((outs[i].detach().cpu().double() - ocap[i].detach().transpose(1, 2)).norm() / ocap[i].detach().norm()).item():.2e}; "
              f"worst time idx {bad.argmax().item()} of {bad.numel()}; err by time (last 4): {[f'{v:.1e}' for v in bad[-4:].tolist()]} "
              f"first 2: {[f'{v:.1e}' for v in bad[:2].tolist()]}")

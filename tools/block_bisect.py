"""Bisect a data-dependent backward discrepancy in the conv blocks: take the fp64 oracle's
intermediate activations / gradients for one case and replay every op of every ResBlock on the
GPU kernels with exactly those tensors.   usage: python tools/block_bisect.py D NL B L ci"""
import os
import random
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
os.environ["SSB_GEMM"] = "simt"
from make_golden_model import make_input  # noqa: E402
from oracle import model as om  # noqa: E402
from silent_speech_b200 import functional as SF  # noqa: E402

D, NL, B, L, CI = (int(v) for v in sys.argv[1:6])
dt = torch.float64
sd = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
          else (v.to(dt) if v.is_floating_point() else v.clone()))
      for k, v in om.formula_state_dict(D, NL).items()}
cap = {}


def res_block_cap(x, sd, prefix, stride, training, update_running=True):
    def keep(name, t):
        if t.requires_grad:
            t.retain_grad()
        cap[prefix + "." + name] = t
        return t
    x = keep("x", x * 1.0)
    c1 = keep("c1", F.conv1d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], stride=stride, padding=1))
    h1 = keep("h1", F.relu(om._bn(c1, sd, prefix + ".bn1", training, update_running)))
    c2 = keep("c2", F.conv1d(h1, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=1))
    b2 = om._bn(c2, sd, prefix + ".bn2", training, update_running)
    cr = keep("cr", F.conv1d(x, sd[prefix + ".residual_path.weight"], sd[prefix + ".residual_path.bias"], stride=stride))
    br = om._bn(cr, sd, prefix + ".res_norm", training, update_running)
    return keep("y", F.relu(b2 + br))


om.res_block = res_block_cap
random.seed(3)
x = make_input(B, L, CI).to(dt)
p, a = om.model_forward(sd, x, training=True, dropout_p=0.0)
p64 = torch.cos(torch.arange(p.numel(), dtype=dt) * 0.91 + 1).reshape(p.shape)
a64 = torch.cos(torch.arange(a.numel(), dtype=dt) * 0.91 + 2).reshape(a.shape)
((p * p64).sum() / p.numel() * 100 + (a * a64).sum() / a.numel() * 100).backward()


def cl(t):   # (B, C, L) fp64 cpu -> (B, L, C) fp32 cuda contiguous
    return t.detach().transpose(1, 2).contiguous().float().cuda()


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


for i in range(3):
    pre = f"conv_blocks.{i}"
    g = lambda n: cap[pre + "." + n]
    f32 = lambda k: sd[pre + "." + k].detach().float().cuda()
    stride = 2
    # --- final bn2 + res_norm + relu
    c2, cr, y = cl(g("c2")).requires_grad_(True), cl(g("cr")).requires_grad_(True), g("y")
    params = [f32(k).requires_grad_(True) for k in ("bn2.weight", "bn2.bias", "res_norm.weight", "res_norm.bias")]
    rm = [torch.zeros_like(params[0]) for _ in range(4)]
    yo = SF.bn_act(c2, params[0], params[1], rm[0], rm[1] + 1, True, True, cr, params[2], params[3], rm[2], rm[3] + 1)
    yo.backward(cl(g("y").grad))
    print(f"block {i}: bn_act2 y {rel(yo.transpose(1,2), y):.1e} dc2 {rel(c2.grad.transpose(1,2), g('c2').grad):.1e} "
          f"dcr {rel(cr.grad.transpose(1,2), g('cr').grad):.1e} dgamma2 {rel(params[0].grad, sd[pre+'.bn2.weight'].grad):.1e} "
          f"dbeta2 {rel(params[1].grad, sd[pre+'.bn2.bias'].grad):.1e} dgr {rel(params[2].grad, sd[pre+'.res_norm.weight'].grad):.1e}")
    # --- conv2 (k3 s1)
    h1 = cl(g("h1")).requires_grad_(True)
    w = sd[pre + ".conv2.weight"].detach().float().cuda().requires_grad_(True)
    bb = f32("conv2.bias").requires_grad_(True)
    Wg = w.permute(2, 1, 0).reshape(-1, w.shape[0]).contiguous()
    o = SF.conv1d_cl(h1, Wg, bb, 3, 1)
    o.backward(cl(g("c2").grad))
    print(f"         conv2 y {rel(o.transpose(1,2), g('c2')):.1e} dh1 {rel(h1.grad.transpose(1,2), g('h1').grad):.1e} "
          f"dW {rel(w.grad, sd[pre+'.conv2.weight'].grad):.1e}")
    # --- bn1 + relu
    c1 = cl(g("c1")).requires_grad_(True)
    ps = [f32(k).requires_grad_(True) for k in ("bn1.weight", "bn1.bias")]
    ho = SF.bn_act(c1, ps[0], ps[1], rm[0] * 0, rm[1] * 0 + 1, True, True)
    ho.backward(cl(g("h1").grad))
    print(f"         bn_act1 y {rel(ho.transpose(1,2), g('h1')):.1e} dc1 {rel(c1.grad.transpose(1,2), g('c1').grad):.1e} "
          f"dgamma {rel(ps[0].grad, sd[pre+'.bn1.weight'].grad):.1e} dbeta {rel(ps[1].grad, sd[pre+'.bn1.bias'].grad):.1e}")
    # --- conv1 (k3 s2) and residual (k1 s2)
    for nm, k, gout in (("conv1", 3, "c1"), ("residual_path", 1, "cr")):
        xin = cl(g("x")).requires_grad_(i > 0)
        w = sd[pre + f".{nm}.weight"].detach().float().cuda().requires_grad_(True)
        bb = f32(f"{nm}.bias").requires_grad_(True)
        Wg = w.permute(2, 1, 0).reshape(-1, w.shape[0]).contiguous()
        o = SF.conv1d_cl(xin, Wg, bb, k, stride)
        o.backward(cl(g(gout).grad))
        print(f"         {nm} y {rel(o.transpose(1,2), g(gout)):.1e} dW {rel(w.grad, sd[pre+f'.{nm}.weight'].grad):.1e}")

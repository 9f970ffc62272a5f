"""Bring-up check of the fused attention kernels: forward and backward separately, each
synchronised, against the dense torch restatement (tests/test_ops_gpu.py::dense_attention_ref).

    python tools/attn_fused_check.py [B T H dh W [p]]
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import functional as SF  # noqa: E402


def dense_ref(qkv, E, B, T, H, dh, W):
    D = H * dh
    q, k, v = (qkv[:, i * D:(i + 1) * D].view(B, T, H, dh).permute(0, 2, 1, 3) for i in range(3))
    logits = q @ k.transpose(-1, -2) / math.sqrt(dh)
    R = torch.einsum('bhqa,hra->bhqr', q, E[:, :2 * W + 1])
    ar = torch.arange(T, device=qkv.device)
    relidx = ar[None, :] - ar[:, None] + W
    inb = (relidx >= 0) & (relidx <= 2 * W)
    pos = torch.gather(R, 3, relidx.clamp(0, 2 * W)[None, None].expand(B, H, T, T))
    logits = torch.where(inb[None, None], logits + pos, torch.full_like(logits, float("-inf")))
    o = torch.softmax(logits, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * T, D)


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def main():
    a = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else [1, 130, 2, 32, 99]
    B, T, H, dh, W = a
    D = H * dh
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * T, 3 * D, generator=g).cuda().requires_grad_(True)
    RW = (2 * W + 1 + 3) // 4 * 4
    E = torch.zeros(H, RW, dh).cuda()
    E[:, :2 * W + 1] = (torch.randn(H, 2 * W + 1, dh, generator=g) * dh ** -0.5).cuda()
    assert SF._fused_attn_ok(B, T, H, dh, W)
    o = SF.band_attention(qkv, E, B, T, H, dh, W, 0.0, 0, 0)
    torch.cuda.synchronize()
    print("forward ran")
    qr = qkv.detach().clone().requires_grad_(True)
    orf = dense_ref(qr, E, B, T, H, dh, W)
    print("O rel-L2", rel(o.detach(), orf.detach()))
    go = torch.randn(B * T, D, generator=g).cuda()
    o.backward(go)
    torch.cuda.synchronize()
    print("backward ran")
    orf.backward(go)
    for n, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        print(n, "rel-L2", rel(qkv.grad[:, sl], qr.grad[:, sl]))


if __name__ == "__main__":
    main()

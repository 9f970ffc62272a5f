"""Bring-up checks of the tcgen05 GEMM (csrc/gemm_tc.cu) against fp64 torch references.
usage: python tools/tc_gemm_check.py {split|kmajor|wgrad|conv1|conv2|perf}   (run each under `timeout`)"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import functional as SF  # noqa: E402

dev = "cuda"


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(dev)


def t_split():
    x = rnd(1000, 768, seed=1)
    p = SF.split_planes(x)
    rec = p[0].float() + p[1].float()
    print("split: rel err", rel(rec, x), "max", (rec - x).abs().max().item())


def t_kmajor():
    for (M, N, K) in [(128, 256, 64), (256, 256, 128), (1000, 768, 768), (333, 80, 768),
                      (16000, 3072, 768), (16000, 768, 3072)]:
        x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
        xp, wp = SF.split_planes(x), SF.split_planes(w)
        y = torch.full((M, N), float("nan"), device=dev)
        SF.gemm_tc_kmajor(SF.tc_operand_plain(xp, M, K), wp, N, K,
                          SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b, relu=1))
        torch.cuda.synchronize()
        ref = torch.relu(x.double() @ w.double().t() + b.double())
        print(f"kmajor {M}x{N}x{K}: rel err {rel(y, ref):.3e} finite={torch.isfinite(y).all().item()}")


def t_wgrad():
    for (M, K, N) in [(64, 128, 256), (640, 128, 256), (1000, 768, 768), (16000, 768, 3072),
                      (16000, 3072, 768), (4000, 768, 2304), (1000, 768, 80)]:
        x, g = rnd(M, K, seed=1), rnd(M, N, seed=2)
        xp, gp = SF.split_planes(x), SF.split_planes(g)
        dW = torch.full((K, N), float("nan"), device=dev)
        SF.gemm_tc_wgrad(SF.tc_operand_plain(xp, M, K), gp, N, K, dW)
        torch.cuda.synchronize()
        ref = x.double().t() @ g.double()
        print(f"wgrad M={M} K={K} N={N}: rel err {rel(dW, ref):.3e} finite={torch.isfinite(dW).all().item()}")


def t_conv(stride):
    for (B, L, C, Co) in [(2, 256, 64, 256), (3, 250, 128, 768), (4, 1000, 768, 768), (2, 37, 64, 64)]:
        x = rnd(B, L, C, seed=1)
        w = rnd(Co, C, 3, seed=2, scale=(3 * C) ** -0.5)
        bias = rnd(Co, seed=3)
        Lout = (L - 1) // stride + 1
        xp = SF.split_planes(x)
        wp = SF.split_planes(w.permute(0, 2, 1).reshape(Co, 3 * C).contiguous())   # [co][(tap, ci)]
        y = torch.full((B, Lout, Co), float("nan"), device=dev)
        op = SF.tc_operand_conv(xp, B, L, C, Lout, stride, 1, -1)
        SF.gemm_tc_kmajor(op, wp, Co, 3 * C, SF._epi(SF.Scatter(y.data_ptr(), Lout * Co, Lout, Co, 1, 0), bias=bias))
        torch.cuda.synchronize()
        ref = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), stride=stride, padding=1).transpose(1, 2)
        print(f"conv s={stride} B={B} L={L} C={C} Co={Co}: rel err {rel(y, ref):.3e} finite={torch.isfinite(y).all().item()}")
        # weight gradient through the MN-major path: dW[(tap,ci), co] = sum x_im2col * g
        g = rnd(B, Lout, Co, seed=4)
        gp = SF.split_planes(g)
        if C % 128 == 0 or True:
            try:
                dW = torch.full((3 * C, Co), float("nan"), device=dev)
                SF.gemm_tc_wgrad(op, gp, Co, 3 * C, dW)
                torch.cuda.synchronize()
                xr = x.double().requires_grad_(False)
                wr = w.double().clone().requires_grad_(True)
                F.conv1d(xr.transpose(1, 2), wr, None, stride=stride, padding=1).transpose(1, 2).backward(g.double())
                refw = wr.grad.permute(2, 1, 0).reshape(3 * C, Co)
                print(f"   conv wgrad: rel err {rel(dW, refw):.3e}")
            except Exception as e:
                print("   conv wgrad skipped:", str(e)[:100])


def t_perf():
    M, N, K = 16000, 3072, 768
    x, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    xp, wp = SF.split_planes(x), SF.split_planes(w)
    y = torch.empty(M, N, device=dev)
    op = SF.tc_operand_plain(xp, M, K)
    ep = SF._epi(SF._scatter_plain(y.data_ptr(), M, N))
    for _ in range(3):
        SF.gemm_tc_kmajor(op, wp, N, K, ep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        SF.gemm_tc_kmajor(op, wp, N, K, ep)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"perf kmajor {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
    g = rnd(M, N, seed=3)
    gp = SF.split_planes(g)
    dW = torch.empty(K, N, device=dev)
    for _ in range(3):
        SF.gemm_tc_wgrad(op, gp, N, K, dW)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        SF.gemm_tc_wgrad(op, gp, N, K, dW)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"perf wgrad  {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
    e0.record()
    for _ in range(10):
        SF.split_planes(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"perf split {M}x{K}: {e0.elapsed_time(e1)/10*1e3:.1f} us")


if __name__ == "__main__":
    what = sys.argv[1]
    {"split": t_split, "kmajor": t_kmajor, "wgrad": t_wgrad, "conv1": lambda: t_conv(1),
     "conv2": lambda: t_conv(2), "perf": t_perf}[what]()

"""Bring-up checks of the tcgen05 GEMM (csrc/gemm_tc.cu) against fp64 torch references.
usage: python tools/tc_gemm_check.py {split|kmajor|wgrad|conv1|conv2|perf|shapes|ffn|convgrad}   (run each under `timeout`)"""
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




def t_shapes():
    """in-step GEMM shapes of cfg-1 (16000 tokens), L2-warm and L2-flushed, epilogue variants"""
    flush = torch.empty(64 * 1024 * 1024, device=dev)   # 256 MB > 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, cold):
        for _ in range(2):
            fn()
        tot = 0.0
        for _ in range(6):
            if cold:
                flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / 6

    M = 16000
    for (name, N, K, kw) in [("qkv", 2304, 768, {}), ("wo", 768, 768, {}),
                             ("ffn1 relu", 3072, 768, dict(relu=1)),
                             ("ffn1 relu+drop", 3072, 768, dict(relu=1, drop_p=0.2, seed=5, site=3)),
                             ("ffn2", 768, 3072, {}), ("conv2-like", 768, 2304, {})]:
        x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
        xp, wp = SF.split_planes(x), SF.split_planes(w)
        y = torch.empty(M, N, device=dev)
        op = SF.tc_operand_plain(xp, M, K)
        ep = SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b, **kw)
        fn = lambda: SF.gemm_tc_kmajor(op, wp, N, K, ep)
        w_ms, c_ms = timeit(fn, False), timeit(fn, True)
        fl = 2.0 * M * N * K / 1e9
        print(f"shape {name:16s} {M}x{N}x{K}: warm {w_ms*1e3:6.1f} us {fl/w_ms:6.1f} TF/s | "
              f"cold {c_ms*1e3:6.1f} us {fl/c_ms:6.1f} TF/s")
        g = rnd(M, N, seed=4)
        gp = SF.split_planes(g)
        dW = torch.empty(K, N, device=dev)
        fn = lambda: SF.gemm_tc_wgrad(op, gp, N, K, dW)
        w_ms, c_ms = timeit(fn, False), timeit(fn, True)
        print(f"      wgrad            K={K} N={N}: warm {w_ms*1e3:6.1f} us {fl/w_ms:6.1f} TF/s | "
              f"cold {c_ms*1e3:6.1f} us {fl/c_ms:6.1f} TF/s")


def t_ffn():
    """FFN-like chain pieces with every epilogue variant, odd sizes; each vs fp64."""
    import itertools
    for (M, D, Fh) in [(777, 64, 3072), (75, 32, 3072), (450, 128, 3072), (1000, 256, 1024)]:
        x, W1, b1 = rnd(M, D, seed=1), rnd(D, Fh, seed=2, scale=D ** -0.5), rnd(Fh, seed=3, scale=0.1)
        W2, b2 = rnd(Fh, D, seed=4, scale=Fh ** -0.5), rnd(D, seed=5, scale=0.1)
        dy = rnd(M, D, seed=6)
        xd, W1d, b1d, W2d, b2d, dyd = (t.double() for t in (x, W1, b1, W2, b2, dy))
        h_ref = torch.relu(xd @ W1d + b1d)
        h = torch.empty(M, Fh, device=dev)
        SF.mm_fwd(x, W1, dict(bias=b1, relu=1), h, M, Fh, D)
        print(f"ffn M={M} D={D} F={Fh}: h {rel(h, h_ref):.2e}", end=" ")
        y = torch.empty(M, D, device=dev)
        SF.mm_fwd(h, W2, dict(bias=b2), y, M, D, Fh)
        print(f"y {rel(y, h.double() @ W2d + b2d):.2e}", end=" ")
        dW2 = torch.empty(Fh, D, device=dev)
        SF.mm_wgrad(h, dy, dW2, M, D, Fh)
        print(f"dW2 {rel(dW2, h.double().t() @ dyd):.2e}", end=" ")
        dh = torch.empty(M, Fh, device=dev)
        SF.mm_dgrad(dy, W2, dict(mask_src=h, mask_scale=1.25), dh, M, D, Fh)
        dh_ref = (dyd @ W2d.t()) * (h.double() > 0) * 1.25
        print(f"dh {rel(dh, dh_ref):.2e}", end=" ")
        dW1 = torch.empty(D, Fh, device=dev)
        SF.mm_wgrad(x, dh, dW1, M, Fh, D)
        print(f"dW1 {rel(dW1, xd.t() @ dh.double()):.2e}", end=" ")
        dx = torch.empty(M, D, device=dev)
        SF.mm_dgrad(dh, W1, {}, dx, M, Fh, D)
        print(f"dx {rel(dx, dh.double() @ W1d.t()):.2e}")


def t_convgrad():
    """conv data / weight gradients through the autograd function, tc vs fp64 F.conv1d"""
    for (B, L, Cin, Cout, k, s) in [(2, 256, 64, 64, 3, 1), (3, 600, 128, 128, 3, 1),
                                    (3, 600, 128, 128, 3, 2), (3, 601, 128, 128, 3, 2),
                                    (3, 600, 128, 128, 1, 2), (2, 301, 64, 128, 1, 2),
                                    (4, 1000, 768, 768, 3, 2), (4, 1000, 768, 768, 1, 2)]:
        x = rnd(B, L, Cin, seed=1).requires_grad_(True)
        w = rnd(Cout, Cin, k, seed=2, scale=(Cin * k) ** -0.5).requires_grad_(True)
        b = rnd(Cout, seed=3, scale=0.1).requires_grad_(True)
        Wg = w.permute(2, 1, 0).reshape(k * Cin, Cout).contiguous()
        y = SF.conv1d_cl(x, Wg, b, k, s)
        xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
        yr = F.conv1d(xr.transpose(1, 2), wr, br, stride=s, padding=1 if k == 3 else 0).transpose(1, 2)
        g = rnd(*y.shape, seed=4)
        y.backward(g)
        yr.backward(g.double())
        print(f"convgrad B={B} L={L} Cin={Cin} Cout={Cout} k={k} s={s}: y {rel(y, yr):.2e} "
              f"dx {rel(x.grad, xr.grad):.2e} dw {rel(w.grad, wr.grad):.2e} db {rel(b.grad, br.grad):.2e}")


def t_streamk():
    """Stream-K remainder (include/ssb.h ssb_gemm_tc_set_streamk_workspace) against the classic schedule
    and fp64: plain, accumulate + plane-emitting (wide) epilogues; repeated launches (flag re-arming);
    timing of both schedules."""
    from silent_speech_b200 import _lib
    lib = _lib.load()
    SF._ensure_streamk(lib, torch.device("cuda", torch.cuda.current_device()))
    ws = SF._streamk_ws[torch.cuda.current_device()]
    assert ws is not None, "run with SSB_STREAMK=1"

    def attach(on):
        _lib.check(lib.ssb_gemm_tc_set_streamk_workspace(ws.data_ptr() if on else None, ws.numel() if on else 0))

    def ms(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    shapes = [(16000, 768, 3072), (16000, 3072, 768), (16000, 768, 768), (16000, 2304, 768), (16000, 768, 2304),
              (32000, 768, 2304), (4800, 256, 768), (4800, 256, 2816), (601, 2048, 1024), (38400, 128, 896),
              (1000, 768, 768), (333, 80, 768), (24000, 768, 3072), (130, 256, 512)]
    for (M, N, K) in shapes:
        x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
        xp, wp = SF.split_planes(x), SF.split_planes(w)
        ref = x.double() @ w.double().t() + b.double()
        res0 = rnd(M, N, seed=4)

        def run(y, pl, variant):
            if variant == 0:
                epi = SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b)
            else:   # accumulate in place + planes of leaky_relu(result): the wide lane map
                y.copy_(res0)
                epi = SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b, accumulate=1, planes_out=pl,
                              planes_lrelu=0.1)
            SF.gemm_tc_kmajor(SF.tc_operand_plain(xp, M, K), wp, N, K, epi)
        out = {}
        for variant in ((0, 1) if N % 8 == 0 else (0,)):
            for on in (1, 0, 1):
                attach(on)
                y = torch.full((M, N), float("nan"), device=dev)
                pl = torch.zeros((2, M, N), dtype=torch.bfloat16, device=dev)
                run(y, pl, variant)
                torch.cuda.synchronize()
                want = ref if variant == 0 else ref + res0.double()
                e = rel(y, want)
                key = (variant, on)
                if key in out:
                    assert torch.equal(out[key][0], y), "stream-K result not reproducible"
                out[key] = (y, e)
                if variant == 1:
                    rec = pl[0].float() + pl[1].float()
                    lr = torch.where(y > 0, y, y * 0.1)
                    assert rel(rec, lr) < 1e-4, rel(rec, lr)
                assert e < 2e-5 and torch.isfinite(y).all(), (M, N, K, variant, on, e)
        attach(1)
        y = torch.empty((M, N), device=dev)
        pl = torch.empty((2, M, N), dtype=torch.bfloat16, device=dev)
        t1 = ms(lambda: run(y, pl, 0))
        attach(0)
        t0 = ms(lambda: run(y, pl, 0))
        attach(1)
        d = (out[(0, 1)][0] - out[(0, 0)][0]).abs().max().item()
        print(f"streamk {M}x{N}x{K}: rel err on {out[(0, 1)][1]:.2e} off {out[(0, 0)][1]:.2e} max|on-off| {d:.2e}  "
              f"{t1 * 1e3:.1f} us on / {t0 * 1e3:.1f} us off ({t0 / t1:.3f}x)", flush=True)
    assert int(ws[:8192].max()) == 0, "stream-K flags not re-armed"
    print("streamk ok")


if __name__ == "__main__":
    what = sys.argv[1]
    {"streamk": t_streamk, "split": t_split, "kmajor": t_kmajor, "wgrad": t_wgrad, "conv1": lambda: t_conv(1),
     "conv2": lambda: t_conv(2), "perf": t_perf, "shapes": t_shapes, "ffn": t_ffn, "convgrad": t_convgrad}[what]()

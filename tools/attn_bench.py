"""CUDA-event timing of the attention schedules at the cfg-1 shape (B=32, T=500, H=8, dh=96).

    python tools/attn_bench.py [fused|tc ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import functional as SF  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B, T, H, dh, W, p = 32, 500, 8, 96, 99, 0.2
    D = H * dh
    qkv = torch.randn(B * T, 3 * D, device="cuda").requires_grad_(True)
    E = torch.zeros(H, 200, dh, device="cuda")
    E[:, :199] = torch.randn(H, 199, dh, device="cuda") * dh ** -0.5
    go = torch.randn(B * T, D, device="cuda")
    for sched in (sys.argv[1:] or ["fused", "tc"]):
        os.environ["SSB_ATTN"] = sched
        o = SF.band_attention(qkv, E, B, T, H, dh, W, p, 1, 0)
        fwd = timeit(lambda: SF.band_attention(qkv, E, B, T, H, dh, W, p, 1, 0))

        def fb():
            qkv.grad = None
            SF.band_attention(qkv, E, B, T, H, dh, W, p, 1, 0).backward(go)
        tot = timeit(fb)
        print(f"{sched}: fwd {fwd:.0f} us, fwd+bwd {tot:.0f} us (bwd {tot - fwd:.0f} us) per layer "
              f"(includes pad/split, positional GEMMs)")
        if sched == "fused":
            lib = SF._lib.load()
            # the two fused kernels alone
            M, BH, RW = B * T, B * H, 200
            qkvp = torch.empty((2, M, 3 * H, 128), dtype=torch.bfloat16, device="cuda")
            SF._lib.check(lib.ssb_pad_split_heads(qkv.data_ptr(), M, 3 * D, 0, 3 * H, dh, qkvp.data_ptr(), SF._stream()))
            R = torch.randn(BH, T, RW, device="cuda")
            O = torch.empty(M, D, device="cuda")
            Op = torch.empty((2, M, D), dtype=torch.bfloat16, device="cuda")
            stats = torch.empty(2, BH, T, device="cuda")
            f = lambda: SF._lib.check(lib.ssb_attn_fused_fwd(
                qkvp.data_ptr(), R.data_ptr(), B, T, H, dh, W, RW, p, 1, 0, O.data_ptr(), Op.data_ptr(),
                stats[0].data_ptr(), stats[1].data_ptr(), 128, SF._stream()))
            print(f"  attn_fused_fwd_kernel alone: {timeit(f):.0f} us")
            dop = torch.empty((2, M, H, 128), dtype=torch.bfloat16, device="cuda")
            SF._lib.check(lib.ssb_pad_split_heads(go.data_ptr(), M, D, 0, H, dh, dop.data_ptr(), SF._stream()))
            delta = torch.zeros(BH, T, device="cuda")
            dqkv = torch.zeros(M, D, device="cuda")
            dqp = torch.empty((2, M, 3 * D), dtype=torch.bfloat16, device="cuda")
            dsb = torch.zeros((2, M, H, 256), dtype=torch.bfloat16, device="cuda")
            g = lambda: SF._lib.check(lib.ssb_attn_fused_bwd(
                qkvp.data_ptr(), dop.data_ptr(), R.data_ptr(), stats[0].data_ptr(),
                stats[1].data_ptr(), delta.data_ptr(), B, T, H, dh, W, RW, p, 1, 0,
                dqkv.data_ptr(), D, dqp.data_ptr(), dsb.data_ptr(), 256, 128, 128, SF._stream()))
            print(f"  attn_fused_bwd_kernel alone: {timeit(g):.0f} us")


if __name__ == "__main__":
    main()

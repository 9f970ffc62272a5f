"""Launch one kernel family at its bench geometry a few times: the command `ncu` wraps.

    python tools/profile_targets.py {mel|attn|dtw|ctc|gemm} [n]

  mel   mel_kernel on 1024 clips x 10 s (bench.py mel_side_metric geometry)
  attn  attn_fused_fwd / attn_fused_bwd at cfg-1 (B=32, T=500, H=8, dh=96, dropout 0.2)
  dtw   dtw_fill + backtrace on 2000 cfg-2 pairs (500 x 600 cdist matrices, 2.4 GB > L2)
  ctc   ctc_fused_kernel at cfg-5 (32 x 750 x 38, 100-char targets)
  gemm  the three in-step FFN launches of bench.gemm_roofline, once each
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import _lib  # noqa: E402
from silent_speech_b200 import functional as SF  # noqa: E402


def mel(n):
    from silent_speech_b200 import data_utils as du
    B = 1024
    y = (torch.rand(B, 220500, device="cuda") * 2 - 1) * 0.5
    lib = _lib.load()
    basis, begin, end = du._basis_for(22050, 1024, 80, 0, 8000, y.device)
    out = torch.empty(B, 80, 861, device="cuda")
    for _ in range(n):
        _lib.check(lib.ssb_mel_fwd(y.data_ptr(), B, 220500, y.stride(0), 1024, 256, 1024,
                                   basis.data_ptr(), begin.data_ptr(), end.data_ptr(), 80, 1e-5,
                                   out.data_ptr(), None, _lib.current_stream()))


def attn(n):
    B, T, H, dh, W, p = 32, 500, 8, 96, 99, 0.2
    D = H * dh
    qkv = torch.randn(B * T, 3 * D, device="cuda").requires_grad_(True)
    E = torch.zeros(H, 200, dh, device="cuda")
    E[:, :199] = torch.randn(H, 199, dh, device="cuda") * dh ** -0.5
    go = torch.randn(B * T, D, device="cuda")
    for _ in range(n):
        qkv.grad = None
        SF.band_attention(qkv, E, B, T, H, dh, W, p, 1, 0).backward(go)


def dtw(n):
    from silent_speech_b200 import align
    P, Tp, Tg = 2000, 500, 600
    g = torch.Generator(device="cuda").manual_seed(1234)
    cost = torch.cdist(torch.randn(P, Tp, 80, device="cuda", generator=g),
                       torch.randn(P, Tg, 80, device="cuda", generator=g))
    for _ in range(n):
        align.align_batch(cost.transpose(1, 2))


def ctc(n):
    from silent_speech_b200.losses import ctc_loss
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(32, 750, 38, device="cuda", generator=g).requires_grad_(True)
    y = torch.randint(0, 37, (32, 100), device="cuda", generator=g)
    il = torch.full((32,), 750, dtype=torch.int64, device="cuda")
    tl = torch.full((32,), 100, dtype=torch.int64, device="cuda")
    for _ in range(n):
        ctc_loss(logits, y, il, tl, blank=37, reduction='sum')


def gemm(n):
    import bench
    launches, keep = bench.ffn_instep_launches()
    for _ in range(n):
        for _, fn, _ in launches:
            fn()


def ln(n):
    """add + dropout + LayerNorm forward / backward at cfg-1 (16000 x 768, p = 0.2), as the encoder blocks call them
    (planes-only branch gradient, parameter gradients into sinks)"""
    M, D = 16000, 768
    res, br = torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda")
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    dy = torch.randn(M, D, device="cuda")
    for _ in range(n):
        y, z, stat = SF._ln_fwd(res, br, g, b, 0.2, 7, 3, 1e-5, True)
        SF._ln_bwd(dy, z, stat, g, 0.2, 7, 3, (None, None), planes_only=True)


def thin(n):
    """first ResBlock at cfg-1 (8-channel EMG in, 768 out, 32 x 4000 samples): the thin-K CUDA-core kernels"""
    from absl import flags
    from silent_speech_b200 import architecture
    if not flags.FLAGS.is_parsed():
        flags.FLAGS(["x"])
    blk = architecture.ResBlock(8, 768, 2).cuda().train()
    x = torch.randn(32, 4000, 8, device="cuda")
    for _ in range(n):
        for p_ in blk.parameters():
            p_.grad = None
        blk.forward_cl(x).square().mean().backward()


if __name__ == "__main__":
    what = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(0)
    {"mel": mel, "attn": attn, "dtw": dtw, "ctc": ctc, "gemm": gemm, "thin": thin, "ln": ln}[what](n)
    torch.cuda.synchronize()
    print(what, "done")

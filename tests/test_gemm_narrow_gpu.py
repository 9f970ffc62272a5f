"""GPU: the GEMM-engine features the vocoder path added (csrc/gemm_tc.cu), each against fp64 torch:
N-adaptive MMAs for outputs narrower than a 256-column tile, dilated / many-tap im2col operands
(ssb_tc_operand_t.s_tap = dilation, up to 16 taps), the leaky-ReLU plane epilogue with in-place fp32
accumulation, and the opt-in stream-K remainder schedule."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("M", [100, 128, 300, 5000])
@pytest.mark.parametrize("N", [32, 64, 96, 128, 160, 224, 256, 288])
def test_narrow_outputs(M, N):
    """N < 256 runs an MMA of N = round32(N) columns (CTA pair and single-CTA forms)."""
    from silent_speech_b200 import functional as SF
    K = 192
    x, w, b = _rnd(M, K, seed=1), _rnd(N, K, seed=2, scale=K ** -0.5), _rnd(N, seed=3)
    y = torch.full((M, N), float("nan"), device="cuda")
    xp, wp = SF.split_planes(x), SF.split_planes(w)     # the operand descriptor holds a raw pointer: keep xp alive
    SF.gemm_tc_kmajor(SF.tc_operand_plain(xp, M, K), wp, N, K,
                      SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b, relu=1))
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    assert torch.isfinite(y).all() and _rel(y, ref) < 2e-5


@pytest.mark.parametrize("C,Cout,k,dil", [(32, 32, 3, 1), (32, 32, 11, 5), (64, 64, 7, 3), (128, 128, 11, 1),
                                           (256, 256, 3, 5), (96, 512, 7, 1)])
def test_dilated_many_tap_convolution_with_lrelu_planes(C, Cout, k, dil):
    """y = x + conv1d(lrelu(x), w, dilation) in place, planes of lrelu(y): one residual step of
    hifi_gan/models.py:40-44 as the vocoder issues it (C_out != C: no residual)."""
    from silent_speech_b200 import functional as SF
    from silent_speech_b200._lib import Scatter, TcOperand
    L = 777
    x = _rnd(L, C, seed=1)
    w = _rnd(Cout, C, k, seed=2, scale=(C * k) ** -0.5)
    b = _rnd(Cout, seed=3, scale=0.1)
    residual = Cout == C
    a = torch.where(x > 0, x, 0.1 * x)
    ap = SF.split_planes(a)
    wp = SF.split_planes(w.permute(0, 2, 1).reshape(Cout, k * C).contiguous())
    y = x.clone() if residual else torch.full((L, Cout), float("nan"), device="cuda")
    pl = torch.zeros((2, L, Cout), dtype=torch.bfloat16, device="cuda")
    pad = (k * dil - dil) // 2
    op = TcOperand(ap.data_ptr(), L * C, L * C, 1, L, L, C, C, 1, dil, -pad)
    SF.gemm_tc_kmajor(op, wp, Cout, k * C,
                      SF._epi(Scatter(y.data_ptr(), 0, L, Cout, 1, 0), bias=b, accumulate=int(residual),
                              planes_out=pl, planes_lrelu=0.1))
    ref = torch.nn.functional.conv1d(a.double().t()[None], w.double(), b.double(), dilation=dil, padding=pad)[0].t()
    if residual:
        ref = ref + x.double()
    assert _rel(y, ref) < 2e-5
    want = torch.where(y > 0, y, 0.1 * y)
    assert _rel(pl[0].float() + pl[1].float(), want) < 1e-5       # 16 mantissa bits of the result just stored


def test_streamk_opt_in_schedule():
    """tools/tc_gemm_check.py streamk: the stream-K remainder against the classic schedule and fp64 on
    14 shapes, two epilogue variants, repeated launches; own process because SSB_STREAMK is read once."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tc_gemm_check.py"), "streamk"],
                       capture_output=True, text=True, timeout=600, env={**os.environ, "SSB_STREAMK": "1"})
    assert r.returncode == 0 and "streamk ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

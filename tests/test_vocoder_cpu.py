"""CPU checks of the HiFi-GAN vocoder path:
  * the oracle (oracle/vocoder.py) against audio produced by the executed reference generator
    (tests/golden/vocoder_golden.npz, made by tests/golden/make_golden_vocoder.py);
  * the host logic of silent_speech_b200/vocoder.py - weight layouts, the stride-phase form of
    ConvTranspose1d, dilated im2col geometry, channel padding, in-place residual accumulation -
    with its four C-ABI primitives re-implemented in plain torch from their documented semantics
    (include/ssb.h).  No kernel runs here; the GPU tests run the same geometry through libssb.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vocoder as ov
from silent_speech_b200 import vocoder as sv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vocoder_golden.npz")


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class EmuGenerator(sv.Generator):
    """Generator with the libssb primitives restated on CPU tensors (bf16 split planes included)."""

    def __init__(self, h):
        super().__init__(h)
        self.device = torch.device("cpu")

    def _split(self, x):
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return torch.stack([hi, lo])

    def _gemm(self, cv, src, rows_out, out=None, accumulate=False, planes=True):
        sp, row0, L_src = src
        C = sp.shape[2]
        assert C == cv.C and cv.K == cv.taps * C
        a = (sp[0].double() + sp[1].double())[row0:row0 + L_src]          # rows outside: zero padding
        cols = []
        r = torch.arange(rows_out)
        for tap in range(cv.taps):
            idx = r + tap * cv.s_tap + cv.off
            ok = (idx >= 0) & (idx < L_src)
            g = torch.zeros((rows_out, C), dtype=torch.float64)
            g[ok] = a[idx[ok]]
            cols.append(g)
        A = torch.cat(cols, 1)                                           # (rows_out, taps * C)
        W = cv.wp[0].double() + cv.wp[1].double()                         # (N, K)
        res = (A @ W.t() + cv.bias.double()).float()
        if out is not None:
            ot, orow = out
            assert ot.shape[1] == cv.N
            if accumulate:
                res = res + ot[orow:orow + rows_out]
            ot[orow:orow + rows_out] = res
        else:
            assert not accumulate
        if not planes:
            return None
        return self._split(torch.where(res > 0, res, res * sv.LRELU_SLOPE))

    def _mix(self, branches, scale, slope):
        v = sum(branches) * scale
        return self._split(torch.where(v > 0, v, v * slope))

    def _sum(self, branches):
        return sum(branches)

    def _post(self, branches, scale, slope, w, bias):
        v = sum(branches) * scale
        v = torch.where(v > 0, v, v * slope)
        taps = w.shape[0]
        L = v.shape[0]
        vp = torch.nn.functional.pad(v, (0, 0, taps // 2, taps // 2))
        acc = torch.full((L,), bias, dtype=torch.float32)
        for tap in range(taps):
            acc = acc + vp[tap:tap + L] @ w[tap]
        return torch.tanh(acc)


def test_oracle_matches_reference_golden():
    gold = np.load(GOLD)
    for name, (cfg, T, gain) in ov.GOLDEN_CASES.items():
        sd = ov.formula_state_dict(cfg, gain=gain)
        audio = ov.generator_forward(sd, ov.formula_mel(T), cfg)
        ref = torch.from_numpy(gold[f"{name}_audio"])
        assert audio.shape == ref.shape == (T * int(np.prod(cfg["upsample_rates"])),)
        assert _rel(audio, ref) < 1e-6, name


def test_oracle_weight_norm_fold_matches_reference_golden():
    gold = np.load(GOLD)
    cfg, T, gain = ov.GOLDEN_CASES["rb1"]
    wn = ov.weight_normed_state_dict(ov.formula_state_dict(cfg, gain=gain))
    audio = ov.generator_forward(ov.fold_state_dict(wn), ov.formula_mel(T), cfg)
    assert _rel(audio, torch.from_numpy(gold["rb1_wn_audio"])) < 1e-5


@pytest.mark.parametrize("name", list(ov.GOLDEN_CASES))
def test_host_geometry_against_oracle(name):
    """Phase-form transposed convolutions, dilated im2col views, padding and in-place residuals,
    evaluated with emulated primitives, reproduce the reference audio to the accuracy of the
    16-mantissa-bit operand planes."""
    cfg, T, gain = ov.GOLDEN_CASES[name]
    sd = ov.formula_state_dict(cfg, gain=gain)
    g = EmuGenerator(cfg)
    g.load_state_dict(sd)
    audio = g.forward_one(ov.formula_mel(T))
    ref = torch.from_numpy(np.load(GOLD)[f"{name}_audio"])
    assert audio.shape == ref.shape
    assert _rel(audio, ref) < 5e-5, _rel(audio, ref)


def test_host_geometry_odd_upsampling_and_four_kernels():
    """k != 2u (three taps per phase, ragged last tap), an odd frame count and four residual
    branches (the pre-sum path)."""
    cfg = dict(resblock="2", upsample_rates=[4, 3], upsample_kernel_sizes=[10, 7],
               upsample_initial_channel=24, resblock_kernel_sizes=[3, 5, 3, 7],
               resblock_dilation_sizes=[[1, 2], [2, 3], [1, 1], [3, 1]])
    sd = ov.formula_state_dict(cfg)
    mel = ov.formula_mel(11)
    g = EmuGenerator(cfg)
    g.load_state_dict(sd)
    audio = g.forward_one(mel)
    ref = ov.generator_forward(sd, mel, cfg)
    assert audio.shape == ref.shape
    assert _rel(audio, ref) < 5e-5, _rel(audio, ref)


def test_load_state_dict_folds_weight_norm_and_is_strict():
    cfg, T, gain = ov.GOLDEN_CASES["rb2"]
    sd = ov.formula_state_dict(cfg, gain=gain)
    g = sv.Generator(cfg)
    g.load_state_dict(ov.weight_normed_state_dict(sd))
    want = ov.fold_state_dict(ov.weight_normed_state_dict(sd))
    got = g.state_dict()
    assert set(got) == set(want) == set(ov.shapes(cfg))
    for k in want:
        assert torch.equal(got[k], want[k]), k
    bad = dict(sd)
    bad.pop("conv_post.bias")
    with pytest.raises(RuntimeError):
        sv.Generator(cfg).load_state_dict(bad)
    bad = dict(sd)
    bad["conv_pre.weight"] = bad["conv_pre.weight"][:, :40]
    with pytest.raises(RuntimeError):
        sv.Generator(cfg).load_state_dict(bad)


def test_vocoder_reads_checkpoint_like_the_reference(tmp_path):
    """vocoder.py:17-26: config.json next to the checkpoint, torch.load(...)['generator']."""
    cfg, T, gain = ov.GOLDEN_CASES["rb1"]
    sd = ov.weight_normed_state_dict(ov.formula_state_dict(cfg, gain=gain))
    with open(tmp_path / "config.json", "w") as f:
        json.dump(dict(cfg, num_mels=80, sampling_rate=22050), f)
    torch.save({"generator": sd}, tmp_path / "g_00000001")
    v = sv.Vocoder(checkpoint_file=str(tmp_path / "g_00000001"))
    assert v.generator.h.upsample_rates == cfg["upsample_rates"]
    folded = ov.fold_state_dict(sd)
    assert all(torch.equal(v.generator.state_dict()[k], folded[k]) for k in folded)


def test_no_cpu_path():
    cfg, T, gain = ov.GOLDEN_CASES["rb2"]
    g = sv.Generator(cfg)
    with pytest.raises(RuntimeError):
        g.to("cpu")
    g.load_state_dict(ov.formula_state_dict(cfg, gain=gain))
    with pytest.raises(Exception):
        g(torch.zeros(1, 80, 4))

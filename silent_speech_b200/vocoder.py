"""HiFi-GAN generator inference on libssb (second half).

Mirrors the reference's vocoder.py:16-36 (`Vocoder`) and the `Generator` of hifi_gan/models.py:75-116
that it drives: same constructor arguments, same checkpoint / state_dict keys (with or without
weight norm), a (seq_len, 80) mel in, a 1-D audio tensor out.  There is no CPU path.

B200-first layout instead of the reference's NCL convolutions:

  * activations are channels-last (samples, C); every convolution is ONE tcgen05 GEMM
    (csrc/gemm_tc.cu, bf16x3 = fp32-class arithmetic) whose A operand is a TMA im2col view of the
    previous layer's bf16 split planes: tap step = dilation, out-of-range rows read as the zero
    padding.  K = taps * C.
  * ConvTranspose1d(k, stride u, padding (k - u) / 2) is evaluated in its stride-phase form: output
    sample p = i' * u + r' - pad receives  sum_j  x[i' - j] . w[:, :, r' + u j]  (j < ceil(k / u)),
    i.e. a GEMM with ceil(k / u) taps and N = u * C_out whose (rows, u * C_out) result, read as
    (rows * u, C_out), IS the upsampled signal shifted by `pad` rows: no zero-stuffing, no scatter.
  * the GEMM epilogue adds the bias, accumulates the residual (x = x + c2(...), models.py:44) in
    fp32 in place and writes leaky_relu(result) as the next convolution's operand planes
    (ssb_epilogue_t.planes_lrelu), so no activation pass and no fp32 -> bf16 split pass runs
    between two convolutions.
  * the multi-receptive-field mean (models.py:101-108) + leaky ReLU is one kernel writing operand
    planes (ssb_voc_mix); the last one is fused with conv_post (one output channel) and tanh
    (ssb_voc_post).
  * channel counts are padded to a multiple of 32 (one tensor-core k-block) with zero weights: padded
    channels stay exactly 0 through bias-free zero filters and leaky ReLU.
"""
import contextlib
import json
import math
import os

import torch

from . import _lib
from .functional import _epi, _stream, gemm_tc_kmajor, split_planes
from ._lib import Scatter, TcOperand

# hifi_gan/config_v1.json (the generator the reference's published vocoder checkpoint uses)
CONFIG_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                 upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])

LRELU_SLOPE = 0.1     # hifi_gan/models.py:8
POST_SLOPE = 0.01     # F.leaky_relu default (models.py:109)
_f32 = torch.float32


def _pad32(c):
    return (c + 31) // 32 * 32


def get_padding(kernel_size, dilation=1):
    """hifi_gan/utils.py:36-37"""
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(g, v):
    """w = v * (g / ||v||), norm over every dimension but the first (torch.nn.utils.weight_norm,
    dim = 0): what Generator.remove_weight_norm() leaves in `.weight` (vocoder.py:26)."""
    dims = tuple(range(1, v.dim()))
    return v * (g / torch.linalg.vector_norm(v, 2, dims, keepdim=True))


class AttrDict(dict):
    """hifi_gan/env.py:5-8"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class _Conv:
    """One convolution as a GEMM: operand planes of the weights + geometry of the im2col view."""
    __slots__ = ("wp", "bias", "N", "K", "C", "taps", "s_tap", "off")

    def __init__(self, wp, bias, N, K, C, taps, s_tap, off):
        self.wp, self.bias, self.N, self.K, self.C = wp, bias, N, K, C
        self.taps, self.s_tap, self.off = taps, s_tap, off


def _operand(ptr, plane_stride, rows_out, L_src, C, s_tap, off):
    return TcOperand(ptr, plane_stride, plane_stride, 1, rows_out, L_src, C, C, 1, s_tap, off)


class Generator:
    """Inference-only HiFi-GAN generator with the reference's interface (hifi_gan/models.py:75-116):
    Generator(h) with h the hyper-parameter dict of config_v*.json, load_state_dict() of a
    reference checkpoint's 'generator' entry, `.to(device)`, `.eval()`, `.remove_weight_norm()`,
    and `generator(x)` with x (B, num_mels, T) -> (B, 1, T * prod(upsample_rates))."""

    def __init__(self, h, num_mels=80):
        self.h = AttrDict(h)
        self.num_mels = num_mels          # models.py:80 hard-codes 80 input channels
        self.num_kernels = len(self.h.resblock_kernel_sizes)
        self.num_upsamples = len(self.h.upsample_rates)
        if str(self.h.resblock) not in ("1", "2"):
            raise ValueError(f"resblock must be '1' or '2', got {self.h.resblock!r}")
        self.device = torch.device("cuda")
        self._sd = None
        self._prep = None

    # ---- nn.Module-shaped surface ------------------------------------------------------------
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("silent_speech_b200.vocoder.Generator has no CPU path")
        if device != self.device:
            self.device, self._prep = device, None
        return self

    def cuda(self):
        return self.to("cuda")

    def eval(self):
        return self

    def remove_weight_norm(self):
        """Weight norm is folded when the checkpoint is loaded (load_state_dict)."""
        return self

    def expected_shapes(self):
        h, c0 = self.h, self.h.upsample_initial_channel
        out = {"conv_pre.weight": (c0, self.num_mels, 7), "conv_pre.bias": (c0,)}
        ch = c0
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
            out[f"ups.{i}.weight"] = (cin, ch, k)
            out[f"ups.{i}.bias"] = (ch,)
            for j, (rk, dil) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                p = f"resblocks.{i * self.num_kernels + j}"
                for m in range(len(dil)):
                    names = ((f"{p}.convs1.{m}", f"{p}.convs2.{m}") if str(h.resblock) == "1"
                             else (f"{p}.convs.{m}",))
                    for n in names:
                        out[n + ".weight"] = (ch, ch, rk)
                        out[n + ".bias"] = (ch,)
        out["conv_post.weight"] = (1, ch, 7)
        out["conv_post.bias"] = (1,)
        return out

    def load_state_dict(self, state_dict, strict=True):
        """Accepts the generator as a checkpoint stores it (weight_g / weight_v per convolution,
        vocoder.py:24) or after remove_weight_norm() (plain .weight)."""
        sd = {}
        for k, t in state_dict.items():
            if k.endswith(".weight_g"):
                base = k[:-len(".weight_g")]
                sd[base + ".weight"] = fold_weight_norm(t.detach().to(_f32),
                                                        state_dict[base + ".weight_v"].detach().to(_f32))
            elif not k.endswith(".weight_v"):
                sd[k] = t.detach().to(_f32)
        want = self.expected_shapes()
        missing = [k for k in want if k not in sd]
        unexpected = [k for k in sd if k not in want]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Generator.load_state_dict: missing {missing[:4]}, unexpected {unexpected[:4]}")
        for k, shp in want.items():
            if k in sd and tuple(sd[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(shp)}")
        self._sd = {k: v.clone() for k, v in sd.items() if k in want}
        self._prep = None
        return missing, unexpected

    def state_dict(self):
        return dict(self._sd or {})

    # ---- weights -> GEMM operands (once per load) -----------------------------------------------
    def _conv(self, name, dilation):
        w = self._sd[name + ".weight"].to(self.device)               # (Cout, Cin, k)
        b = self._sd[name + ".bias"].to(self.device)
        Cout, Cin, k = w.shape
        Np, Cp = _pad32(Cout), _pad32(Cin)
        wg = torch.zeros((Np, k, Cp), dtype=_f32, device=self.device)
        wg[:Cout, :, :Cin] = w.permute(0, 2, 1)                       # K index = tap * Cp + ci
        bp = torch.zeros(Np, dtype=_f32, device=self.device)
        bp[:Cout] = b
        return _Conv(self._split(wg.view(Np, k * Cp).contiguous()), bp, Np, k * Cp, Cp, k, dilation,
                     -get_padding(k, dilation))

    def _up(self, name, u, k):
        w = self._sd[name + ".weight"].to(self.device)               # (Cin, Cout, k)
        b = self._sd[name + ".bias"].to(self.device)
        Cin, Cout, _ = w.shape
        Np, Cp = _pad32(Cout), _pad32(Cin)
        J = (k + u - 1) // u
        # B[(r', co), (tau, ci)] = w[ci, co, r' + u * (J - 1 - tau)]: tap tau reads x[i' + tau - (J - 1)]
        wg = torch.zeros((u, Np, J, Cp), dtype=_f32, device=self.device)
        for tau in range(J):
            j = J - 1 - tau
            kk = min(k, u * (j + 1)) - u * j                          # phases with r' + u j < k
            if kk > 0:
                wg[:kk, :Cout, tau, :Cin] = w[:, :, u * j:u * j + kk].permute(2, 1, 0)
        bp = torch.zeros((u, Np), dtype=_f32, device=self.device)
        bp[:, :Cout] = b
        return _Conv(self._split(wg.view(u * Np, J * Cp).contiguous()), bp.view(-1).contiguous(),
                     u * Np, J * Cp, Cp, J, 1, -(J - 1))

    def _prepare(self):
        if self._sd is None:
            raise RuntimeError("Generator: load_state_dict() first")
        h = self.h
        prep = {"pre": self._conv("conv_pre", 1), "ups": [], "blocks": []}
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            prep["ups"].append(self._up(f"ups.{i}", u, k))
            stage = []
            for j, (rk, dil) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                p = f"resblocks.{i * self.num_kernels + j}"
                if str(h.resblock) == "1":
                    stage.append([(self._conv(f"{p}.convs1.{m}", d), self._conv(f"{p}.convs2.{m}", 1))
                                  for m, d in enumerate(dil)])
                else:
                    stage.append([(self._conv(f"{p}.convs.{m}", d),) for m, d in enumerate(dil)])
            prep["blocks"].append(stage)
        wpost = self._sd["conv_post.weight"].to(self.device)          # (1, C, 7)
        Cp = _pad32(wpost.shape[1])
        wp = torch.zeros((wpost.shape[2], Cp), dtype=_f32, device=self.device)
        wp[:, :wpost.shape[1]] = wpost[0].t()
        prep["post_w"] = wp.contiguous()
        prep["post_taps"] = int(wpost.shape[2])
        prep["post_b"] = float(self._sd["conv_post.bias"][0])
        self._prep = prep
        return prep

    # ---- forward -------------------------------------------------------------------------------
    # The four primitives below are the only places that touch the C ABI; everything above them is
    # geometry (tests/test_vocoder_cpu.py re-implements them in plain torch to check that geometry
    # against the oracle without a GPU).
    def _split(self, x):
        return split_planes(x)

    def _gemm(self, cv, src, rows_out, out=None, accumulate=False, planes=True):
        """One convolution.  src = (planes (2, rows_total, C) bf16, first row, rows): the A operand is
        the im2col view of those rows (rows outside read as zero padding).  out = (fp32 (rows_total,
        ld) tensor, first row) receives / accumulates the result; returns the (2, rows_out, N) planes
        of leaky_relu(result) when planes is set."""
        sp, row0, L_src = src
        C = sp.shape[2]
        pl = torch.empty((2, rows_out, cv.N), dtype=torch.bfloat16, device=sp.device) if planes else None
        if out is not None:
            ot, orow = out
            scat = Scatter(ot.data_ptr() + 4 * orow * ot.shape[1], 0, rows_out, ot.shape[1], 1, 0)
        else:
            scat = Scatter(None, 0, rows_out, cv.N, 1, 0)
        gemm_tc_kmajor(_operand(sp.data_ptr() + 2 * row0 * C, sp.shape[1] * C, rows_out, L_src, C,
                                cv.s_tap, cv.off),
                       cv.wp, cv.N, cv.K,
                       _epi(scat, bias=cv.bias, accumulate=int(accumulate), planes_out=pl,
                            planes_lrelu=LRELU_SLOPE if planes else None))
        return pl

    def _mix(self, branches, scale, slope):
        """planes of leaky_relu(scale * sum(branches)); branches: 1..3 contiguous (L, C) fp32."""
        lib = _lib.load()
        L, C = branches[0].shape
        ptrs = [b.data_ptr() for b in branches] + [None] * (3 - len(branches))
        out = torch.empty((2, L, C), dtype=torch.bfloat16, device=branches[0].device)
        _lib.check(lib.ssb_voc_mix(ptrs[0], ptrs[1], ptrs[2], L * C, scale, 1, slope, None,
                                   out.data_ptr(), L * C, _stream()))
        return out

    def _sum(self, branches):
        lib = _lib.load()
        ptrs = [b.data_ptr() for b in branches] + [None] * (3 - len(branches))
        out = torch.empty_like(branches[0])
        _lib.check(lib.ssb_voc_mix(ptrs[0], ptrs[1], ptrs[2], out.numel(), 1.0, 0, 0.0,
                                   out.data_ptr(), None, 0, _stream()))
        return out

    def _post(self, branches, scale, slope, w, bias):
        """tanh(conv_post(leaky_relu(scale * sum(branches)))) -> (L,)"""
        lib = _lib.load()
        L, C = branches[0].shape
        ptrs = [b.data_ptr() for b in branches] + [None] * (3 - len(branches))
        audio = torch.empty(L, dtype=_f32, device=branches[0].device)
        _lib.check(lib.ssb_voc_post(ptrs[0], ptrs[1], ptrs[2], L, C, w.shape[0], scale, slope,
                                    w.data_ptr(), bias, audio.data_ptr(), _stream()))
        return audio

    def _branch_streams(self, n):
        """Side streams for the residual branches (SSB_VOC_STREAMS=0: everything on the caller's stream)."""
        if self.device.type != "cuda" or os.environ.get("SSB_VOC_STREAMS", "1") == "0":
            return None
        pool = getattr(self, "_side_streams", None)
        if pool is None or len(pool) < n:
            pool = [torch.cuda.Stream(device=self.device) for _ in range(n)]
            self._side_streams = pool
        return pool

    def forward_one(self, mel):
        """mel (T, num_mels) fp32 on self.device -> audio (T * prod(rates),)."""
        prep = self._prep or self._prepare()
        h, dev, nk = self.h, self.device, self.num_kernels
        L = mel.shape[0]
        pre = prep["pre"]
        x0 = torch.zeros((L, pre.C), dtype=_f32, device=dev)
        x0[:, :self.num_mels] = mel
        cur = self._gemm(pre, (self._split(x0), 0, L), L)            # planes of lrelu(conv_pre(x))
        audio = None
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            up = prep["ups"][i]
            pad = (k - u) // 2
            Lout = (L - 1) * u - 2 * pad + k
            R = (pad + Lout - 1) // u + 1                             # rows i' of the phase GEMM
            C = up.N // u                                             # padded channel count
            full = torch.empty((R, up.N), dtype=_f32, device=dev)
            fullp = self._gemm(up, (cur, 0, L), R, out=(full, 0))
            L = Lout
            xf = full.view(R * u, C)                                  # rows [pad, pad + L): the signal
            xp = fullp.view(2, R * u, C)                              # planes of lrelu(signal)
            x = xf[pad:pad + L]
            branches = [x.clone() for _ in range(nk - 1)]
            # The residual branches of a stage are independent chains: each runs on its own stream.
            # Where one GEMM's tiles leave most CTA pairs idle (the first stage: 4800 rows x 256
            # channels = 19 tiles for 74 pairs) the branches run side by side; where a GEMM fills the
            # machine (150 / 300 / 600 tiles = 2.03 / 4.05 / 8.1 waves) the next branch's CTAs take the
            # SMs the partial last wave leaves idle.  Measured (600 frames, graph replay): one stream
            # 3.15 ms, first stage only 2.73 ms, every stage 2.36 ms.
            side = self._branch_streams(nk) if nk > 1 else None
            main = torch.cuda.current_stream() if side else None
            for j, convs in enumerate(prep["blocks"][i]):
                # branch 0 accumulates into the phase GEMM's output in place, the others into copies
                dst = (xf, pad) if j == 0 else (branches[j - 1], 0)
                src = (xp, pad, L)
                if side:
                    side[j].wait_stream(main)
                with (torch.cuda.stream(side[j]) if side else contextlib.nullcontext()):
                    for m, pair in enumerate(convs):
                        last = m == len(convs) - 1
                        if len(pair) == 2:                            # ResBlock1: models.py:40-44
                            hp = self._gemm(pair[0], src, L)
                            nxt = self._gemm(pair[1], (hp, 0, L), L, out=dst, accumulate=True,
                                             planes=not last)
                        else:                                         # ResBlock2: models.py:65-68
                            nxt = self._gemm(pair[0], src, L, out=dst, accumulate=True, planes=not last)
                        if not last:
                            src = (nxt, 0, L)
                    del src
            if side:
                for st in side[:nk]:
                    main.wait_stream(st)
            branches.insert(0, x)
            while len(branches) > 3:                                  # more than 3 kernels: pre-sum
                branches = branches[:-3] + [self._sum([b.contiguous() for b in branches[-3:]])]
            if i + 1 < self.num_upsamples:
                cur = self._mix(branches, 1.0 / nk, LRELU_SLOPE)
            else:
                audio = self._post(branches, 1.0 / nk, POST_SLOPE, prep["post_w"], prep["post_b"])
        return audio

    def __call__(self, x):
        """x: (B, num_mels, T) (or (num_mels, T)) -> (B, 1, samples) like Generator.forward."""
        if x.dim() == 2:
            x = x[None]
        _lib.require_cuda(x, "mel")
        if x.shape[1] != self.num_mels:
            raise ValueError(f"expected (B, {self.num_mels}, T), got {tuple(x.shape)}")
        if x.device != self.device:
            self.to(x.device)
        with torch.no_grad(), torch.cuda.device(self.device):
            outs = [self.forward_one(x[b].to(_f32).t().contiguous()) for b in range(x.shape[0])]
        return torch.stack(outs)[:, None, :]


class Vocoder(object):
    """vocoder.py:16-36 with the generator on libssb: Vocoder(device)(mel (seq_len, 80)) -> 1-D audio.
    checkpoint_file defaults to the reference's --hifigan_checkpoint flag; config.json is read from the
    checkpoint's directory exactly as the reference does."""

    def __init__(self, device="cuda", checkpoint_file=None):
        if checkpoint_file is None:
            from absl import flags
            checkpoint_file = flags.FLAGS.hifigan_checkpoint
        assert checkpoint_file is not None
        config_file = os.path.join(os.path.split(checkpoint_file)[0], "config.json")
        with open(config_file) as f:
            hparams = AttrDict(json.load(f))
        self.generator = Generator(hparams).to(device)
        self.generator.load_state_dict(torch.load(checkpoint_file, map_location="cpu")["generator"])
        self.generator.eval()
        self.generator.remove_weight_norm()

    def __call__(self, mel_spectrogram):
        """mel_spectrogram: (seq_len, 80) tensor; returns a 1-D audio tensor (vocoder.py:28-36)."""
        with torch.no_grad():
            audio = self.generator(mel_spectrogram.T[None, :, :])
        return audio.squeeze()


def samples_per_frame(h):
    return int(math.prod(h["upsample_rates"]))

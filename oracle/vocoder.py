"""ORACLE — test infrastructure, not product code.

Plain torch (CPU, fp32 or fp64) restatement of the HiFi-GAN generator the reference uses as its
vocoder, written against a reference-shaped state_dict.  Follows (paths relative to the reference):
  Generator.__init__ / forward      hifi_gan/models.py:75-112
  ResBlock1 / ResBlock2              hifi_gan/models.py:11-72
  get_padding                        hifi_gan/utils.py:36-37
  Vocoder.__call__                   vocoder.py:28-36   ((seq_len, 80) mel -> 1-D audio)
  weight normalisation               torch.nn.utils.weight_norm (third-party PyTorch, pinned 2.0 in
                                     environment.yml:10): w = g * v / ||v||, the norm taken over all
                                     dimensions but the first - restated in fold_weight_norm
The arithmetic of the convolutions lives in PyTorch (F.conv1d / F.conv_transpose1d); this file
restates the generator's structure on top of those primitives.
Pinned by tests/golden/vocoder_golden.npz: audio produced by the executed reference Generator
(tests/golden/make_golden_vocoder.py) from formula-defined weights, with and without weight norm.
"""
import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1          # hifi_gan/models.py:8
POST_SLOPE = 0.01          # F.leaky_relu default, models.py:109

V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
          upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
          resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])          # hifi_gan/config_v1.json
V3 = dict(resblock="2", upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8],
          upsample_initial_channel=256, resblock_kernel_sizes=[3, 5, 7],
          resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]])                  # hifi_gan/config_v3.json

# name -> (config, mel frames, weight gain) of the cases tests/golden/vocoder_golden.npz holds: small
# generators of both residual-block kinds, widths below one tensor-core k-block on purpose (the GPU
# path pads them); at gain 1 the audio neither vanishes nor saturates the tanh (rms 0.14 - 0.19)
GOLDEN_CASES = {
    "rb1": (dict(V1, upsample_initial_channel=64), 9, 1.0),
    "rb2": (dict(V3, upsample_initial_channel=32), 7, 1.0),
    "rb1_wide": (dict(V1, upsample_initial_channel=128), 5, 1.0),
}


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def shapes(cfg, num_mels=80):
    """name -> shape of the generator's parameters after remove_weight_norm()."""
    c0 = cfg["upsample_initial_channel"]
    out = {"conv_pre.weight": (c0, num_mels, 7), "conv_pre.bias": (c0,)}
    ch = c0
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
        out[f"ups.{i}.weight"] = (cin, ch, k)
        out[f"ups.{i}.bias"] = (ch,)
        for j, (rk, dil) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            for m in range(len(dil)):
                names = (f"{p}.convs1.{m}", f"{p}.convs2.{m}") if cfg["resblock"] == "1" else (f"{p}.convs.{m}",)
                for n in names:
                    out[n + ".weight"] = (ch, ch, rk)
                    out[n + ".bias"] = (ch,)
    out["conv_post.weight"] = (1, ch, 7)
    out["conv_post.bias"] = (1,)
    return out


def _hash_uniform(numel, salt):
    """Portable pseudo-random U(-1, 1): murmur3's 32-bit finaliser over the element index, exact
    integer arithmetic (no libm, no torch generator) so every box regenerates the same weights."""
    h = (torch.arange(numel, dtype=torch.int64) * 2654435761 + salt * 0x9E3779B1) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return h.to(torch.float64) / 2147483648.0 - 1.0


def formula_state_dict(cfg, dtype=torch.float32, gain=1.0, res_gain=0.5):
    """Deterministic weights: hashed U(-a, a) per tensor with std = gain / sqrt(fan_in) on the trunk
    (conv_pre, ups, conv_post) so that activations stay O(1) through the stack (the reference's
    N(0, 0.01) init would drive every intermediate towards 0) and res_gain / sqrt(fan_in) inside the
    residual blocks.  Unstructured on purpose: plane-wave weights (sin of the flat index) make every
    filter a sinusoid whose responses cancel to a small residual, and the generator then amplifies a
    1e-7 perturbation a thousandfold (fp32 itself sat 2e-4 from fp64)."""
    sd = {}
    for n, (name, shp) in enumerate(sorted(shapes(cfg).items())):
        numel = int(math.prod(shp))
        u = _hash_uniform(numel, n + 1)
        if name.endswith(".bias"):
            w = 0.05 * u
        else:
            if name.startswith("ups."):
                fan_in = shp[0] * shp[2] / max(1, cfg["upsample_rates"][int(name.split(".")[1])])
            else:
                fan_in = shp[1] * shp[2]
            w = (res_gain if name.startswith("resblocks.") else gain) * math.sqrt(3.0 / fan_in) * u
        sd[name] = w.reshape(shp).to(dtype)
    return sd


def formula_mel(T, num_mels=80, dtype=torch.float32):
    """(T, num_mels) normalised-mel-like input."""
    i = torch.arange(T * num_mels, dtype=torch.float64)
    return (1.5 * torch.sin(0.173 * i) + 0.5 * torch.sin(0.0031 * i * i)).reshape(T, num_mels).to(dtype)


def fold_weight_norm(g, v):
    """torch.nn.utils.weight_norm(dim=0): w = v * (g / ||v||), norm over every dim but 0 (ATen's
    _weight_norm evaluates it in this order; the deep generator amplifies a different rounding of
    the weights to ~5e-5 in the audio)."""
    dims = tuple(range(1, v.dim()))
    return v * (g / torch.linalg.vector_norm(v, 2, dims, keepdim=True))


def weight_normed_state_dict(sd):
    """The same generator as a checkpoint holds it (weight_g / weight_v, vocoder.py:24): v = the
    weight scaled per output row (which the normalisation undoes), g = the weight's row norm
    modulated by a few percent, so that v * g / ||v|| is NOT the given weight itself."""
    wn = {}
    for k, t in sd.items():
        if k.endswith(".weight"):
            rows = torch.arange(t.shape[0], dtype=t.dtype).view(-1, *([1] * (t.dim() - 1)))
            wn[k[:-7] + ".weight_v"] = t * (1.0 + 0.25 * torch.cos(rows))
            wn[k[:-7] + ".weight_g"] = t.pow(2).sum(tuple(range(1, t.dim())), keepdim=True).sqrt() * \
                (1.0 + 0.04 * torch.sin(rows))
        else:
            wn[k] = t
    return wn


def fold_state_dict(sd):
    """A checkpoint saved with weight norm in place (weight_g / weight_v, as vocoder.py:24 loads it
    before remove_weight_norm) -> plain weights."""
    out = {}
    for k, t in sd.items():
        if k.endswith(".weight_g"):
            base = k[:-len(".weight_g")]
            out[base + ".weight"] = fold_weight_norm(t, sd[base + ".weight_v"])
        elif not k.endswith(".weight_v"):
            out[k] = t
    return out


def generator_forward(sd, mel, cfg):
    """mel: (T, num_mels) -> audio (T * prod(rates),), hifi_gan/models.py:96-112 behind vocoder.py:32-36."""
    x = mel.T[None]                                                   # vocoder.py:33
    x = F.conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j, (rk, dil) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            y = x
            for m, d in enumerate(dil):
                if cfg["resblock"] == "1":
                    t = F.leaky_relu(y, LRELU_SLOPE)
                    t = F.conv1d(t, sd[f"{p}.convs1.{m}.weight"], sd[f"{p}.convs1.{m}.bias"],
                                 dilation=d, padding=get_padding(rk, d))
                    t = F.leaky_relu(t, LRELU_SLOPE)
                    t = F.conv1d(t, sd[f"{p}.convs2.{m}.weight"], sd[f"{p}.convs2.{m}.bias"],
                                 padding=get_padding(rk, 1))
                else:
                    t = F.leaky_relu(y, LRELU_SLOPE)
                    t = F.conv1d(t, sd[f"{p}.convs.{m}.weight"], sd[f"{p}.convs.{m}.bias"],
                                 dilation=d, padding=get_padding(rk, d))
                y = t + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x).squeeze()

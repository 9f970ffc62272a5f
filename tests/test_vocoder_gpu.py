"""HiFi-GAN vocoder on libssb against the oracle, the reference-generated
fixtures and - when the reference copy travelled (baseline/_ref) - the reference's own Generator
run on the CPU of the GPU box.  All through the C ABI (ssb_gemm_tc_kmajor with dilated / phase
im2col views, ssb_voc_mix, ssb_voc_post)."""
import os

import numpy as np
import pytest
import torch

from oracle import vocoder as ov

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vocoder_golden.npz")
# bf16x3 GEMMs keep 16 mantissa bits per operand: ~4e-6 per convolution, 1e-5 after the stack on
# CPU emulation of the same arithmetic (tests/test_vocoder_cpu.py); north star bar: 1e-3
TOL = 1e-4


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def _gen(cfg, sd):
    from silent_speech_b200 import vocoder as sv
    g = sv.Generator(cfg).to("cuda")
    g.load_state_dict(sd)
    return g


@pytest.mark.parametrize("name", list(ov.GOLDEN_CASES))
def test_generator_matches_reference_golden(name):
    cfg, T, gain = ov.GOLDEN_CASES[name]
    g = _gen(cfg, ov.formula_state_dict(cfg, gain=gain))
    audio = g(ov.formula_mel(T).t()[None].cuda())
    assert audio.shape == (1, 1, T * int(np.prod(cfg["upsample_rates"])))
    ref = torch.from_numpy(np.load(GOLD)[f"{name}_audio"])
    err = _rel(audio[0, 0], ref)
    print(f"vocoder {name}: rel-L2 {err:.2e}")
    assert err < TOL


def test_weight_normed_checkpoint():
    cfg, T, gain = ov.GOLDEN_CASES["rb1"]
    g = _gen(cfg, ov.weight_normed_state_dict(ov.formula_state_dict(cfg, gain=gain)))
    audio = g(ov.formula_mel(T).t()[None].cuda())[0, 0]
    assert _rel(audio, torch.from_numpy(np.load(GOLD)["rb1_wn_audio"])) < TOL


@pytest.mark.parametrize("cfg_name,T", [("V1", 40), ("V3", 33), ("V1", 257)])
def test_shipped_configs_against_fp64_oracle(cfg_name, T):
    """config_v1 / config_v3 at full width against the oracle in fp64, beside the fp32 oracle's own
    distance from it (the floor of any fp32 implementation)."""
    cfg = getattr(ov, cfg_name)
    sd64 = ov.formula_state_dict(cfg, dtype=torch.float64)
    mel64 = ov.formula_mel(T, dtype=torch.float64)
    ref = ov.generator_forward(sd64, mel64, cfg)
    sd = {k: v.float() for k, v in sd64.items()}
    floor = _rel(ov.generator_forward(sd, mel64.float(), cfg), ref)
    g = _gen(cfg, sd)
    audio = g(mel64.float().t()[None].cuda())[0, 0]
    err = _rel(audio, ref)
    print(f"vocoder {cfg_name} T={T}: rel-L2 vs fp64 {err:.2e} (fp32 CPU oracle {floor:.2e})")
    assert audio.shape == ref.shape
    assert err < TOL


def test_odd_geometry_and_batch():
    """k != 2u, an odd frame count, four residual branches; a batch of two mels."""
    cfg = dict(resblock="2", upsample_rates=[4, 3], upsample_kernel_sizes=[10, 7],
               upsample_initial_channel=24, resblock_kernel_sizes=[3, 5, 3, 7],
               resblock_dilation_sizes=[[1, 2], [2, 3], [1, 1], [3, 1]])
    sd = ov.formula_state_dict(cfg)
    g = _gen(cfg, sd)
    mels = torch.stack([ov.formula_mel(11), ov.formula_mel(11).flip(0)])
    audio = g(mels.transpose(1, 2).cuda())
    for b in range(2):
        assert _rel(audio[b, 0], ov.generator_forward(sd, mels[b], cfg)) < TOL


def test_vocoder_class_against_the_reference_itself(tmp_path):
    """The reference's Vocoder (vocoder.py) on CPU and the drop-in Vocoder on the B200 from the same
    checkpoint file and config.json; skipped when the reference copy did not travel."""
    import json
    from baseline import refenv
    if refenv.reference_dir() is None:
        pytest.skip("reference copy (baseline/_ref) not present")
    cfg, T = dict(ov.V1, upsample_initial_channel=256), 50
    sd = ov.weight_normed_state_dict(ov.formula_state_dict(cfg))
    with open(tmp_path / "config.json", "w") as f:
        json.dump(cfg, f)
    ckpt = str(tmp_path / "g_00000001")
    torch.save({"generator": sd}, ckpt)
    models, env = refenv.import_reference("models", "env")
    gen = models.Generator(env.AttrDict(cfg))
    gen.load_state_dict(torch.load(ckpt)["generator"])
    gen.eval()
    gen.remove_weight_norm()
    mel = ov.formula_mel(T)
    with torch.no_grad():
        ref = gen(mel.T[np.newaxis, :, :]).squeeze()               # vocoder.py:32-36
    from silent_speech_b200.vocoder import Vocoder
    audio = Vocoder("cuda", checkpoint_file=ckpt)(mel.cuda())
    assert audio.dim() == 1 and audio.is_cuda
    err = _rel(audio, ref)
    print(f"Vocoder vs reference Generator (CPU): rel-L2 {err:.2e}")
    assert err < TOL

"""Timing of HiFi-GAN generator inference (vocoder.py:28-36 / hifi_gan/models.py:96-112, config_v1) for
one utterance: libssb (silent_speech_b200/vocoder.py) vs the same generator on stock PyTorch on the
same GPU (fp32 and TF32; oracle/vocoder.py's F.conv1d / F.conv_transpose1d formulation = the reference's
modules) and on the host cores.  Not the headline bench (bench.py); side metric of SURVEY.md 8 f4.
Usage: python tools/vocoder_bench.py [frames] [--no-cpu] [--kernels]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vocoder as ov  # noqa: E402  (checker + the stock-torch legs only)
from silent_speech_b200 import vocoder as sv  # noqa: E402


def gpu_time(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flops(cfg, T):
    """2 * MACs of the generator (conv_pre, ups, residual blocks, conv_post)."""
    c0 = cfg["upsample_initial_channel"]
    f = 2.0 * T * 80 * c0 * 7
    L, ch = T, c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
        f += 2.0 * L * cin * ch * k
        L *= u
        per = 2 if cfg["resblock"] == "1" else 1
        for rk, dil in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
            f += 2.0 * L * ch * ch * rk * per * len(dil)
    return f + 2.0 * L * ch * 7


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    T = int(args[0]) if args else 600
    cfg = ov.V1
    sd = ov.formula_state_dict(cfg)
    mel = ov.formula_mel(T)
    g = sv.Generator(cfg).to("cuda")
    g.load_state_dict(sd)
    melc = mel.cuda()
    x = melc.t()[None].contiguous()
    audio = g(x)[0, 0]
    sd_c = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cudnn.allow_tf32 = False            # cuDNN convolutions default to TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref = ov.generator_forward(sd_c, melc, cfg)
        ref64 = ov.generator_forward({k: v.double() for k, v in sd_c.items()}, melc.double(), cfg)
    err = float((audio - ref).norm() / ref.norm())
    err64 = float((audio.double() - ref64).norm() / ref64.norm())
    floor64 = float((ref.double() - ref64).norm() / ref64.norm())
    from silent_speech_b200 import _lib
    n0 = _lib.launch_count
    g(x)
    launches = _lib.launch_count - n0
    ms = gpu_time(lambda: g(x))
    # the same forward as one CUDA graph (fixed frame count): no Python between the kernels
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        g(x)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            ag = g(x)
    torch.cuda.synchronize()
    ms_graph = gpu_time(graph.replay)
    gerr = float((ag[0, 0] - ref).norm() / ref.norm())
    res = {"config": "hifi_gan/config_v1.json", "frames": T, "samples": int(audio.numel()),
           "gflop": flops(cfg, T) / 1e9, "rel_l2_vs_torch_fp32_gpu": err, "rel_l2_vs_torch_fp64_gpu": err64,
           "torch_fp32_rel_l2_vs_fp64": floor64, "rel_l2_graph": gerr,
           "libssb_ms": ms, "libssb_graph_ms": ms_graph, "libssb_launches": launches,
           "libssb_graph_tflops": flops(cfg, T) / ms_graph / 1e9,
           "libssb_graph_samples_per_s": audio.numel() / ms_graph * 1e3}
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        res["torch_fp32_gpu_ms"] = gpu_time(lambda: ov.generator_forward(sd_c, melc, cfg))
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        res["torch_tf32_gpu_ms"] = gpu_time(lambda: ov.generator_forward(sd_c, melc, cfg))
        tf = ov.generator_forward(sd_c, melc, cfg)
        res["torch_tf32_rel_l2_vs_fp64"] = float((tf.double() - ref64).norm() / ref64.norm())
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        if "--no-cpu" not in sys.argv:
            Tc = min(T, 100)
            mc = ov.formula_mel(Tc)
            ov.generator_forward(sd, mc, cfg)
            t0 = time.perf_counter()
            ov.generator_forward(sd, mc, cfg)
            dt = time.perf_counter() - t0
            res["cpu_frames"] = Tc
            res["cpu_threads"] = torch.get_num_threads()
            res["cpu_ms_scaled_to_frames"] = 1e3 * dt * T / Tc
    if "--kernels" in sys.argv:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            graph.replay()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        agg = {}
        for e in evs:
            k = e.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "").split("(")[0][:60]
            a = agg.setdefault(k, [0.0, 0])
            a[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
            a[1] += 1
        evs.sort(key=lambda e: e.time_range.start)
        res["seq_us"] = [[e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:24],
                          round(e.time_range.end - e.time_range.start, 1)] for e in evs]
        res["kernels_us"] = {k: [round(v[0], 1), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

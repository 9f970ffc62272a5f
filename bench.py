"""Benchmark of the B200-native silent_speech transduction hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one training step of the reference's transduction model on BASELINE.json's
configs[1]: d_model 768, 6 encoder layers, batch 32 utterances x 4000 raw EMG samples x 8
channels per GPU (T = 500 frames; 16 silent utterances with 600-frame targets aligned by DTW,
16 voiced), forward + dtw_loss + backward + AdamW — transduction_model.py:196-212.
`value` = bs-32 steps per second summed over the N GPUs (weak scaling: every GPU runs its own
32 utterances; gradients are all-reduced over NCCL), inputs resident in HBM; `e2e` = the same
through silent_speech_b200.training.train_step with HOST (pinned) batches, H2D copies and the
loss read-back inside the timed region.  Rank 0 prints ONE JSON line.

`--impl reference`: the reference's own CPU implementation of the step is Python and cannot
travel to the GPU box, so this arm times its CPU port (oracle/, torch fp32 on all host cores)
on a bounded sample of the same workload.
"""
import argparse
import contextlib
import io
import json
import os
import random
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "transduction steps/sec (seq_len=4000, bs=32)"
D_MODEL, N_LAYERS, BS, SEQ, FRAMES = 768, 6, 32, 4000, 500
STEP_GFLOP = 6229.0       # algorithmic GFLOP per bs-32 step per GPU, SURVEY.md §8(d)


def cpu_threads():
    """Threads for the CPU arms: physical cores (logical / 2 with SMT; oversubscribing the 128
    logical CPUs of this pool's hosts made the torch CPU step 20x slower)."""
    n = int(os.environ.get("SSB_CPU_THREADS", "0"))
    return n if n > 0 else max(1, (os.cpu_count() or 2) // 2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_IB_DISABLE", "1")     # single node: NVLink / NVSwitch only
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def timed(fn, steps, warmup, world):
    """fn() is one step; CUDA-event timing, barrier + synchronize on both sides, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def build_model():
    from absl import flags
    from silent_speech_b200 import architecture
    F = flags.FLAGS
    if not F.is_parsed():
        F([sys.argv[0]])
    F.model_size, F.num_layers, F.dropout = D_MODEL, N_LAYERS, 0.2
    torch.manual_seed(0)
    return architecture.Model(112, 80, 48).cuda().train()


def batch_bytes(batch):
    n = 0
    for k in ('raw_emg', 'audio_features', 'phonemes'):
        n += sum(t.numel() * t.element_size() for t in batch[k])
    return n


def gemm_roofline(peaks):
    """Dominant kernel = the tcgen05 GEMM (gemm_tc_kernel).  Time the largest forward GEMM of
    the step alone (FFN1: 16000 x 768 -> 3072, bias + ReLU epilogue) with CUDA events on the
    launching stream, L2 flushed between launches.  `achieved` counts ALGORITHMIC flops
    (2*M*N*K); the kernel issues 3 bf16 MMAs per logical MMA (hi/lo split for fp32-class
    accuracy), so the tensor pipe itself runs at 3x that rate (`tensor_pipe_frac`)."""
    from silent_speech_b200 import functional as SF
    M, K, N = BS * FRAMES, D_MODEL, 3072
    x = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    xp, wp = SF.split_planes(x), SF.split_planes(W)
    op = SF.tc_operand_plain(xp, M, K)
    ep = SF._epi(SF._scatter_plain(y.data_ptr(), M, N), bias=b, relu=1)
    scratch = torch.empty(64 * 1024 * 1024, device="cuda")   # 256 MB > L2: flushed between launches

    def launch():
        SF.gemm_tc_kmajor(op, wp, N, K, ep)
    for _ in range(3):
        launch()
    times = []
    for _ in range(5):
        scratch.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    tf = 2.0 * M * N * K / ms / 1e9
    peak = peaks.get("bf16_tflops", 1590.0)
    return {"bound": "tensor", "kernel": "gemm_tc_kernel bf16x3 (FFN1 16000x768x3072)",
            "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full capture
            # committed as profiles/r1_gemm_tc_persistent.txt (algorithmic: 58 MB in + 197 MB out)
            "traffic": 201.5e6, "traffic_unit": "B/launch",
            "mma_per_logical_mma": 3, "tensor_pipe_frac": 3.0 * tf / peak,
            "peak_source": f"{peaks['_source']} cuBLAS bf16 burst (this kernel is timed alone)",
            "ms_per_launch": ms}


def dtw_side_metric(peaks, world, rank, with_cpu):
    """cfg-2: DTW on 10k (500 x 600) cdist matrices; HBM roofline at 4 B/cell."""
    from silent_speech_b200 import align
    P, Tp, Tg = 10000, 500, 600
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    cost = torch.empty(P, Tp, Tg, device="cuda")
    for s in range(0, P, 1000):
        cost[s:s + 1000] = torch.cdist(torch.randn(1000, Tp, 80, device="cuda", generator=g),
                                       torch.randn(1000, Tg, 80, device="cuda", generator=g))
    view = cost.transpose(1, 2)
    ms = timed(lambda: align.align_batch(view), 5, 3, world) / 5
    cells = P * Tp * Tg * world
    gbs = 4.0 * P * Tp * Tg / ms / 1e6
    out = {"metric": "DTW Mcells/s (10k pairs 500x600 per GPU)", "value": cells / ms / 1e3,
           "ms": ms, "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                  "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                  "peak_source": peaks["_source"],
                                  "algorithmic_bytes_per_cell": 4}}
    if with_cpu:   # CPU port of align.py (oracle C, OpenMP over pairs) on a bounded sample
        from oracle import dtw as odtw
        n = 512
        host = cost[:n].cpu().numpy().transpose(0, 2, 1)
        odtw.align_batch(host[:8])
        t0 = time.perf_counter()
        odtw.align_batch(host, threads=cpu_threads())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n * Tp * Tg / dt / 1e6, "unit": "Mcells/s",
                               "cores": cpu_threads(), "kind": "port",
                               "sample": f"{n} of the 10000 pairs"}
    del cost
    return out


def mel_side_metric(peaks, with_cpu):
    """log-mel of 32 x 10 s clips at 22.05 kHz; 1344 B/frame algorithmic."""
    from silent_speech_b200 import data_utils as du
    y = (torch.rand(32, 220500, device="cuda") * 2 - 1) * 0.5
    f = lambda: du.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
    ms = timed(f, 10, 3, 1) / 10       # public call: includes the reference's range check (a host read)
    frames = 32 * 861
    # the kernel alone through the C ABI (what the roofline is about)
    from silent_speech_b200 import _lib
    lib = _lib.load()
    basis, begin, end = du._basis_for(22050, 1024, 80, 0, 8000, y.device)
    out_k = torch.empty(32, 80, 861, device="cuda")

    def k():
        _lib.check(lib.ssb_mel_fwd(y.data_ptr(), 32, 220500, y.stride(0), 1024, 256, 1024,
                                   basis.data_ptr(), begin.data_ptr(), end.data_ptr(), 80, 1e-5,
                                   out_k.data_ptr(), _lib.current_stream()))
    ms_k = timed(k, 20, 3, 1) / 20
    gbs = 1344.0 * frames / ms_k / 1e6
    out = {"metric": "mel kframes/s (32 clips x 10 s)", "value": frames / ms, "ms": ms,
           "kernel_ms": ms_k, "kernel_kframes_per_s": frames / ms_k,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_frame": 1344,
                        "note": "kernel time; fp32 16x16x4 register FFT + sparse mel: compute / "
                                "shared-memory bound at this size (37 MB of algorithmic traffic), "
                                "not HBM-bound"}}
    if with_cpu:
        from oracle import mel as omel
        yh = y[:4].cpu().numpy()
        t0 = time.perf_counter()
        omel.mel_spectrogram(yh)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 4 * 861 / dt / 1e3, "unit": "kframes/s", "cores": 1,
                               "kind": "port", "sample": "4 of the 32 clips (numpy)"}
    return out


def cpu_port_step_time(n_utt, steps, warmup):
    """The reference step's CPU port (oracle/) on `n_utt` of the 32 utterances, all host cores."""
    from oracle import model as om
    from oracle import step as ostep
    from silent_speech_b200.read_emg import synthetic_batch
    torch.manual_seed(0)
    sd = om.formula_state_dict(D_MODEL, N_LAYERS)
    params = ostep.make_params(sd)
    optim = ostep.make_optimizer(params)
    batch = synthetic_batch(n_utt, FRAMES, seed=1234)
    batch = {k: ([t.clone() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else v)
             for k, v in batch.items()}
    random.seed(0)
    for _ in range(warmup):
        ostep.train_step(params, optim, batch, FRAMES, dropout_p=0.2)
    t0 = time.perf_counter()
    for _ in range(steps):
        ostep.train_step(params, optim, batch, FRAMES, dropout_p=0.2)
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 per rank; the CPU arm uses every physical host core
    torch.set_num_threads(cpu_threads())
    n_utt = 4
    sec = cpu_port_step_time(n_utt, max(1, min(args.steps, 2)), 1 if args.warmup else 0)
    value = (n_utt / BS) / sec
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3 * BS / n_utt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg-1 step (768/6, seq_len 4000) on the CPU port, sample = "
                                   f"{n_utt} of {BS} utterances per step, scaled to bs-32 steps/s"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n_utt}/{BS} utterances x {max(1, min(args.steps, 2))} step(s)"},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    from silent_speech_b200 import _lib
    from silent_speech_b200.read_emg import synthetic_batch
    from silent_speech_b200.training import GradientBucket, GraphedTrainStep, train_step
    world, rank, local = dist_setup(args.gpus)
    peaks = load_peaks()
    _lib.load()   # fails loudly if the CUDA library is missing

    from silent_speech_b200.optim import FlatAdamW
    model = build_model()
    bucket = GradientBucket(model)
    # transduction_model.py:178: AdamW(weight_decay=1e-7), here as one fused pass over the flat
    # parameter / gradient buckets (csrc/optim.cu)
    optim = FlatAdamW(bucket, lr=1e-3, weight_decay=1e-7)
    host_batch = synthetic_batch(BS, FRAMES, seed=1234 + rank)      # pinned host tensors
    dev_batch = dict(host_batch)
    for k in ('raw_emg', 'audio_features', 'phonemes'):
        dev_batch[k] = [t.cuda() for t in host_batch[k]]
    random.seed(rank)
    torch.manual_seed(1000 + rank)

    # the public training call: eager train_step, or (default) the same step with
    # zero_grad + forward + dtw_loss + backward replayed as one CUDA graph per batch signature
    graphed = None if args.eager else GraphedTrainStep(model, optim, "cuda", FRAMES, bucket)

    def step_device():
        # combine_fixed_length copies, so the in-place augmentation never touches dev_batch
        if graphed is not None:
            graphed(dev_batch, sync_loss=False)
        else:
            train_step(model, optim, dev_batch, "cuda", FRAMES, bucket, sync_loss=False)

    def step_e2e():
        if graphed is not None:
            graphed(host_batch, sync_loss=True)
        else:
            train_step(model, optim, host_batch, "cuda", FRAMES, bucket, sync_loss=True)

    with ClockSampler(local) as clk:
        ms = timed(step_device, args.steps, args.warmup, world)
        l0 = _lib.launch_count          # libssb kernels of one steady-state step (graph replays
        step_device()                   # count the kernels captured in the graph)
        launches = _lib.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2), world)
    # host-side enqueue time per step (no synchronisation inside): shows whether the step is
    # bounded by the GPU or by launching ~1000 kernels from Python
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    host_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    torch.cuda.synchronize()

    graph_ms = None
    if args.graph_probe and graphed is None:
        # EXPERIMENT (not a reported number): replay the step as one CUDA graph to see the
        # GPU-only time.  Seeds / augmentation are frozen inside the graph, so this is not a
        # valid training loop.
        optim_c = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-7, fused=True,
                                    capturable=True)
        sidestream = torch.cuda.Stream()
        sidestream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sidestream):
            for _ in range(3):
                train_step(model, optim_c, dev_batch, "cuda", FRAMES, bucket, sync_loss=False)
        torch.cuda.current_stream().wait_stream(sidestream)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            train_step(model, optim_c, dev_batch, "cuda", FRAMES, bucket, sync_loss=False)
        graph_ms = timed(g.replay, args.steps, 2, world) / args.steps

    value = world * args.steps / (ms / 1e3)
    e2e = world * args.steps / (ms_e2e / 1e3)
    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "cfg-1: d_model 768, 6 layers, bs 32/GPU, seq_len 4000 (T=500), "
                                   "16 silent (DTW, 600-frame targets) + 16 voiced synthetic "
                                   "utterances; fwd + dtw_loss + bwd + grad all-reduce + AdamW",
                       "dropout": 0.2, "parallelism": f"dp{world}",
                       "launch": "eager" if graphed is None else
                                 "CUDA graph (zero_grad+fwd+loss+bwd) + eager all-reduce + fused flat AdamW",
                       "l2": "per-step working set (~10 GB of activations) >> 126 MB L2"},
            "e2e": {"value": e2e, "unit": "steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": batch_bytes(host_batch), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms,
            "graph_probe_ms_per_step": graph_ms,
            "step_tflops": STEP_GFLOP * value / world / 1e3,
            "clocks": clk.summary()}
    if rank == 0:
        line["roofline"] = gemm_roofline(peaks)
    if not args.no_side:
        side = dtw_side_metric(peaks, world, rank, rank == 0 and world == 1 and not args.no_cpu)
        if rank == 0:
            line["dtw"] = side
        if rank == 0 and world == 1:
            line["mel"] = mel_side_metric(peaks, not args.no_cpu)
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(cpu_threads())
        n_utt = 4
        sec = cpu_port_step_time(n_utt, 1, 0)
        line["cpu_baseline"] = {"value": (n_utt / BS) / sec, "unit": "steps/s",
                                "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n_utt}/{BS} utterances x 1 step, scaled to bs-32 steps/s"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-side", action="store_true", help="skip the DTW side metric")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--eager", action="store_true",
                    help="launch every kernel from Python instead of replaying the CUDA graph")
    ap.add_argument("--graph-probe", action="store_true",
                    help="experiment (with --eager): also time a whole-step CUDA graph")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that write to fd 1 on their own
    # (NCCL prints its version banner there) go to stderr while the run is in progress
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    for l in lines[:-1]:
        print(l, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()

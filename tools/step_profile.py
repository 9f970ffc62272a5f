"""Per-kernel time of ONE steady-state cfg-1 training step, measured in process with CUPTI
(torch.profiler) instead of ncu: kernels run back to back as they do in the bench (no per-kernel
serialisation / cache flush), and a run costs seconds.  Not a bench number; the ncu launch list
under profiles/ stays the committed evidence.

    python tools/step_profile.py [--eager] [--top 40] [--seq FILE]
"""
import argparse
import os
import random
import re
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")
    name = re.sub(r"\(.*", "", name)
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--seq", default=None, help="write the kernel sequence (name, us) to this file")
    ap.add_argument("--steady", type=int, default=0,
                    help="profile this many back-to-back steps (no host sync between them) and report "
                         "the LAST one: the host runs ahead as in the bench, so idle time between "
                         "activities is the GPU's own, not the enqueue latency of a cold step")
    args = ap.parse_args()
    from silent_speech_b200.read_emg import synthetic_batch
    from silent_speech_b200.training import GradientBucket, GraphedTrainStep, train_step
    model = bench.build_model()
    from silent_speech_b200.optim import FlatAdamW
    bucket = GradientBucket(model)
    optim = FlatAdamW(bucket, lr=1e-3, weight_decay=1e-7)
    batch = synthetic_batch(bench.BS, bench.FRAMES, seed=1234)
    for k in ('raw_emg', 'audio_features', 'phonemes'):
        batch[k] = [t.cuda() for t in batch[k]]
    random.seed(0)
    graphed = None if args.eager else GraphedTrainStep(model, optim, "cuda", bench.FRAMES, bucket)

    def step():
        if graphed is not None:
            graphed(batch, sync_loss=False)
        else:
            train_step(model, optim, batch, "cuda", bench.FRAMES, bucket, sync_loss=False)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(max(1, args.steady)):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if args.steady > 1:
        # a step ends with the fused optimiser kernel: keep what lies between the last two of them
        ends = [i for i, e in enumerate(evs) if "adamw_flat_kernel" in e.name]
        assert len(ends) >= 2, "no optimiser kernels found"
        prev_end = evs[ends[-2]].time_range.end
        evs = evs[ends[-2] + 1:ends[-1] + 1]
        print(f"# steady state: step {args.steady} of {args.steady} back-to-back steps; "
              f"{(evs[0].time_range.start - prev_end):.1f} us idle after the previous step's optimiser kernel")
    agg, cnt = defaultdict(float), defaultdict(int)
    for e in evs:
        d = e.time_range.end - e.time_range.start
        agg[short(e.name)] += d
        cnt[short(e.name)] += 1
    tot = sum(agg.values())
    span = evs[-1].time_range.end - evs[0].time_range.start
    print(f"# {len(evs)} device activities, busy {tot / 1e3:.2f} ms, span {span / 1e3:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:args.top]:
        print(f"{v:10.1f} us {100 * v / tot:5.1f}% {cnt[k]:5d}  {k}")
    # idle time between consecutive activities, attributed to the activity that FOLLOWS the gap
    # (its launch latency / dependency wait): where the span - busy difference sits
    gap, gcnt = defaultdict(float), defaultdict(int)
    for a, b in zip(evs[:-1], evs[1:]):
        g = b.time_range.start - a.time_range.end
        if g > 0:
            gap[short(b.name)] += g
            gcnt[short(b.name)] += 1
    gtot = sum(gap.values())
    print(f"# idle between activities: {gtot / 1e3:.2f} ms; by following kernel:")
    for k, v in sorted(gap.items(), key=lambda kv: -kv[1])[:15]:
        print(f"#  gap {v:8.1f} us {100 * v / max(gtot, 1e-9):5.1f}% {gcnt[k]:5d} (avg {v / gcnt[k]:5.2f} us)  {k}")
    if args.seq:
        with open(args.seq, "w") as f:
            prev_end = None
            for e in evs:
                g = 0.0 if prev_end is None else e.time_range.start - prev_end
                prev_end = e.time_range.end
                f.write(f"{short(e.name)}\t{e.time_range.end - e.time_range.start:.1f}\t{g:.1f}\n")


if __name__ == "__main__":
    main()

"""Micro-benchmark of the batched DTW path (cfg-2 geometry). CUDA-event timing, inputs
(12 GB at 10k pairs) far larger than L2.  Usage: python tools/dtw_bench.py [npairs] [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import align  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    Tp, Tg = 500, 600
    g = torch.Generator(device="cuda").manual_seed(1234)
    cost = torch.empty(P, Tp, Tg, device="cuda")
    for s in range(0, P, 500):
        e = min(P, s + 500)
        pred = torch.randn(e - s, Tp, 80, device="cuda", generator=g)
        tgt = torch.randn(e - s, Tg, 80, device="cuda", generator=g)
        cost[s:e] = torch.cdist(pred, tgt)
    view = cost.transpose(1, 2)
    for _ in range(3):
        path = align.align_batch(view)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        path = align.align_batch(view)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    best, mean = min(ms), sum(ms) / len(ms)
    cells = P * Tp * Tg
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = {"npairs": P, "ms_mean": mean, "ms_best": best, "mcells_per_s": cells / mean / 1e3,
           "algorithmic_gbs": 4 * cells / mean / 1e6, "hbm_peak_gbs": hbm,
           "frac": 4 * cells / mean / 1e6 / hbm, "all_ms": ms,
           "path_checksum": int(path.sum().item())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Small DTW run for compute-sanitizer: the band pipeline (5 warps per pair, progress words in shared
memory) and the shared-memory backtrace on a few pairs, checked against the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dtw as odtw  # noqa: E402
from silent_speech_b200 import align  # noqa: E402

rs = np.random.RandomState(0)
for P, Tp, Tg in ((6, 500, 600), (3, 130, 257)):
    cost = np.abs(rs.randn(P, Tp, Tg)).astype(np.float32)
    got = align.align_batch(torch.from_numpy(cost).cuda().transpose(1, 2)).cpu().numpy()
    want = odtw.align_batch(cost.transpose(0, 2, 1))
    assert (got == want).all()
torch.cuda.synchronize()
print("dtw small check ok")

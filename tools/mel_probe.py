"""Launch the log-mel kernel on the bench geometry (32 clips x 10 s @ 22.05 kHz) a few times
(for `ncu --set full -k regex:mel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import data_utils as du  # noqa: E402

y = (torch.rand(32, 220500, device="cuda") * 2 - 1) * 0.5
for _ in range(4):
    out = du.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
torch.cuda.synchronize()
print(tuple(out.shape))

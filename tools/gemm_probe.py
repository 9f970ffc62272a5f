"""Launch the step's dominant GEMM shapes alone (for `ncu --set full -k regex:gemm_tc`):
  0: FFN1 forward   16000 x 768 -> 3072, bias + ReLU + dropout, split-plane output only
  1: FFN2 masked data gradient 16000 x 768 -> 3072, mask from the hi plane, split-plane output
  2: FFN2 forward   16000 x 3072 -> 768, bias, fp32 output
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from silent_speech_b200 import functional as SF  # noqa: E402

M, D, F = 16000, 768, 3072
x, w1, w2 = torch.randn(M, D).cuda(), (torch.randn(F, D) * D ** -0.5).cuda(), (torch.randn(D, F) * F ** -0.5).cuda()
b1, b2 = torch.zeros(F).cuda(), torch.zeros(D).cuda()
xp, w1p, w2p = SF.split_planes(x), SF.split_planes(w1), SF.split_planes(w2)
hp = torch.empty(2, M, F, dtype=torch.bfloat16, device="cuda")
dhp = torch.empty_like(hp)
y = torch.empty(M, D, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    SF.gemm_tc_kmajor(SF.tc_operand_plain(xp, M, D), w1p, F, D,
                      SF._epi(SF._scatter_plain(None, M, F), bias=b1, relu=1, drop_p=0.2, seed=1, site=2,
                              planes_out=hp))
    SF.gemm_tc_kmajor(SF.tc_operand_plain(xp, M, D), SF.split_planes_t(w2), F, D,
                      SF._epi(SF._scatter_plain(None, M, F), mask_planes=hp[0], mask_scale=1.25,
                              planes_out=dhp))
    SF.gemm_tc_kmajor(SF.tc_operand_plain(hp, M, F), w2p, D, F,
                      SF._epi(SF._scatter_plain(y.data_ptr(), M, D), bias=b2))
torch.cuda.synchronize()
print("ok")

"""Mirror of the reference's architecture.py on the B200 kernels.

`Model(num_features, num_outs, num_aux_outs=None)` with `forward(x_feat, x_raw, session_ids)`
keeps the reference's signature, flags, parameter / buffer names and shapes
(architecture.py:10-12, 14-84; state_dict contract in SURVEY.md §8b), so the reference's
transduction_model.py / recognition_model.py run unchanged against it and its checkpoints load.

Data flow (channels-last, no NCL transposes): x_raw (B, L, 8) -> 3 x ResBlock as gathered
GEMMs + fused BatchNorm/ReLU/residual kernels -> (B*T, D) tokens -> w_raw_in -> N encoder
layers (transformer.py mirror) -> output heads.
"""
import os
import random

import torch
from torch import nn
from absl import flags

from . import functional as F_
from .transformer import TransformerEncoder, TransformerEncoderLayer
from .weights import WeightPlanes

FLAGS = flags.FLAGS
for _define, _name, _default, _help in (
        (flags.DEFINE_integer, 'model_size', 768, 'number of hidden dimensions'),
        (flags.DEFINE_integer, 'num_layers', 6, 'number of layers'),
        (flags.DEFINE_float, 'dropout', .2, 'dropout')):
    try:
        _define(_name, _default, _help)
    except flags.DuplicateFlagError:   # the reference's own architecture.py is also imported
        pass


def _conv_weight(conv):
    """nn.Conv1d weight (Cout, Cin, k) -> GEMM layout (k*Cin, Cout), row = tap*Cin + ci."""
    w = conv.weight
    return w.permute(2, 1, 0).reshape(w.shape[2] * w.shape[1], w.shape[0]).contiguous()


def _shift_rows_device(x_raw, r_dev):
    """architecture.py:64-68 with the shift r held in device memory (CUDA-graph friendly):
    x_raw[:, :-r] = x_raw[:, r:]; x_raw[:, -r:] = 0, in place; r = 0 is the identity."""
    L = x_raw.shape[1]
    idx = torch.arange(L, device=x_raw.device) + r_dev
    valid = (idx < L).to(x_raw.dtype).view(1, L, 1)
    x_raw.copy_(x_raw.index_select(1, idx.clamp_(max=L - 1)) * valid)


class ResBlock(nn.Module):
    """architecture.py:14-40.  Submodules exist for their parameters/buffers (checkpoint
    contract); forward runs the fused channels-last kernels."""

    def __init__(self, num_ins, num_outs, stride=1):
        super().__init__()
        self.conv1 = nn.Conv1d(num_ins, num_outs, 3, padding=1, stride=stride)
        self.bn1 = nn.BatchNorm1d(num_outs)
        self.conv2 = nn.Conv1d(num_outs, num_outs, 3, padding=1)
        self.bn2 = nn.BatchNorm1d(num_outs)
        if stride != 1 or num_ins != num_outs:
            self.residual_path = nn.Conv1d(num_ins, num_outs, 1, stride=stride)
            self.res_norm = nn.BatchNorm1d(num_outs)
        else:
            self.residual_path = None
        self.stride = stride
        self._bump_counters = True

    def _bn_args(self, bn):
        if self.training and self._bump_counters:
            bn.num_batches_tracked += 1
        return bn.weight, bn.bias, bn.running_mean, bn.running_var

    def batch_counters(self):
        return [bn.num_batches_tracked for bn in (self.bn1, self.bn2, getattr(self, "res_norm", None))
                if bn is not None]

    def forward_cl(self, x, wp=None, bump_counters=True):
        """x: (B, L, Cin) channels-last -> (B, Lout, Cout).  wp: the model's WeightPlanes arena.
        bump_counters=False: the caller has already advanced num_batches_tracked (Model does all
        nine in one launch)."""
        self._bump_counters = bump_counters
        if self.residual_path is None:
            raise NotImplementedError("identity residual is never instantiated by the reference "
                                      "(architecture.py:46-50) and is not built")
        tr = self.training
        # every convolution here feeds a BatchNorm: in training mode its bias gradient is exactly 0
        # conv1 and residual_path read the same input: one autograd node when x needs a gradient
        pair = F_.conv_pair_w(x, self.conv1, self.residual_path, wp, tr) if self.stride == 2 else None
        if pair is not None:
            c1, cr = pair
        else:
            c1 = F_.conv1d_w(x, self.conv1, wp, 3, self.stride, lambda: _conv_weight(self.conv1), tr)
            cr = None
        h1 = F_.bn_act(c1, *self._bn_args(self.bn1), training=tr, relu=True,
                       momentum=self.bn1.momentum, eps=self.bn1.eps)
        c2 = F_.conv1d_w(h1, self.conv2, wp, 3, 1, lambda: _conv_weight(self.conv2), tr)
        if cr is None:
            cr = F_.conv1d_w(x, self.residual_path, wp, 1, self.stride,
                             lambda: _conv_weight(self.residual_path), tr)
        ga, ba, rma, rva = self._bn_args(self.bn2)
        gb, bb, rmb, rvb = self._bn_args(self.res_norm)
        return F_.bn_act(c2, ga, ba, rma, rva, tr, True, cr, gb, bb, rmb, rvb,
                         momentum=self.bn2.momentum, eps=self.bn2.eps)

    def forward(self, x):
        """x: (B, Cin, L) as in the reference -> (B, Cout, Lout)."""
        return self.forward_cl(x.transpose(1, 2).contiguous()).transpose(1, 2)


class Model(nn.Module):
    def __init__(self, num_features, num_outs, num_aux_outs=None):
        super().__init__()
        model_size = FLAGS['model_size'].value
        num_layers = FLAGS['num_layers'].value
        dropout = FLAGS['dropout'].value
        self.conv_blocks = nn.Sequential(
            ResBlock(8, model_size, 2),
            ResBlock(model_size, model_size, 2),
            ResBlock(model_size, model_size, 2),
        )
        self.w_raw_in = nn.Linear(model_size, model_size)
        encoder_layer = TransformerEncoderLayer(d_model=model_size, nhead=8,
                                                relative_positional=True,
                                                relative_positional_distance=100,
                                                dim_feedforward=3072, dropout=dropout)
        self.transformer = TransformerEncoder(encoder_layer, num_layers)
        self.w_out = nn.Linear(model_size, num_outs)
        self.has_aux_out = num_aux_outs is not None
        if self.has_aux_out:
            self.w_aux = nn.Linear(model_size, num_aux_outs)
        # None: the augmentation shift is drawn with Python's `random` per forward, as in the
        # reference.  A 0-d int64 device tensor: the shift is read from it at execution time
        # (training.GraphedTrainStep draws it on the host and fills the cell before each replay).
        self.shift_source = None
        self._wp = None     # WeightPlanes arena, created at the first CUDA forward

    def weight_planes(self, device):
        """Every weight's tensor-core operand layouts, made current by one launch (weights.py).
        SSB_WPLANES=0 falls back to deriving them per use (round-1 path, A/B testing)."""
        if device.type != "cuda" or not F_._tc_enabled() or os.environ.get("SSB_WPLANES", "1") == "0":
            return None
        if self._wp is None or self._wp.model is not self:
            self._wp = WeightPlanes(self)
        self._wp.refresh()
        return self._wp

    def forward(self, x_feat, x_raw, session_ids):
        # x_raw is (batch, time, electrode); x_feat and session_ids are ignored, as in the
        # reference (architecture.py:61)
        if self.training:
            if self.shift_source is not None:
                _shift_rows_device(x_raw, self.shift_source)
            else:
                r = random.randrange(8)       # architecture.py:64-68, mutates the caller's tensor
                if r > 0:
                    x_raw[:, :-r, :] = x_raw[:, r:, :].clone()
                    x_raw[:, -r:, :] = 0
        x = x_raw.to(torch.float32).contiguous()
        wp = self.weight_planes(x.device)
        counters = [c for blk in self.conv_blocks if blk.training for c in blk.batch_counters()]
        if counters:          # the 9 BatchNorm step counters in one launch instead of nine
            torch._foreach_add_(counters, 1)
        for blk in self.conv_blocks:
            x = blk.forward_cl(x, wp, bump_counters=False)
        B, T, D = x.shape
        x2 = F_.linear_w(x.view(B * T, D), self.w_raw_in, wp)
        x2 = self.transformer.forward_tokens(x2, B, T, wp)
        if self.has_aux_out:
            hf = wp.get(self, "heads_f") if wp is not None else None
            M, N = B * T, self.w_out.out_features + self.w_aux.out_features
            if (hf is not None and F_._tc_fwd_ok(M, N, D) and F_._tc_fwd_ok(M, D, N)
                    and F_._tc_wgrad_ok(M, N, D)):
                out, aux = F_._HeadsFn.apply(x2, self.w_out.weight, self.w_out.bias, self.w_aux.weight,
                                             self.w_aux.bias, hf, wp.get(self, "heads_b"))
                return out.view(B, T, -1), aux.view(B, T, -1)
            out = F_.linear_w(x2, self.w_out, wp).view(B, T, -1)
            aux = F_.linear_w(x2, self.w_aux, wp).view(B, T, -1)
            return out, aux
        return F_.linear_w(x2, self.w_out, wp).view(B, T, -1)

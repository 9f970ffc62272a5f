"""The training step of transduction_model.py:196-212 as one reusable call, plus the
data-parallel gradient exchange the north star adds (the reference is single-process).

train_step(model, optim, batch, device, seq_len_frames):
    H2D of the collate_raw-style batch -> combine_fixed_length -> Model forward -> dtw_loss ->
    backward -> [flat-bucket NCCL all-reduce of gradients] -> optimizer step.
"""
import random

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F_
from .data_utils import combine_fixed_length
from .losses import dtw_loss


class GradientBucket:
    """All gradients of a replica in ONE flat fp32 buffer (p.grad are views into it), so the
    data-parallel exchange is a single NCCL all-reduce over NVLink (213 MB at 768/6) instead
    of ~100 small ones.  The relative-position tables never get gradients and
    are left out."""

    def __init__(self, model):
        self.params = [p for n, p in model.named_parameters()
                       if p.requires_grad and not n.endswith("relative_positional.embeddings")]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        F_.register_grad_sinks(self.params)      # backward bodies may accumulate in place

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, optim=None):
        """Sum over ranks, then the mean: as a pass over the bucket, or (optim = FlatAdamW)
        folded into the fused optimiser kernel as its grad_scale."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if hasattr(optim, "grad_scale"):
                optim.grad_scale = 1.0 / dist.get_world_size()
            else:
                self.flat.mul_(1.0 / dist.get_world_size())


def to_device(batch, device):
    """transduction_model.py:200-202: per-tensor non_blocking H2D of the list-valued batch."""
    out = dict(batch)
    for k in ('raw_emg', 'emg', 'session_ids', 'audio_features', 'phonemes'):
        if k in batch:
            out[k] = [t.to(device, non_blocking=True) for t in batch[k]]
    return out


def train_step(model, optim, batch, device, seq_len_frames=200, bucket=None, sync_loss=True):
    """One optimisation step.  `batch` follows EMGDataset.collate_raw (read_emg.py:262-296).
    Returns the loss (python float if sync_loss, else a 0-d device tensor)."""
    if bucket is not None:
        bucket.zero()
    else:
        optim.zero_grad(set_to_none=True)
    b = to_device(batch, device)
    X_raw = combine_fixed_length(b['raw_emg'], seq_len_frames * 8)
    pred, phoneme_pred = model(None, X_raw, None)
    loss, _ = dtw_loss(pred, phoneme_pred, b)
    loss.backward()
    if bucket is not None:
        bucket.allreduce_mean(optim)
    optim.step()
    return loss.item() if sync_loss else loss.detach()


_BATCH_KEYS = ('raw_emg', 'audio_features', 'phonemes')


class GraphedTrainStep:
    """train_step with zero_grad + forward + dtw_loss + backward replayed as ONE CUDA graph.

    An eager step issues ~1000 kernel launches from Python (~36 ms of host time at cfg-1, as
    long as the GPU work itself); the graph replays them in one driver call.  What makes a
    captured training step valid rather than a frozen one:
      * dropout: every site keys Philox with seed + *seed_cell (ssb_set_seed_source); the cell is
        bumped by the first node of the graph, so each replay draws fresh masks and the
        forward / backward of one replay agree;
      * the train-mode shift augmentation (architecture.py:64-68) is drawn on the host with
        `random.randrange(8)` exactly like the reference and handed over in a device cell;
      * inputs are copied into static device buffers (H2D from the caller's pinned tensors, or
        D2D) before each replay; BatchNorm running statistics and `num_batches_tracked` are
        updated by the replayed kernels in place.
    The gradient all-reduce and `optim.step()` stay outside the graph (a handful of launches), so
    any optimizer / LR schedule works unchanged (transduction_model.py:178-189).

    One graph per batch signature (utterance lengths, silent flags, train/eval): the first step
    of a new signature runs eagerly (lazy initialisation must not happen under capture), the
    second captures, later ones replay.
    """

    def __init__(self, model, optim, device, seq_len_frames=200, bucket=None):
        self.model, self.optim, self.device = model, optim, torch.device(device)
        self.seq_len_frames, self.bucket = seq_len_frames, bucket
        if bucket is None:
            raise ValueError("GraphedTrainStep needs a GradientBucket: gradients must live at "
                             "fixed addresses across replays")
        self.seed_cell = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.shift_cell = torch.zeros((), dtype=torch.int64, device=self.device)
        self._seen = set()
        self._graphs = {}
        self.kernels_per_replay = 0

    @staticmethod
    def signature(batch, training):
        shapes = tuple(tuple(tuple(t.shape) for t in batch[k]) for k in _BATCH_KEYS)
        return (shapes, tuple(bool(s) for s in batch['silent']), tuple(batch['lengths']), training)

    def _body(self, b):
        self.bucket.zero()
        X_raw = combine_fixed_length(b['raw_emg'], self.seq_len_frames * 8)
        pred, phoneme_pred = self.model(None, X_raw, None)
        loss, _ = dtw_loss(pred, phoneme_pred, b)
        loss.backward()
        return loss.detach()

    def _capture(self, batch):
        static = dict(batch)
        for k in _BATCH_KEYS:
            static[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch[k]]
        graph = torch.cuda.CUDAGraph()
        F_.set_seed_source(self.seed_cell)
        self.model.shift_source = self.shift_cell
        n0 = _lib.launch_count
        try:
            with torch.cuda.graph(graph):
                self.seed_cell.add_(0x9E3779B97F4A7C15 - (1 << 64))   # odd 64-bit stride
                loss = self._body(static)
        finally:
            self.model.shift_source = None
            F_.set_seed_source(None)
        return graph, static, loss, _lib.launch_count - n0

    def __call__(self, batch, sync_loss=True):
        sig = self.signature(batch, self.model.training)
        if sig not in self._graphs:
            if sig not in self._seen:      # first sight: a plain eager step
                self._seen.add(sig)
                return train_step(self.model, self.optim, batch, self.device, self.seq_len_frames,
                                  self.bucket, sync_loss)
            self._graphs[sig] = self._capture(batch)
        graph, static, loss, nk = self._graphs[sig]
        for k in _BATCH_KEYS:
            for dst, src in zip(static[k], batch[k]):
                dst.copy_(src, non_blocking=True)
        if self.model.training:
            self.shift_cell.fill_(random.randrange(8))
        graph.replay()
        _lib.launch_count += nk
        self.kernels_per_replay = nk
        self.bucket.allreduce_mean(self.optim)
        self.optim.step()
        return loss.item() if sync_loss else loss

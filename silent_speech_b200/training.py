"""The training steps of the reference as reusable calls, plus the data-parallel gradient exchange
the north star adds (the reference is single-process).

  transduction   transduction_model.py:196-212   train_step / GraphedTrainStep(task="transduction")
      H2D of the collate_raw-style batch -> combine_fixed_length -> Model forward -> dtw_loss ->
      backward -> [flat-bucket NCCL all-reduce of gradients] -> optimizer step.
  recognition    recognition_model.py:89-107     ctc_train_step / GraphedTrainStep(task="recognition")
      same Model with one 38-way head -> fused log-softmax + CTC -> backward, optimiser every
      `accumulate`-th batch (the reference steps every 2nd batch, :104-107).
"""
import random

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F_
from .data_utils import combine_fixed_length
from .losses import ctc_loss_from_chunks, dtw_loss


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_model(model, src=0):
    """Make every replica start from rank `src`'s parameters AND buffers (BatchNorm running
    statistics, num_batches_tracked).  The reference is single-process; under data parallelism
    the replicas otherwise agree only if every rank seeded identically before building the model.
    BatchNorm statistics stay per-rank afterwards (stock-DDP semantics):
    rank 0's buffers are the ones to checkpoint."""
    if _world() == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


class GradientBucket:
    """All gradients of a replica in ONE flat fp32 buffer (p.grad are views into it), so the
    data-parallel exchange is a single NCCL all-reduce over NVLink (213 MB at 768/6) instead
    of ~100 small ones.  The relative-position tables never get gradients and
    are left out.

    The buffer is laid out in named_parameters() order (conv stack, w_raw_in, encoder layers
    0..N-1, heads), i.e. backward completes it from the END towards the start.  `segments`
    cuts it at encoder-layer boundaries into `n_segments` contiguous ranges, last range first
    complete; allreduce_segment(i) exchanges one of them (training.OverlappedAllReduce issues
    them from backward hooks so the exchange overlaps the rest of backward)."""

    def __init__(self, model, n_segments=4):
        named = [(n, p) for n, p in model.named_parameters()
                 if p.requires_grad and not n.endswith("relative_positional.embeddings")]
        self.params = [p for _, p in named]
        self.names = [n for n, _ in named]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = {}
        for name, p in named:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.offsets[name] = off
            off += p.numel()
        F_.register_grad_sinks(self.params)      # backward bodies may accumulate in place
        self.segments = self._make_segments(n_segments)

    def _make_segments(self, n_segments):
        """[(start, end, first_layer)] in COMPLETION order: the tail (last layers + heads) first.
        first_layer = index of the lowest encoder layer whose gradients lie in the range (the
        range is complete once that layer's backward has run); -1 for the conv-stack range."""
        layer_off = {}
        for name in self.names:
            if name.startswith("transformer.layers."):
                l = int(name.split(".")[2])
                layer_off.setdefault(l, self.offsets[name])
        L = len(layer_off)
        n = self.flat.numel()
        if L == 0 or n_segments <= 1:
            return [(0, n, -1)]
        k = max(1, min(n_segments - 1, L))
        cuts = sorted({(L * i) // k for i in range(k)})            # first layers of the groups
        segs = []
        end = n
        for first in reversed(cuts):
            segs.append((layer_off[first], end, first))
            end = layer_off[first]
        if end > 0:
            segs.append((0, end, -1))
        return segs

    def zero(self):
        self.flat.zero_()

    def allreduce_segment(self, i):
        s, e, _ = self.segments[i]
        dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM)

    def allreduce_mean(self, optim=None, already_summed=False):
        """Sum over ranks, then the mean: as a pass over the bucket, or (optim = FlatAdamW)
        folded into the fused optimiser kernel as its grad_scale."""
        if _world() > 1:
            if not already_summed:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if hasattr(optim, "grad_scale"):
                optim.grad_scale = 1.0 / _world()
            else:
                self.flat.mul_(1.0 / _world())


class OverlappedAllReduce:
    """Issue the bucket's all-reduce segment by segment FROM INSIDE backward, on a side stream,
    as soon as a segment's gradients are complete, so NCCL runs under the remaining backward
    kernels instead of after them ("overlapped with backward").

    A segment [layer k .. end) is complete when encoder layer k's backward has run, which is
    when autograd delivers the gradient of layer k's INPUT tensor: `watch(k, x)` registers a
    tensor hook on it.  The conv-stack segment is exchanged by `finish()` after backward.
    Everything is stream-ordered with events, so the same code is valid eagerly and under CUDA
    graph capture (the side stream forks from and joins the capturing stream)."""

    def __init__(self, bucket):
        self.bucket = bucket
        self.stream = torch.cuda.Stream() if bucket.flat.is_cuda else None   # None: gloo/CPU tests
        self._by_layer = {first: i for i, (_, _, first) in enumerate(bucket.segments)}
        self._done = set()
        self.active = False

    def begin(self):
        self._done = set()
        self.active = _world() > 1

    def _issue(self, i):
        if i in self._done:
            return
        self._done.add(i)
        if self.stream is None:
            self.bucket.allreduce_segment(i)
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.bucket.allreduce_segment(i)

    def watch(self, layer_index, x):
        """Call in forward with encoder layer `layer_index`'s input tensor."""
        if not self.active or not x.requires_grad:
            return
        i = self._by_layer.get(layer_index)
        if i is None:
            return
        x.register_hook(lambda g, i=i: (self._issue(i), None)[1])

    def finish(self):
        """After backward: exchange what is left and join the side stream."""
        if not self.active:
            return
        for i in range(len(self.bucket.segments)):
            self._issue(i)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.active = False


def to_device(batch, device):
    """transduction_model.py:200-202: per-tensor non_blocking H2D of the list-valued batch."""
    out = dict(batch)
    for k in ('raw_emg', 'emg', 'session_ids', 'audio_features', 'phonemes', 'text_int'):
        if k in batch:
            out[k] = [t.to(device, non_blocking=True) for t in batch[k]]
    return out


def _forward_backward(model, b, seq_len_frames, task, blank, overlap=None):
    X_raw = combine_fixed_length(b['raw_emg'], seq_len_frames * 8)
    if overlap is not None:
        overlap.begin()
        model.transformer.layer_input_hook = overlap.watch if overlap.active else None
    try:
        if task == "transduction":
            pred, phoneme_pred = model(None, X_raw, None)
            loss, _ = dtw_loss(pred, phoneme_pred, b)
        else:
            loss = ctc_loss_from_chunks(model(None, X_raw, None), b, blank)
        loss.backward()
    finally:
        if overlap is not None:
            model.transformer.layer_input_hook = None
    if overlap is not None:
        overlap.finish()
    return loss.detach()


def train_step(model, optim, batch, device, seq_len_frames=200, bucket=None, sync_loss=True,
               overlap=None):
    """One optimisation step.  `batch` follows EMGDataset.collate_raw (read_emg.py:262-296).
    Returns the loss (python float if sync_loss, else a 0-d device tensor)."""
    if bucket is not None:
        bucket.zero()
    else:
        optim.zero_grad(set_to_none=True)
    b = to_device(batch, device)
    loss = _forward_backward(model, b, seq_len_frames, "transduction", None, overlap)
    if bucket is not None:
        bucket.allreduce_mean(optim, already_summed=overlap is not None and _world() > 1)
    optim.step()
    return loss.item() if sync_loss else loss


class CtcAccumulator:
    """recognition_model.py:86,104-107: `optim.zero_grad()` once, then every batch adds its
    gradients and every `accumulate`-th batch steps the optimiser and clears them."""

    def __init__(self, accumulate=2):
        self.accumulate, self.micro = accumulate, 0

    def should_zero(self):
        return self.micro % self.accumulate == 0

    def should_step(self):
        self.micro += 1
        return self.micro % self.accumulate == 0


def ctc_train_step(model, optim, batch, device, seq_len_frames=200, bucket=None, blank=37,
                   acc=None, sync_loss=True):
    """One batch of recognition_model.py:89-107 (fused log-softmax + CTC, csrc/ctc.cu)."""
    acc = acc or CtcAccumulator(1)
    if acc.should_zero():
        if bucket is not None:
            bucket.zero()
        else:
            optim.zero_grad(set_to_none=True)
    b = to_device(batch, device)
    loss = _forward_backward(model, b, seq_len_frames, "recognition", blank)
    if acc.should_step():
        if bucket is not None:
            bucket.allreduce_mean(optim)
        optim.step()
    return loss.item() if sync_loss else loss


_BATCH_KEYS = {"transduction": ('raw_emg', 'audio_features', 'phonemes'),
               "recognition": ('raw_emg', 'text_int')}


class GraphedTrainStep:
    """train_step / ctc_train_step with forward + loss + backward replayed as ONE CUDA graph.

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
    The gradient clear and `optim.step()` stay outside the graph (a handful of launches), so any
    optimizer / LR schedule / accumulation count works unchanged (transduction_model.py:178-189,
    recognition_model.py:104-107).  With world_size > 1 and overlap=True the segmented NCCL
    all-reduce is captured INSIDE the graph on a side stream (OverlappedAllReduce) and runs under
    the remaining backward kernels; otherwise one all-reduce follows the replay.

    One graph per batch signature (utterance lengths, silent flags, train/eval): the first step
    of a new signature runs eagerly (lazy initialisation must not happen under capture), the
    second captures, later ones replay.
    """

    def __init__(self, model, optim, device, seq_len_frames=200, bucket=None, task="transduction",
                 accumulate=1, blank=37, overlap=False):
        self.model, self.optim, self.device = model, optim, torch.device(device)
        self.seq_len_frames, self.bucket = seq_len_frames, bucket
        if bucket is None:
            raise ValueError("GraphedTrainStep needs a GradientBucket: gradients must live at "
                             "fixed addresses across replays")
        if task not in _BATCH_KEYS:
            raise ValueError(task)
        self.task, self.blank = task, blank
        self.acc = CtcAccumulator(accumulate)
        # (accumulating several batches before the exchange leaves nothing to overlap per batch)
        self.overlap = (OverlappedAllReduce(bucket)
                        if (overlap and _world() > 1 and accumulate == 1) else None)
        self.seed_cell = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.shift_cell = torch.zeros((), dtype=torch.int64, device=self.device)
        self._seen = set()
        self._graphs = {}
        self.kernels_per_replay = 0

    def signature(self, batch, training):
        keys = _BATCH_KEYS[self.task]
        shapes = tuple(tuple(tuple(t.shape) for t in batch[k]) for k in keys)
        extra = tuple(bool(s) for s in batch['silent']) if self.task == "transduction" else ()
        return (shapes, extra, tuple(batch['lengths']), training)

    def _eager(self, batch, sync_loss):
        if self.acc.should_zero():
            self.bucket.zero()
        b = to_device(batch, self.device)
        loss = _forward_backward(self.model, b, self.seq_len_frames, self.task, self.blank,
                                 self.overlap)
        self._finish_step(self.overlap is not None)
        return loss.item() if sync_loss else loss

    def _finish_step(self, summed):
        if self.acc.should_step():
            self.bucket.allreduce_mean(self.optim, already_summed=summed)
            self.optim.step()

    def _capture(self, batch):
        static = dict(batch)
        for k in _BATCH_KEYS[self.task]:
            static[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch[k]]
        if self.task == "recognition":      # length vectors on the device before capture starts
            static['lengths_dev'] = torch.tensor(batch['lengths'], dtype=torch.int64,
                                                 device=self.device)
            static['text_int_lengths_dev'] = torch.tensor(batch['text_int_lengths'],
                                                          dtype=torch.int64, device=self.device)
        graph = torch.cuda.CUDAGraph()
        F_.take_capture_keepalive()
        F_.set_capture_execution_stream(torch.cuda.current_stream().cuda_stream)
        F_.set_seed_source(self.seed_cell)
        self.model.shift_source = self.shift_cell
        n0 = _lib.launch_count
        try:
            with torch.cuda.graph(graph):
                self.seed_cell.add_(0x9E3779B97F4A7C15 - (1 << 64))   # odd 64-bit stride
                loss = _forward_backward(self.model, static, self.seq_len_frames, self.task,
                                         self.blank, self.overlap)
        finally:
            self.model.shift_source = None
            F_.set_seed_source(None)
            F_.set_capture_execution_stream(None)
        # cached scratch buffers whose addresses the graph replays into live as long as the graph
        static['_keepalive'] = F_.take_capture_keepalive()
        return graph, static, loss, _lib.launch_count - n0

    def __call__(self, batch, sync_loss=True):
        sig = self.signature(batch, self.model.training)
        if sig not in self._graphs:
            if sig not in self._seen:      # first sight: a plain eager step
                self._seen.add(sig)
                return self._eager(batch, sync_loss)
            self._graphs[sig] = self._capture(batch)
        graph, static, loss, nk = self._graphs[sig]
        if self.acc.should_zero():
            self.bucket.zero()
        for k in _BATCH_KEYS[self.task]:
            if batch[k] and batch[k][0].is_cuda:
                torch._foreach_copy_(static[k], list(batch[k]))      # one fused D2D launch per key
            else:
                for dst, src in zip(static[k], batch[k]):            # pinned host -> device DMA
                    dst.copy_(src, non_blocking=True)
        if self.model.training:
            self.shift_cell.fill_(random.randrange(8))
        graph.replay()
        _lib.launch_count += nk
        self.kernels_per_replay = nk
        self._finish_step(self.overlap is not None)
        # the graph's loss tensor is overwritten by the next replay: hand out a copy
        return loss.item() if sync_loss else loss.clone()

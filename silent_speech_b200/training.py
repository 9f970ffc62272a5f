"""The training step of transduction_model.py:196-212 as one reusable call, plus the
data-parallel gradient exchange the north star adds (the reference is single-process).

train_step(model, optim, batch, device, seq_len_frames):
    H2D of the collate_raw-style batch -> combine_fixed_length -> Model forward -> dtw_loss ->
    backward -> [flat-bucket NCCL all-reduce of gradients] -> optimizer step.
"""
import torch
import torch.distributed as dist

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

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
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
        bucket.allreduce_mean()
    optim.step()
    return loss.item() if sync_loss else loss.detach()

"""Fused AdamW on the flat gradient bucket.

Opt-in replacement for `torch.optim.AdamW(model.parameters(), weight_decay=FLAGS.l2)` +
the per-iteration `param_group['lr'] = ...` schedule of transduction_model.py:178-189,210.
Parameters are re-pointed into ONE flat fp32 buffer (names, shapes and `state_dict` unchanged),
gradients already live in the `GradientBucket` that the data-parallel all-reduce uses, and
`step()` is a single streaming kernel (csrc/optim.cu) that also applies the 1/world_size of the
gradient mean.  Learning rate and step count are device cells, so the update can sit inside a
captured CUDA graph.  It is a `torch.optim.Optimizer`, so `ReduceLROnPlateau`
(transduction_model.py:179) and the reference's `set_lr` loop work on it unchanged.
"""
import torch

from . import _lib
from . import weights


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        """bucket: training.GradientBucket of the model (defines the parameter order)."""
        params = list(bucket.params)
        if not params or any(p.dtype != torch.float32 or not p.is_cuda for p in params):
            raise TypeError("FlatAdamW: CUDA float32 parameters only (libssb has no CPU path)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.bucket = bucket
        dev = params[0].device
        n = bucket.flat.numel()
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view              # same Parameter object, storage now inside the bucket
                off += p.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.lr_cell = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self.step_cell = torch.zeros((), dtype=torch.int64, device=dev)
        self._lr_on_device = float(lr)
        self.grad_scale = 1.0              # set to 1/world_size when the bucket holds a SUM

    def zero_grad(self, set_to_none=False):
        # p.grad must stay views of the bucket (fixed addresses for CUDA graphs / the all-reduce)
        self.bucket.zero()

    def push_lr(self):
        """Write param_groups[0]['lr'] to the device cell if the schedule changed it."""
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_on_device:
            self.lr_cell.fill_(lr)
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.push_lr()
        lib = _lib.load()
        _lib.check(lib.ssb_adamw_flat(
            self.flat_param.data_ptr(), self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(),
            self.exp_avg_sq.data_ptr(), self.flat_param.numel(), self.lr_cell.data_ptr(),
            self.step_cell.data_ptr(), g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'],
            self.grad_scale, _lib.current_stream()))
        weights.bump_epoch()     # parameters were written through raw pointers: planes are stale
        return loss

    def state_dict(self):
        return {"step": int(self.step_cell.item()), "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"}
                                 for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_cell.fill_(int(sd["step"]))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        self.push_lr()

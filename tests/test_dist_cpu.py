"""CPU, world_size 2, gloo: the data-parallel gradient exchange (GradientBucket) used by the
N > 1 path — flat bucket aliasing, mean all-reduce, and that the no-gradient position tables
stay out of the bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(4, 3)
        self.relative_positional = nn.Module()
        self.relative_positional.embeddings = nn.Parameter(torch.zeros(2, 3))
        self.b = nn.Linear(3, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from silent_speech_b200.training import GradientBucket
    torch.manual_seed(0)
    m = Tiny()
    bucket = GradientBucket(m)
    assert bucket.flat.numel() == sum(p.numel() for n, p in m.named_parameters()
                                      if "embeddings" not in n)
    bucket.zero()
    x = torch.full((5, 4), float(rank + 1))
    m.b(m.a(x)).sum().backward()
    local = bucket.flat.clone()
    assert torch.equal(m.a.weight.grad.flatten(), bucket.flat[:12])     # grads alias the bucket
    bucket.allreduce_mean()
    mean = bucket.flat.clone()
    # with a fused optimiser (FlatAdamW exposes grad_scale) the bucket keeps the SUM and the
    # 1/world of the mean is handed to the optimiser kernel instead of a pass over the bucket
    bucket.flat.copy_(local)

    class Opt:
        grad_scale = 1.0
    opt = Opt()
    bucket.allreduce_mean(opt)
    assert opt.grad_scale == 0.5
    assert torch.allclose(bucket.flat * opt.grad_scale, mean)
    q.put((rank, local, mean))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (res[0][1] + res[1][1]) / 2
    assert torch.allclose(res[0][2], mean) and torch.allclose(res[1][2], mean)
    assert not torch.allclose(res[0][1], res[1][1])

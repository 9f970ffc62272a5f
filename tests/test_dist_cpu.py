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


class TinyEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(4, 4) for _ in range(5)])
        self.layer_input_hook = None

    def forward(self, x):
        for i, l in enumerate(self.layers):
            if self.layer_input_hook is not None:
                self.layer_input_hook(i, x)
            x = torch.tanh(l(x)) + x
        return x


class TinyModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Linear(4, 4)
        self.transformer = TinyEncoder()
        self.w_out = nn.Linear(4, 2)
        self.bn = nn.BatchNorm1d(2)

    def forward(self, x):
        return self.w_out(self.transformer(self.conv(x)))


def _overlap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from silent_speech_b200.training import GradientBucket, OverlappedAllReduce, broadcast_model
    torch.manual_seed(rank)                      # replicas start DIFFERENT ...
    m = TinyModel()
    m.bn.running_mean.fill_(float(rank))
    broadcast_model(m)                           # ... and are made equal to rank 0's
    w0 = m.conv.weight.detach().clone()
    rm = m.bn.running_mean.clone()
    bucket = GradientBucket(m, n_segments=3)
    n = bucket.flat.numel()
    segs = bucket.segments
    # segments tile the bucket, in completion order (tail first), cut at layer boundaries
    assert segs[0][1] == n and segs[-1][0] == 0 and segs[-1][2] == -1
    assert all(segs[i][0] == segs[i + 1][1] for i in range(len(segs) - 1))
    assert [f for _, _, f in segs] == [2, 0, -1]
    x = torch.randn(6, 4, generator=torch.Generator().manual_seed(10 + rank))
    bucket.zero()
    m(x).square().sum().backward()
    local = bucket.flat.clone()
    ov = OverlappedAllReduce(bucket)
    bucket.zero()
    ov.begin()
    issued = []
    orig = bucket.allreduce_segment
    bucket.allreduce_segment = lambda i: (issued.append(i), orig(i))[1]
    m.transformer.layer_input_hook = ov.watch
    m(x).square().sum().backward()
    m.transformer.layer_input_hook = None
    during_backward = list(issued)
    ov.finish()
    q.put((rank, local, bucket.flat.clone(), during_backward, list(issued), w0, rm))
    dist.destroy_process_group()


def test_overlapped_segment_allreduce_and_broadcast_world2():
    """The segmented exchange issued from backward hooks gives exactly the SUM over ranks of the
    local gradients (every segment reduced once, none before its gradients were complete), and
    broadcast_model makes replicas identical (parameters and BatchNorm buffers)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    for r in res:
        assert torch.allclose(r[2], total, rtol=1e-6, atol=1e-7)
        assert r[3] == [0, 1]            # layer-2 and layer-0 hooks fired inside backward
        assert r[4] == [0, 1, 2]         # conv-stack segment by finish()
    assert torch.equal(res[0][5], res[1][5]) and torch.equal(res[0][6], res[1][6])
    assert float(res[1][6][0]) == 0.0    # rank 1's buffer now holds rank 0's value

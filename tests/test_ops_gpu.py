"""GPU unit parity of every model kernel family against plain PyTorch fp32 on the same
device (forward and backward).  fp32 CUDA-core arithmetic: tolerances 2e-5 relative L2."""
import math

import pytest
import torch
import torch.nn.functional as F

from silent_speech_b200 import functional as SF

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def close(a, b, tol=2e-5, what=""):
    r = rel(a.double(), b.double())
    assert r < tol, f"{what}: rel-L2 {r:.3e} >= {tol}"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


@pytest.mark.parametrize("M,K,N", [(64, 32, 48), (300, 768, 80), (1000, 24, 768), (257, 100, 132),
                                   (4096, 768, 768), (5000, 24, 768), (4097, 32, 260),
                                   (6000, 20, 64)])
def test_linear_fwd_bwd(M, K, N):
    x = rnd(M, K, seed=1).requires_grad_(True)
    W = rnd(K, N, seed=2, scale=K ** -0.5).requires_grad_(True)
    b = rnd(N, seed=3).requires_grad_(True)
    y = SF.linear(x, W, b)
    yr = x.detach() @ W.detach() + b.detach()
    close(y, yr, what="y")
    g = rnd(M, N, seed=4)
    y.backward(g)
    close(x.grad, g @ W.detach().t(), what="dx")
    close(W.grad, x.detach().t() @ g, what="dW")
    close(b.grad, g.sum(0), what="db")


def test_engines_agree_linear(monkeypatch):
    """tcgen05 (bf16x3) and CUDA-core engines on the same problem: both within 2e-5 of fp64."""
    M, K, N = 2048, 768, 1024
    x, W, b = rnd(M, K, seed=1), rnd(K, N, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    ref = (x.double() @ W.double() + b.double())
    outs = {}
    for eng in ("tc", "simt"):
        monkeypatch.setenv("SSB_GEMM", eng)
        xg, Wg = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
        y = SF.linear(xg, Wg, b)
        y.backward(torch.ones_like(y))
        close(y, ref, what=eng + " y")
        close(Wg.grad, x.double().t() @ torch.ones(M, N, device=dev, dtype=torch.float64), what=eng + " dW")
        close(xg.grad, torch.ones(M, N, device=dev, dtype=torch.float64) @ W.double().t(), what=eng + " dx")
        outs[eng] = y.detach()
    assert not torch.equal(outs["tc"], outs["simt"])      # really two different engines


def test_ffn_fwd_bwd_no_dropout(monkeypatch):
    monkeypatch.setenv("SSB_GEMM", "simt")    # exact fp32: masks identical to torch's
    M, D, Fh = 777, 64, 3072
    x = rnd(M, D, seed=1).requires_grad_(True)
    W1 = rnd(D, Fh, seed=2, scale=D ** -0.5).requires_grad_(True)
    b1 = rnd(Fh, seed=3, scale=0.1).requires_grad_(True)
    W2 = rnd(Fh, D, seed=4, scale=Fh ** -0.5).requires_grad_(True)
    b2 = rnd(D, seed=5, scale=0.1).requires_grad_(True)
    y = SF.ffn(x, W1, b1, W2, b2, 0.0, 0, 0)
    g = rnd(M, D, seed=6)
    y.backward(g)
    xs = [t.detach().clone().requires_grad_(True) for t in (x, W1, b1, W2, b2)]
    yr = F.relu(xs[0] @ xs[1] + xs[2]) @ xs[3] + xs[4]
    yr.backward(g)
    close(y, yr, what="y")
    for a, b, n in zip((x, W1, b1, W2, b2), xs, "x W1 b1 W2 b2".split()):
        close(a.grad, b.grad, what="d" + n)


@pytest.mark.parametrize("M,D,Fh", [(777, 64, 3072), (450, 128, 3072), (1000, 768, 3072)])
def test_ffn_tensor_core_engine_vs_fp64(M, D, Fh):
    """tcgen05 bf16x3 engine: forward within 2e-5 of fp64; backward checked against fp64
    formulas evaluated with OUR ReLU mask (gradients through a ReLU kink are only comparable
    at identical masks: a 4e-6 forward perturbation flips ~1e-5 of them)."""
    x = rnd(M, D, seed=1).requires_grad_(True)
    W1 = rnd(D, Fh, seed=2, scale=D ** -0.5).requires_grad_(True)
    b1 = rnd(Fh, seed=3, scale=0.1).requires_grad_(True)
    W2 = rnd(Fh, D, seed=4, scale=Fh ** -0.5).requires_grad_(True)
    b2 = rnd(D, seed=5, scale=0.1).requires_grad_(True)
    y = SF.ffn(x, W1, b1, W2, b2, 0.0, 0, 0)
    g = rnd(M, D, seed=6)
    y.backward(g)
    xd, W1d, b1d, W2d, b2d, gd = (t.detach().double() for t in (x, W1, b1, W2, b2, g))
    z = xd @ W1d + b1d
    close(y, torch.relu(z) @ W2d + b2d, what="y")
    # recover our mask from an identical forward GEMM
    h = torch.empty(M, Fh, device=dev)
    SF.mm_fwd(x.detach(), W1.detach(), dict(bias=b1.detach(), relu=1), h, M, Fh, D)
    mask = (h > 0).double()
    assert (mask != (z > 0).double()).float().mean().item() < 1e-4     # only kink-adjacent flips
    hd = h.double()
    dh = (gd @ W2d.t()) * mask
    close(W2.grad, hd.t() @ gd, tol=5e-5, what="dW2")
    close(b2.grad, gd.sum(0), what="db2")
    close(W1.grad, xd.t() @ dh, tol=5e-5, what="dW1")
    close(b1.grad, dh.sum(0), tol=5e-5, what="db1")
    close(x.grad, dh @ W1d.t(), tol=5e-5, what="dx")


def test_gemm_epilogue_split_plane_operands():
    """tcgen05 epilogue: planes_out (with and without the fp32 copy) equals split_planes of the
    fp32 result bit for bit; mask_planes (hi plane) equals mask_src; colsum_planes = colsum."""
    from silent_speech_b200.functional import (_epi, _scatter_plain, colsum, colsum_planes,
                                               gemm_tc_kmajor, split_planes, tc_operand_plain)
    M, K, N = 1000, 128, 520
    x, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    xp, wp = split_planes(x), split_planes(W)
    y = torch.empty(M, N, device=dev)
    gemm_tc_kmajor(tc_operand_plain(xp, M, K), wp, N, K,
                   _epi(_scatter_plain(y.data_ptr(), M, N), bias=b, relu=1))
    want = split_planes(y)
    for with_fp32 in (True, False):
        y2 = torch.full((M, N), float("nan"), device=dev)
        pl = torch.zeros(2, M, N, dtype=torch.bfloat16, device=dev)
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), wp, N, K,
                       _epi(_scatter_plain(y2.data_ptr() if with_fp32 else None, M, N), bias=b,
                            relu=1, planes_out=pl))
        assert torch.equal(pl, want)
        assert torch.equal(y2, y) if with_fp32 else torch.isnan(y2).all()
    g = rnd(M, K, seed=4)
    outs = []
    for kw in (dict(mask_src=y), dict(mask_planes=want[0])):
        o = torch.empty(M, N, device=dev)
        gemm_tc_kmajor(tc_operand_plain(split_planes(g), M, K), wp, N, K,
                       _epi(_scatter_plain(o.data_ptr(), M, N), mask_scale=1.25, **kw))
        outs.append(o)
    assert torch.equal(outs[0], outs[1]) and (outs[0] == 0).float().mean() > 0.3
    close(colsum_planes(want), colsum(y), tol=1e-5, what="colsum_planes")
    with pytest.raises(RuntimeError):      # CUDA-core engine rejects split-plane epilogue operands
        from silent_speech_b200.functional import _gather_plain, gemm_nn
        gemm_nn(_gather_plain(x.data_ptr(), M, K, K), W.t().contiguous(),
                _epi(_scatter_plain(y.data_ptr(), M, N), planes_out=pl), M, N, K)


def test_ffn_native_layout_and_gradient_sinks():
    """ffn_native (nn.Linear-layout weights, transposed split, parameter-layout weight gradients)
    equals ffn on the transposed copies; with registered gradient sinks the gradients are
    accumulated into the persistent .grad views and autograd receives None."""
    M, D, Fh, p = 1000, 768, 3072, 0.2
    x = rnd(M, D, seed=1)
    w1 = rnd(Fh, D, seed=2, scale=D ** -0.5)
    b1 = rnd(Fh, seed=3, scale=0.1)
    w2 = rnd(D, Fh, seed=4, scale=Fh ** -0.5)
    b2 = rnd(D, seed=5, scale=0.1)
    g = rnd(M, D, seed=6)

    def run(native, sinks):
        ps = [torch.nn.Parameter(t.clone()) for t in (w1, b1, w2, b2)]
        xx = x.clone().requires_grad_(True)
        if sinks:
            for q in ps:
                q.grad = torch.full_like(q, 0.5)          # pre-existing content must be kept
            SF.register_grad_sinks(ps)
        if native:
            y = SF.ffn_native(xx, ps[0], ps[1], ps[2], ps[3], p, 11, 3)
        else:
            y = SF.ffn(xx, ps[0].t().contiguous(), ps[1], ps[2].t().contiguous(), ps[3], p, 11, 3)
        y.backward(g)
        return [y.detach(), xx.grad] + [q.grad - (0.5 if sinks else 0.0) for q in ps]

    base = run(False, False)
    for native, sinks in ((True, False), (True, True)):
        got = run(native, sinks)
        assert torch.equal(got[0], base[0])                                  # same forward GEMMs
        for a, b, n in zip(got[1:], base[1:], "dx dw1 db1 dw2 db2".split()):
            close(a, b, tol=2e-5, what=f"{n} native={native} sinks={sinks}")
    # a parameter whose .grad was replaced no longer has a sink
    q = torch.nn.Parameter(w1.clone())
    q.grad = torch.zeros_like(q)
    SF.register_grad_sinks([q])
    assert SF._sink(q) is q.grad
    q.grad = torch.zeros_like(q)
    assert SF._sink(q) is None


def test_producer_kernels_emit_exact_operand_planes():
    """BatchNorm apply / backward and add+dropout+LayerNorm forward / backward attach the bf16
    split planes of their result (written by the same kernel): bit-identical to a split pass over
    the fp32 result, picked up by planes_of(), dropped after an in-place modification."""
    rows, C = 520, 128
    xa = rnd(1, rows, C, seed=1).requires_grad_(True)
    xb = rnd(1, rows, C, seed=2).requires_grad_(True)
    ps = [(rnd(C, seed=10 + i, scale=0.3) + (1.0 if i % 2 == 0 else 0.0)).requires_grad_(True)
          for i in range(4)]
    rm = [torch.zeros(C, device=dev) for _ in range(2)]
    rv = [torch.ones(C, device=dev) for _ in range(2)]
    y = SF.bn_act(xa, ps[0], ps[1], rm[0], rv[0], True, True, xb, ps[2], ps[3], rm[1], rv[1])
    assert torch.equal(SF.planes_of(y), SF.split_planes(y.detach())) and hasattr(y, "_ssb_planes")
    got = {}

    def grab(key):
        def hook(g):
            got[key] = (g, getattr(g, "_ssb_planes", None))
        return hook
    xa.register_hook(grab("a"))
    y.backward(rnd(1, rows, C, seed=3))
    g, ent = got["a"]
    assert ent is not None and torch.equal(ent[1], SF.split_planes(g.contiguous()))

    res, br = rnd(rows, C, seed=4).requires_grad_(True), rnd(rows, C, seed=5).requires_grad_(True)
    gam, bet = (rnd(C, seed=6) + 1).requires_grad_(True), rnd(C, seed=7).requires_grad_(True)
    o = SF.add_dropout_layernorm(res, br, gam, bet, 0.2, 99, 5)
    assert torch.equal(SF.planes_of(o), SF.split_planes(o.detach())) and hasattr(o, "_ssb_planes")
    br.register_hook(grab("b"))
    o.backward(rnd(rows, C, seed=8))
    g, ent = got["b"]
    assert ent is not None and torch.equal(ent[1], SF.split_planes(g.contiguous()))
    # stale planes are never used
    with torch.no_grad():
        o.add_(1.0)
    assert torch.equal(SF.planes_of(o), SF.split_planes(o.detach()))


def test_split_planes_transposed():
    x = rnd(200, 136, seed=9)
    assert torch.equal(SF.split_planes_t(x), SF.split_planes(x.t().contiguous()))


def test_ffn_dropout_statistics_and_consistency():
    M, D, Fh, p = 512, 32, 3072, 0.2
    x = rnd(M, D, seed=1).requires_grad_(True)
    W1 = rnd(D, Fh, seed=2, scale=D ** -0.5)
    b1 = torch.ones(Fh, device=dev)           # keep pre-activations mostly positive
    W2 = torch.eye(Fh, D, device=dev).contiguous()
    b2 = torch.zeros(D, device=dev)
    # h itself is internal; observe it through W2 = identity-ish rows: use a direct GEMM instead
    from silent_speech_b200.functional import _epi, _gather_plain, _scatter_plain, gemm_nn
    h = torch.empty(M, Fh, device=dev)
    gemm_nn(_gather_plain(x.data_ptr(), M, D, D), W1,
            _epi(_scatter_plain(h.data_ptr(), M, Fh), bias=b1, relu=1, drop_p=p, seed=123, site=7),
            M, Fh, D)
    h0 = F.relu(x.detach() @ W1 + b1)
    pos = h0 > 0
    kept = (h > 0) & pos
    rate = kept.sum().item() / pos.sum().item()
    assert abs(rate - (1 - p)) < 0.005, rate                      # keep-rate
    close(h[kept], h0[kept] / (1 - p), what="scaling 1/(1-p)")    # inverted dropout scale
    h2 = torch.empty_like(h)                                        # same (seed, site) -> same mask
    gemm_nn(_gather_plain(x.data_ptr(), M, D, D), W1,
            _epi(_scatter_plain(h2.data_ptr(), M, Fh), bias=b1, relu=1, drop_p=p, seed=123, site=7),
            M, Fh, D)
    assert torch.equal(h, h2)
    gemm_nn(_gather_plain(x.data_ptr(), M, D, D), W1,
            _epi(_scatter_plain(h2.data_ptr(), M, Fh), bias=b1, relu=1, drop_p=p, seed=123, site=8),
            M, Fh, D)
    assert not torch.equal(h, h2)                                   # other site -> other mask


@pytest.mark.parametrize("B,L,Cin,Cout,k,s", [(3, 50, 8, 32, 3, 2), (2, 37, 8, 32, 3, 2),
                                              (2, 40, 32, 32, 3, 1), (2, 41, 32, 48, 3, 2),
                                              (2, 41, 32, 48, 1, 2), (2, 40, 32, 48, 1, 2),
                                              (4, 500, 64, 64, 3, 1), (1, 1, 8, 16, 3, 2),
                                              (2, 2, 16, 16, 3, 2), (2, 250, 768, 768, 3, 2),
                                              # thin-K kernels (K <= 32, >= 4096 output rows):
                                              (5, 2001, 8, 768, 3, 2), (3, 4000, 8, 132, 1, 2),
                                              (2, 4100, 4, 64, 3, 1)])
def test_conv_fwd_bwd(B, L, Cin, Cout, k, s):
    x = rnd(B, L, Cin, seed=1).requires_grad_(True)
    w = rnd(Cout, Cin, k, seed=2, scale=(Cin * k) ** -0.5).requires_grad_(True)
    b = rnd(Cout, seed=3, scale=0.1).requires_grad_(True)
    Wg = w.permute(2, 1, 0).reshape(k * Cin, Cout).contiguous()
    y = SF.conv1d_cl(x, Wg, b, k, s)
    xr = x.detach().clone().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        yr = F.conv1d(xr.transpose(1, 2), wr, br, stride=s, padding=1 if k == 3 else 0).transpose(1, 2)
        assert y.shape == yr.shape == (B, (L - 1) // s + 1, Cout)
        g = rnd(*y.shape, seed=4)
        y.backward(g)
        yr.backward(g)
    close(y, yr, what="y")
    close(x.grad, xr.grad, what="dx")
    close(w.grad, wr.grad, what="dw")
    close(b.grad, br.grad, what="db")


@pytest.mark.parametrize("rows,C,two,relu,training", [(600, 32, False, True, True),
                                                      (1000, 768, True, True, True),
                                                      (333, 64, True, True, False),
                                                      (257, 32, False, False, True)])
def test_batchnorm_act_fwd_bwd(rows, C, two, relu, training):
    B, L = 1, rows
    xa = (rnd(B, L, C, seed=1) * 2 + 0.5).requires_grad_(True)
    xb = (rnd(B, L, C, seed=2) - 0.3).requires_grad_(True) if two else None
    ps = [rnd(C, seed=10 + i, scale=0.3) + (1.0 if i % 2 == 0 else 0.0) for i in range(4)]
    ps = [p.requires_grad_(True) for p in ps]
    rm = [rnd(C, seed=20 + i, scale=0.2) for i in range(2)]
    rv = [rnd(C, seed=30 + i).abs() + 0.5 for i in range(2)]
    rm_r = [t.clone() for t in rm]
    rv_r = [t.clone() for t in rv]
    y = SF.bn_act(xa, ps[0], ps[1], rm[0], rv[0], training, relu,
                  xb, ps[2] if two else None, ps[3] if two else None,
                  rm[1] if two else None, rv[1] if two else None)
    xar = xa.detach().clone().requires_grad_(True)
    psr = [p.detach().clone().requires_grad_(True) for p in ps]
    yr = F.batch_norm(xar.view(-1, C), rm_r[0], rv_r[0], psr[0], psr[1], training, 0.1, 1e-5)
    if two:
        xbr = xb.detach().clone().requires_grad_(True)
        yr = yr + F.batch_norm(xbr.view(-1, C), rm_r[1], rv_r[1], psr[2], psr[3], training, 0.1, 1e-5)
    if relu:
        yr = F.relu(yr)
    yr = yr.view_as(xa)
    close(y, yr, what="y")
    g = rnd(*y.shape, seed=5)
    y.backward(g)
    yr.backward(g)
    close(xa.grad, xar.grad, tol=5e-5, what="dxa")
    close(ps[0].grad, psr[0].grad, tol=5e-5, what="dgamma_a")
    close(ps[1].grad, psr[1].grad, tol=5e-5, what="dbeta_a")
    if two:
        close(xb.grad, xbr.grad, tol=5e-5, what="dxb")
        close(ps[2].grad, psr[2].grad, tol=5e-5, what="dgamma_b")
    for a, b in zip(rm + rv, rm_r + rv_r):
        close(a, b, tol=1e-5, what="running stats")


@pytest.mark.parametrize("rows,D", [(100, 32), (1000, 768), (37, 256), (64, 1024)])
def test_add_layernorm_fwd_bwd(rows, D):
    res = rnd(rows, D, seed=1).requires_grad_(True)
    br = rnd(rows, D, seed=2).requires_grad_(True)
    g_ = (rnd(D, seed=3, scale=0.2) + 1).requires_grad_(True)
    b_ = rnd(D, seed=4, scale=0.2).requires_grad_(True)
    y = SF.add_dropout_layernorm(res, br, g_, b_, 0.0, 0, 0)
    rs = [t.detach().clone().requires_grad_(True) for t in (res, br, g_, b_)]
    yr = F.layer_norm(rs[0] + rs[1], (D,), rs[2], rs[3], 1e-5)
    close(y, yr, what="y")
    g = rnd(rows, D, seed=5)
    y.backward(g)
    yr.backward(g)
    for a, b, n in zip((res, br, g_, b_), rs, ("dres", "dbranch", "dgamma", "dbeta")):
        close(a.grad, b.grad, tol=5e-5, what=n)


def test_add_layernorm_dropout_mask_consistency():
    rows, D, p = 2000, 256, 0.2
    res = torch.zeros(rows, D, device=dev)
    br = torch.ones(rows, D, device=dev).requires_grad_(True)
    g_ = torch.ones(D, device=dev)
    b_ = torch.zeros(D, device=dev)
    # forward z = dropout(1): recover the mask from the backward: d_branch = dz * mask / (1-p)
    y = SF.add_dropout_layernorm(res.requires_grad_(True), br, g_, b_, p, 99, 5)
    y.backward(torch.randn_like(y))
    zero_frac = (br.grad == 0).float().mean().item()
    assert abs(zero_frac - p) < 0.01, zero_frac
    # where the branch gradient is non-zero it equals d_res / (1-p)
    nz = br.grad != 0
    close(br.grad[nz], res.grad[nz] / (1 - p), what="mask scale")


def dense_attention_ref(qkv, E, B, T, H, dh, W):
    """Dense restatement: logits = qk/sqrt(dh) + q.E[k-q+W] inside the band, -inf outside."""
    D = H * dh
    q, k, v = (qkv[:, i * D:(i + 1) * D].view(B, T, H, dh).permute(0, 2, 1, 3) for i in range(3))
    logits = q @ k.transpose(-1, -2) / math.sqrt(dh)
    R = torch.einsum('bhqa,hra->bhqr', q, E[:, :2 * W + 1])
    ar = torch.arange(T, device=qkv.device)
    relidx = ar[None, :] - ar[:, None] + W
    inb = (relidx >= 0) & (relidx <= 2 * W)
    pos = torch.gather(R, 3, relidx.clamp(0, 2 * W)[None, None].expand(B, H, T, T))
    logits = torch.where(inb[None, None], logits + pos, torch.full_like(logits, float("-inf")))
    o = torch.softmax(logits, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * T, D)


@pytest.mark.parametrize("sched", ["simt", "tc", "fused"])
@pytest.mark.parametrize("B,T,H,dh,W", [(2, 25, 8, 4, 99), (2, 125, 8, 4, 99), (1, 130, 2, 32, 99),
                                        (2, 250, 8, 96, 99), (1, 33, 1, 8, 5), (1, 64, 2, 16, 31),
                                        (3, 500, 8, 96, 99), (2, 126, 4, 16, 99),
                                        (1, 64, 2, 32, 31), (2, 300, 4, 64, 99), (1, 129, 1, 96, 5),
                                        (1, 384, 2, 96, 99)])
def test_band_attention_fwd_bwd(B, T, H, dh, W, sched, monkeypatch):
    """Both schedules of the same op: CUDA-core band kernels and the tensor-core (batched
    tcgen05 GEMM) schedule, against a dense torch restatement."""
    monkeypatch.setenv("SSB_ATTN", sched)
    if sched == "tc" and not SF._tc_attn_ok(T, dh, W):
        pytest.skip("shape not eligible for the tensor-core schedule")
    if sched == "fused" and not SF._fused_attn_ok(B, T, H, dh, W):
        pytest.skip("shape not eligible for the fused kernels")
    D = H * dh
    qkv = rnd(B * T, 3 * D, seed=1).requires_grad_(True)
    RW = (2 * W + 1 + 3) // 4 * 4
    E = torch.zeros(H, RW, dh, device=dev)
    E[:, :2 * W + 1] = rnd(H, 2 * W + 1, dh, seed=2, scale=dh ** -0.5)
    o = SF.band_attention(qkv, E, B, T, H, dh, W, 0.0, 0, 0)
    qr = qkv.detach().clone().requires_grad_(True)
    orf = dense_attention_ref(qr, E, B, T, H, dh, W)
    close(o, orf, what="O")
    g = rnd(B * T, D, seed=3)
    o.backward(g)
    orf.backward(g)
    close(qkv.grad[:, :D], qr.grad[:, :D], tol=5e-5, what="dq")
    close(qkv.grad[:, D:2 * D], qr.grad[:, D:2 * D], tol=5e-5, what="dk")
    close(qkv.grad[:, 2 * D:], qr.grad[:, 2 * D:], tol=5e-5, what="dv")


@pytest.mark.parametrize("sched,T", [("simt", 40), ("tc", 96), ("fused", 96)])
def test_band_attention_dropout_is_consistent_between_fwd_and_bwd(sched, T, monkeypatch):
    """With V = one-hot positions the forward output exposes the dropped probabilities; the
    backward must use the same mask: check d(sum O)/dV against the forward's P_drop."""
    monkeypatch.setenv("SSB_ATTN", sched)
    B, H, dh, W, p = 1, 1, T, 99, 0.3
    D = H * dh
    qkv = torch.zeros(B * T, 3 * D, device=dev)
    qkv[:, :2 * D] = rnd(B * T, 2 * D, seed=1)
    qkv[:, 2 * D:] = torch.eye(T, device=dev)            # v[k] = e_k  => O[q, :] = P_drop[q, :]
    qkv.requires_grad_(True)
    E = torch.zeros(H, 200, dh, device=dev)
    o = SF.band_attention(qkv, E, B, T, H, dh, W, p, 1234, 3)
    Pdrop = o.detach()                                    # (T, T)
    zero = (Pdrop == 0).float().mean().item()
    assert abs(zero - p) < 0.05, zero
    if sched == "tc":
        assert SF._tc_attn_ok(T, dh, W)
    if sched == "fused":
        assert SF._fused_attn_ok(B, T, H, dh, W)
    g = rnd(T, D, seed=5)
    o.backward(g)
    dV_expected = Pdrop.t() @ g                           # dV = P_drop^T dO
    close(qkv.grad[:, 2 * D:], dV_expected, tol=5e-5, what="dV with dropout")
    o2 = SF.band_attention(qkv.detach(), E, B, T, H, dh, W, p, 1234, 3)
    assert torch.equal(o2, Pdrop)


@pytest.mark.parametrize("B,T,H,dh", [(2, 500, 8, 96), (1, 200, 2, 32)])
def test_fused_attention_matches_tc_schedule_with_dropout(B, T, H, dh, monkeypatch):
    """The fused kernels key dropout exactly like the multi-kernel tensor-core schedule (same
    Philox counters per (row, 4 keys)), so with dropout ON both schedules must agree on the
    output and on every gradient: pins the mask plumbing of the fused backward (lanes = keys,
    4 x 4 cross-lane transpose of the Philox words)."""
    W, p = 99, 0.2
    D = H * dh
    E = torch.zeros(H, 200, dh, device=dev)
    E[:, :2 * W + 1] = rnd(H, 2 * W + 1, dh, seed=2, scale=dh ** -0.5)
    g = rnd(B * T, D, seed=3)
    out = {}
    for sched in ("tc", "fused"):
        monkeypatch.setenv("SSB_ATTN", sched)
        qkv = rnd(B * T, 3 * D, seed=1).requires_grad_(True)
        o = SF.band_attention(qkv, E, B, T, H, dh, W, p, 4321, 7)
        o.backward(g)
        out[sched] = (o.detach(), qkv.grad.detach())
    assert SF._fused_attn_ok(B, T, H, dh, W)
    close(out["fused"][0], out["tc"][0], tol=2e-5, what="O (dropout)")
    close(out["fused"][1][:, :D], out["tc"][1][:, :D], tol=5e-5, what="dq (dropout)")
    close(out["fused"][1][:, D:2 * D], out["tc"][1][:, D:2 * D], tol=5e-5, what="dk (dropout)")
    close(out["fused"][1][:, 2 * D:], out["tc"][1][:, 2 * D:], tol=5e-5, what="dv (dropout)")

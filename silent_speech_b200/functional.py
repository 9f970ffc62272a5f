"""torch.autograd bindings of the libssb model kernels (GEMM family, BatchNorm, band attention,
add+dropout+LayerNorm).  Autograd lives here in Python; every forward/backward body is a
sequence of C-ABI calls on raw device pointers (csrc/*.cu).  There is no CPU path: tensors
must be CUDA fp32.

Layout conventions (B200-first, not the reference's NCL / seq-first):
  activations are channels-last / token-major: conv stack (B, L, C), transformer (B*T, D);
  weights are handed to the kernels in "GEMM layout" W[k, n] (n contiguous); the model
  derives them from the reference-shaped parameters with differentiable torch views.
"""
import ctypes

import torch

from . import _lib
import os

from ._lib import Epilogue, Gather, Scatter, TcOperand

_f32 = torch.float32


def _chk(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != _f32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def _stream():
    return _lib.current_stream()


def set_seed_source(cell):
    """Register (or detach with None) the device-resident dropout-seed offset: a 1-element int64
    CUDA tensor the caller keeps alive and bumps between CUDA-graph replays (include/ssb.h,
    ssb_set_seed_source)."""
    lib = _lib.load()
    if cell is None:
        _lib.check(lib.ssb_set_seed_source(None))
        return
    _lib.require_cuda(cell, "seed cell")
    if cell.dtype != torch.int64 or cell.numel() != 1:
        raise TypeError("seed cell must be a 1-element int64 CUDA tensor")
    _lib.check(lib.ssb_set_seed_source(cell.data_ptr()))


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _gather_plain(ptr, M, K, ld):
    return Gather(ptr, 0, M, K, M, ld, 1, 0, 0)


def _scatter_plain(ptr, M, ld):
    return Scatter(ptr, 0, M, ld, 1, 0)


def _epi(out, bias=None, relu=0, accumulate=0, mask_src=None, mask_scale=1.0, drop_p=0.0, seed=0,
         site=0, planes_out=None, mask_planes=None, mask_bits=None, mask_bits_out=None,
         planes_lrelu=None):
    """ssb_epilogue_t.  planes_out: (2, M, N) bf16 tensor receiving the result as split planes
    (tcgen05 engine); mask_planes: bf16 hi plane (M, N) standing in for mask_src; mask_bits /
    mask_bits_out: (M, N / 8) uint8 bit masks (result > 0) read / written by the epilogue;
    planes_lrelu: negative slope of a leaky ReLU applied to the plane copy only."""
    return Epilogue(out, bias.data_ptr() if bias is not None else None,
                    mask_src.data_ptr() if mask_src is not None else None, mask_scale, int(relu),
                    int(accumulate), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site),
                    planes_out.data_ptr() if planes_out is not None else None,
                    planes_out[0].numel() if planes_out is not None else 0,
                    mask_planes.data_ptr() if mask_planes is not None else None,
                    mask_bits.data_ptr() if mask_bits is not None else None,
                    mask_bits_out.data_ptr() if mask_bits_out is not None else None,
                    int(planes_lrelu is not None),
                    float(planes_lrelu) if planes_lrelu is not None else 0.0)


# ------------------------------------------------------------------------------------------
# raw (non-differentiable) wrappers
# ------------------------------------------------------------------------------------------
def gemm_nn(ga, W, epi, M, N, K):
    lib = _lib.load()
    _lib.check(lib.ssb_gemm_nn(ctypes.byref(ga), W.data_ptr(), W.stride(0), ctypes.byref(epi), M,
                               N, K, _stream()))


def gemm_nt(ga, W, Cb, tapmap, epi, M, N, K):
    lib = _lib.load()
    t = list(tapmap) + [0, 0, 0]
    _lib.check(lib.ssb_gemm_nt(ctypes.byref(ga), W.data_ptr(), W.stride(0), Cb, t[0], t[1], t[2],
                               ctypes.byref(epi), M, N, K, _stream()))


def gemm_tn(ga, G, dW, M, N, K, accumulate=False):
    lib = _lib.load()
    _lib.check(lib.ssb_gemm_tn(ctypes.byref(ga), G.data_ptr(), G.stride(0), dW.data_ptr(),
                               dW.stride(0), int(accumulate), M, N, K, _stream()))


def colsum(x2d, out=None, accumulate=False):
    lib = _lib.load()
    rows, C = x2d.shape
    if out is None:
        out = torch.empty(C, dtype=_f32, device=x2d.device)
    nb = lib.ssb_col_partials_bytes(rows, C)
    ws = _ws(nb, x2d.device)
    _lib.check(lib.ssb_colsum(x2d.data_ptr(), rows, C, out.data_ptr(), int(accumulate),
                              ws.data_ptr(), ws.numel(), _stream()))
    return out


# ------------------------------------------------------------------------------------------
# tcgen05 path: bf16 hi/lo split planes + tensor-core GEMMs (csrc/gemm_tc.cu)
# ------------------------------------------------------------------------------------------


def colsum_planes(planes, out=None, accumulate=False):
    """Column sums of a (2, rows, C) split-plane tensor (x = hi + lo)."""
    lib = _lib.load()
    _, rows, C = planes.shape
    if out is None:
        out = torch.empty(C, dtype=_f32, device=planes.device)
    ws = _ws(lib.ssb_col_partials_bytes(rows, C), planes.device)
    _lib.check(lib.ssb_colsum_planes(planes.data_ptr(), rows * C, rows, C, out.data_ptr(),
                                     int(accumulate), ws.data_ptr(), ws.numel(), _stream()))
    return out


# ------------------------------------------------------------------------------------------
# gradient sinks: with a training.GradientBucket the .grad of every parameter is a fixed view of
# one flat buffer.  Backward bodies that own a parameter gradient may then ACCUMULATE it straight
# into that view (the weight-gradient GEMM's epilogue / split-K atomics, the column-sum finalize)
# and return None to autograd, instead of materialising a temporary that AccumulateGrad adds with
# one more launch and pass per parameter.
# ------------------------------------------------------------------------------------------
_sinks = {}


def register_grad_sinks(params):
    """params: iterable of nn.Parameter whose .grad are persistent views (GradientBucket)."""
    import weakref
    for p in params:
        _sinks[id(p)] = (weakref.ref(p), p.grad)


def _sink(t):
    """The registered gradient view of parameter object `t`, if it is still t.grad."""
    ent = _sinks.get(id(t))
    if ent is None or ent[0]() is not t or not t.requires_grad:
        return None
    g = t.grad
    if g is None or g.data_ptr() != ent[1].data_ptr():
        return None
    return ent[1]


def attach_planes(t, planes):
    """Remember that `planes` (2, *t.shape) are the split planes of t's CURRENT contents (a
    producer kernel wrote both).  Consumers pick them up through planes_of()."""
    t._ssb_planes = (t._version, planes)
    return t


def planes_of(t):
    """Split planes of t: the ones its producer already wrote, if t was not modified since
    (in-place autograd accumulation bumps _version), else a split pass."""
    ent = getattr(t, "_ssb_planes", None)
    if ent is not None and ent[0] == t._version and tuple(ent[1].shape[1:]) == tuple(t.shape):
        return ent[1]
    return split_planes(t)


def _want_planes(rows, C):
    """Would a tcgen05 GEMM consume a (rows, C) activation?  (else planes would be wasted writes)"""
    return _tc_enabled() and rows >= 64 and C % 64 == 0


def split_planes_t(x2d):
    """fp32 (rows, cols) -> bf16 planes (2, cols, rows) of the TRANSPOSE (one fused pass)."""
    lib = _lib.load()
    _chk(x2d, "x")
    rows, cols = x2d.shape
    out = torch.empty((2, cols, rows), dtype=torch.bfloat16, device=x2d.device)
    _lib.check(lib.ssb_split_bf16_t(x2d.data_ptr(), rows, cols, out.data_ptr(), _stream()))
    return out


def split_planes(x):
    """fp32 tensor -> bf16 tensor (2, *x.shape): plane 0 = hi = bf16(x), plane 1 = lo = bf16(x - hi)."""
    lib = _lib.load()
    _chk(x, "x")
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.ssb_split_bf16(x.data_ptr(), x.numel(), out.data_ptr(), _stream()))
    return out


def tc_operand_plain(planes, M, K):
    """(2, M, K) split planes as a plain row-major matrix operand."""
    return TcOperand(planes.data_ptr(), M * K, M * K, 1, M, M, K, K, 1, 0, 0)


def tc_operand_conv(planes, B, L, C, rows_out, stride, taps_step, off):
    """(2, B, L, C) split planes as the im2col operand of a k3/k1 convolution."""
    return TcOperand(planes.data_ptr(), B * L * C, L * C, B, rows_out, L, C, C, stride, taps_step,
                     off)


_streamk_ws = {}     # device index -> the registered stream-K workspace (kept alive here)


def _ensure_streamk(lib, device):
    """Register the per-device stream-K workspace of ssb_gemm_tc_kmajor (include/ssb.h) on first use:
    zero-filled once, owned by this module.  All tcgen05 GEMMs of a device share it, so they must be
    stream-ordered (one execution stream per process).  Opt-in (SSB_STREAMK=1): measured slower than
    the classic schedule on every shape of the cfg-1 step and the vocoder (csrc/gemm_tc.cu)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _streamk_ws or torch.cuda.is_current_stream_capturing():
        return
    if os.environ.get("SSB_STREAMK", "0") != "1":
        _streamk_ws[idx] = None
        return
    with torch.cuda.device(idx):
        ws = torch.zeros(int(lib.ssb_gemm_tc_streamk_workspace_bytes()), dtype=torch.uint8, device=device)
        _lib.check(lib.ssb_gemm_tc_set_streamk_workspace(ws.data_ptr(), ws.numel()))
    _streamk_ws[idx] = ws


def gemm_tc_kmajor(opA, Bplanes, N, K, epi):
    lib = _lib.load()
    _ensure_streamk(lib, Bplanes.device)
    _lib.check(lib.ssb_gemm_tc_kmajor(ctypes.byref(opA), Bplanes.data_ptr(), N, K,
                                      ctypes.byref(epi), _stream()))


def gemm_tc_wgrad(opX, Gplanes, N, K, dW, accumulate=False, group=None):
    """group = (lddw, group_w, group_stride): grouped output columns (ssb.h), dW a raw base tensor."""
    lib = _lib.load()
    ld, gw, gs = group if group is not None else (dW.stride(0), 0, 0)
    _lib.check(lib.ssb_gemm_tc_wgrad(ctypes.byref(opX), Gplanes.data_ptr(), Gplanes[0].numel(), N,
                                     K, dW.data_ptr(), ld, gw, gs, int(accumulate), _stream()))


# ------------------------------------------------------------------------------------------
# engine selection: tcgen05 (gemm_tc.cu) when the shape meets its tiling constraints, the
# CUDA-core engine (gemm_simt.cu) otherwise (thin first conv C=8, K=dh=96 positional GEMMs,
# unit-test sized models).  SSB_GEMM=simt forces the CUDA-core engine (A/B testing).
# ------------------------------------------------------------------------------------------
def _tc_enabled():
    return os.environ.get("SSB_GEMM", "tc") != "simt"


def _tc_fwd_ok(M, N, K, C=None):
    C = K if C is None else C
    return _tc_enabled() and K % 64 == 0 and C % 64 == 0 and N % 4 == 0 and M >= 64


def _tc_wgrad_ok(M, N, K, C=None):
    C = K if C is None else C
    return _tc_enabled() and K % 128 == 0 and C % 128 == 0 and N % 8 == 0 and M >= 64


def mm_fwd(x, Wg, epi_kwargs, out, M, N, K, xp=None):
    """out[M,N] = epi(x[M,K] @ Wg[K,N]).  Returns the split planes of x if they were made."""
    scat = _scatter_plain(out.data_ptr(), M, N)
    if _tc_fwd_ok(M, N, K):
        if xp is None:
            xp = planes_of(x)
        wp = split_planes(Wg.t().contiguous())                    # [N][K], K contiguous
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), wp, N, K, _epi(scat, **epi_kwargs))
        return xp
    gemm_nn(_gather_plain(x.data_ptr(), M, K, K), Wg, _epi(scat, **epi_kwargs), M, N, K)
    return None


def mm_dgrad(dy, Wg, epi_kwargs, out, M, N, K, dyp=None):
    """out[M,K] = epi(dy[M,N] @ Wg[K,N]^T)."""
    scat = _scatter_plain(out.data_ptr(), M, K)
    if _tc_fwd_ok(M, K, N):
        if dyp is None:
            dyp = planes_of(dy)
        gemm_tc_kmajor(tc_operand_plain(dyp, M, N), split_planes(Wg), K, N, _epi(scat, **epi_kwargs))
        return dyp
    gemm_nt(_gather_plain(dy.data_ptr(), M, N, N), Wg, N, (0,), _epi(scat, **epi_kwargs), M, K, N)
    return dyp


def mm_wgrad(x, dy, dW, M, N, K, xp=None, dyp=None):
    """dW[K,N] = x[M,K]^T @ dy[M,N]."""
    if _tc_wgrad_ok(M, N, K):
        if xp is None:
            xp = planes_of(x)
        if dyp is None:
            dyp = planes_of(dy)
        gemm_tc_wgrad(tc_operand_plain(xp, M, K), dyp, N, K, dW)
        return dyp
    gemm_tn(_gather_plain(x.data_ptr(), M, K, K), dy, dW, M, N, K)
    return dyp


# ------------------------------------------------------------------------------------------
# Linear:  y = x @ Wg + b     (x: (M, K), Wg: (K, N))
# ------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wg, bias):
        _chk(x, "x"), _chk(Wg, "Wg")
        M, K = x.shape
        N = Wg.shape[1]
        y = torch.empty((M, N), dtype=_f32, device=x.device)
        xp = mm_fwd(x, Wg, dict(bias=bias), y, M, N, K)
        # keep the split planes of x (same bytes as x) so backward does not split it again
        keep_planes = xp is not None and ctx.needs_input_grad[1] and _tc_wgrad_ok(M, N, K)
        ctx.save_for_backward(x, Wg, xp if keep_planes else None)
        ctx.has_bias = bias is not None
        ctx.sink_b = _sink(bias) if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wg, xp = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = Wg.shape[1]
        dx = dW = db = None
        dyp = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            dyp = mm_dgrad(dy, Wg, {}, dx, M, N, K)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(Wg)
            mm_wgrad(x, dy, dW, M, N, K, xp=xp, dyp=dyp)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if ctx.sink_b is not None:
                colsum(dy, out=ctx.sink_b, accumulate=True)
            else:
                db = colsum(dy)
        return dx, dW, db


def linear(x2d, Wg, bias=None):
    """y = x @ Wg + bias.  Output widths that are not a multiple of 4 (the 38-way CTC head of
    recognition_model.py:66) are zero-padded to the kernels' float4 granularity and sliced."""
    N = Wg.shape[1]
    if N % 4:
        pad = 4 - N % 4
        Wp = torch.nn.functional.pad(Wg, (0, pad)).contiguous()
        bp = torch.nn.functional.pad(bias, (0, pad)) if bias is not None else None
        return _LinearFn.apply(x2d, Wp, bp)[:, :N].contiguous()
    return _LinearFn.apply(x2d, Wg, bias)


# ------------------------------------------------------------------------------------------
# FFN:  y = dropout(relu(x @ W1 + b1)) @ W2 + b2          transformer.py:57
# ------------------------------------------------------------------------------------------
class _FFNFn(torch.autograd.Function):
    """On the tcgen05 engine the hidden activation h and its gradient dh exist only as bf16
    split planes: the producing GEMM's epilogue writes the operand format of the consuming
    GEMMs directly (no fp32 copy, no split pass), the ReLU/dropout mask of the backward is read
    off h's hi plane, and the bias gradient sums the planes."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, p, seed, site):
        _chk(x, "x"), _chk(W1, "W1"), _chk(W2, "W2")
        M, K = x.shape
        F_ = W1.shape[1]
        N = W2.shape[1]
        dev = x.device
        y = torch.empty((M, N), dtype=_f32, device=dev)
        ctx.p = p
        ctx.planes = (_tc_fwd_ok(M, F_, K) and _tc_fwd_ok(M, N, F_) and _tc_wgrad_ok(M, F_, K)
                      and _tc_wgrad_ok(M, N, F_) and _tc_fwd_ok(M, K, F_) and _tc_fwd_ok(M, F_, N))
        if ctx.planes:
            xp = planes_of(x)
            hp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
            gemm_tc_kmajor(tc_operand_plain(xp, M, K), split_planes(W1.t().contiguous()), F_, K,
                           _epi(_scatter_plain(None, M, F_), bias=b1, relu=1, drop_p=p, seed=seed,
                                site=site, planes_out=hp))
            gemm_tc_kmajor(tc_operand_plain(hp, M, F_), split_planes(W2.t().contiguous()), N, F_,
                           _epi(_scatter_plain(y.data_ptr(), M, N), bias=b2))
            ctx.save_for_backward(W1, W2, xp, hp)
            return y
        h = torch.empty((M, F_), dtype=_f32, device=dev)
        xp = mm_fwd(x, W1, dict(bias=b1, relu=1, drop_p=p, seed=seed, site=site), h, M, F_, K)
        hp = mm_fwd(h, W2, dict(bias=b2), y, M, N, F_)
        ctx.save_for_backward(x, W1, W2, h, xp if _tc_wgrad_ok(M, F_, K) else None,
                              hp if _tc_wgrad_ok(M, N, F_) else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        scale = 1.0 / (1.0 - ctx.p) if ctx.p > 0 else 1.0
        if ctx.planes:
            W1, W2, xp, hp = ctx.saved_tensors
            M, K = xp.shape[1:]
            F_ = W1.shape[1]
            N = W2.shape[1]
            dev = dy.device
            dyp = planes_of(dy)
            dW2 = torch.empty_like(W2)
            gemm_tc_wgrad(tc_operand_plain(hp, M, F_), dyp, N, F_, dW2)
            db2 = colsum(dy)
            # dh = (dy @ W2^T) * (h > 0) / (1 - p): relu and dropout masks both read off h
            dhp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
            gemm_tc_kmajor(tc_operand_plain(dyp, M, N), split_planes(W2), F_, N,
                           _epi(_scatter_plain(None, M, F_), mask_planes=hp[0], mask_scale=scale,
                                planes_out=dhp))
            dW1 = torch.empty_like(W1)
            gemm_tc_wgrad(tc_operand_plain(xp, M, K), dhp, F_, K, dW1)
            db1 = colsum_planes(dhp)
            dx = torch.empty((M, K), dtype=_f32, device=dev)
            gemm_tc_kmajor(tc_operand_plain(dhp, M, F_), split_planes(W1), K, F_,
                           _epi(_scatter_plain(dx.data_ptr(), M, K)))
            return dx, dW1, db1, dW2, db2, None, None, None
        x, W1, W2, h, xp, hp = ctx.saved_tensors
        M, K = x.shape
        F_ = W1.shape[1]
        N = W2.shape[1]
        dW2 = torch.empty_like(W2)
        dyp = mm_wgrad(h, dy, dW2, M, N, F_, xp=hp)
        db2 = colsum(dy)
        dh = torch.empty_like(h)
        mm_dgrad(dy, W2, dict(mask_src=h, mask_scale=scale), dh, M, N, F_, dyp=dyp)
        dW1 = torch.empty_like(W1)
        dhp = mm_wgrad(x, dh, dW1, M, F_, K, xp=xp)
        db1 = colsum(dh)
        dx = torch.empty_like(x)
        mm_dgrad(dh, W1, {}, dx, M, F_, K, dyp=dhp)
        return dx, dW1, db1, dW2, db2, None, None, None


class _FFNNativeFn(torch.autograd.Function):
    """_FFNFn's split-plane path on nn.Linear-layout weights (w1: (F, K), w2: (N, F)), i.e. on the
    parameters themselves: forward splits them as they lie, the data gradients split their
    transpose in one fused pass, and the weight gradients come out of the tensor-core GEMM in
    parameter layout (dW^T = dy^T x) - straight into the gradient bucket when one is registered."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, p, seed, site, w1f=None, w1t=None, w2f=None, w2t=None):
        """w1f / w2f: [F][K] / [N][F] planes of the weights as they lie, w1t / w2t: planes of their
        transposes (WeightPlanes arena); None -> derived here with a split pass."""
        _chk(x, "x"), _chk(w1, "w1"), _chk(w2, "w2")
        M, K = x.shape
        F_, N = w1.shape[0], w2.shape[0]
        dev = x.device
        xp = planes_of(x)
        hp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), w1f if w1f is not None else split_planes(w1), F_, K,
                       _epi(_scatter_plain(None, M, F_), bias=b1, relu=1, drop_p=p, seed=seed,
                            site=site, planes_out=hp))
        y = torch.empty((M, N), dtype=_f32, device=dev)
        gemm_tc_kmajor(tc_operand_plain(hp, M, F_), w2f if w2f is not None else split_planes(w2), N, F_,
                       _epi(_scatter_plain(y.data_ptr(), M, N), bias=b2))
        ctx.save_for_backward(w1, w2, xp, hp, w1t, w2t)
        ctx.p = p
        ctx.sinks = tuple(_sink(t) for t in (w1, b1, w2, b2))
        return y

    @staticmethod
    def backward(ctx, dy):
        w1, w2, xp, hp, w1t, w2t = ctx.saved_tensors
        s_w1, s_b1, s_w2, s_b2 = ctx.sinks
        dy = dy.contiguous()
        M, K = xp.shape[1:]
        F_, N = w1.shape[0], w2.shape[0]
        dev = dy.device
        scale = 1.0 / (1.0 - ctx.p) if ctx.p > 0 else 1.0
        dyp = planes_of(dy)
        dW2 = s_w2 if s_w2 is not None else torch.empty_like(w2)          # (N, F) = dy^T h
        gemm_tc_wgrad(tc_operand_plain(dyp, M, N), hp, F_, N, dW2, accumulate=s_w2 is not None)
        db2 = colsum(dy, out=s_b2, accumulate=s_b2 is not None)
        dhp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
        gemm_tc_kmajor(tc_operand_plain(dyp, M, N), w2t if w2t is not None else split_planes_t(w2), F_, N,
                       _epi(_scatter_plain(None, M, F_), mask_planes=hp[0], mask_scale=scale,
                            planes_out=dhp))
        dW1 = s_w1 if s_w1 is not None else torch.empty_like(w1)          # (F, K) = dh^T x
        gemm_tc_wgrad(tc_operand_plain(dhp, M, F_), xp, K, F_, dW1, accumulate=s_w1 is not None)
        db1 = colsum_planes(dhp, out=s_b1, accumulate=s_b1 is not None)
        dx = torch.empty((M, K), dtype=_f32, device=dev)
        gemm_tc_kmajor(tc_operand_plain(dhp, M, F_), w1t if w1t is not None else split_planes_t(w1), K, F_,
                       _epi(_scatter_plain(dx.data_ptr(), M, K)))
        return (dx, None if s_w1 is not None else dW1, None if s_b1 is not None else db1,
                None if s_w2 is not None else dW2, None if s_b2 is not None else db2,
                None, None, None, None, None, None, None)


def ffn_native(x2d, w1, b1, w2, b2, p=0.0, seed=0, site=0, wp=None):
    """FFN on nn.Linear-layout weights (linear1.weight (F, K), linear2.weight (N, F)); `wp`: the
    model's WeightPlanes arena (weight operands without per-use split passes)."""
    M, K = x2d.shape
    F_, N = w1.shape[0], w2.shape[0]
    ok = (_tc_fwd_ok(M, F_, K) and _tc_fwd_ok(M, N, F_) and _tc_fwd_ok(M, K, F_)
          and _tc_fwd_ok(M, F_, N) and _tc_wgrad_ok(M, F_, N) and _tc_wgrad_ok(M, K, F_)
          and K % 8 == 0 and b1 is not None and b2 is not None)
    if ok:
        pl = [None] * 4
        if wp is not None:
            pl = [wp.get(w1, "f"), wp.get(w1, "b"), wp.get(w2, "f"), wp.get(w2, "b")]
            if any(t is None for t in pl):
                pl = [None] * 4
        return _FFNNativeFn.apply(x2d, w1, b1, w2, b2, float(p), int(seed), int(site), *pl)
    return ffn(x2d, w1.t().contiguous(), b1, w2.t().contiguous(), b2, p, seed, site)


def ffn(x2d, W1g, b1, W2g, b2, p=0.0, seed=0, site=0):
    return _FFNFn.apply(x2d, W1g, b1, W2g, b2, float(p), int(seed), int(site))


# ------------------------------------------------------------------------------------------
# Weight-plane variants: the same ops on the PARAMETERS themselves (reference shapes), with every
# weight operand layout taken from the model's WeightPlanes arena (weights.py, csrc/prep.cu: one
# launch per step) instead of `.t().contiguous()` / cat / split passes per use, and with weight
# gradients written straight into the gradient bucket where the GEMM can produce the parameter's
# own layout.
# ------------------------------------------------------------------------------------------
class _LinearWFn(torch.autograd.Function):
    """y = x @ weight^T + bias on an nn.Linear-layout weight (N, K).  wf: [N][K] planes (forward
    operand), wb: [K][N] planes (data-gradient operand); both required."""

    @staticmethod
    def forward(ctx, x, weight, bias, wf, wb):
        _chk(x, "x")
        M, K = x.shape
        N = weight.shape[0]
        y = torch.empty((M, N), dtype=_f32, device=x.device)
        xp = planes_of(x)
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), wf, N, K,
                       _epi(_scatter_plain(y.data_ptr(), M, N), bias=bias))
        native = K % 8 == 0 and N % 128 == 0          # dW^T = dy^T x in parameter layout
        ctx.save_for_backward(x, weight, wb, xp if ctx.needs_input_grad[1] else None)
        ctx.native = native
        ctx.sink_w = _sink(weight) if native else None
        ctx.sink_b = _sink(bias) if bias is not None else None
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, wb, xp = ctx.saved_tensors
        dy = dy.contiguous()
        M, K = x.shape
        N = weight.shape[0]
        dx = dW = db = None
        dyp = planes_of(dy)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            gemm_tc_kmajor(tc_operand_plain(dyp, M, N), wb, K, N, _epi(_scatter_plain(dx.data_ptr(), M, K)))
        if ctx.needs_input_grad[1]:
            if ctx.native:
                dW = ctx.sink_w if ctx.sink_w is not None else torch.empty_like(weight)
                gemm_tc_wgrad(tc_operand_plain(dyp, M, N), xp, K, N, dW, accumulate=ctx.sink_w is not None)
                if ctx.sink_w is not None:
                    dW = None
            elif _tc_wgrad_ok(M, N, K):
                dWg = torch.empty((K, N), dtype=_f32, device=x.device)
                gemm_tc_wgrad(tc_operand_plain(xp, M, K), dyp, N, K, dWg)
                dW = dWg.t()
            else:
                dWg = torch.empty((K, N), dtype=_f32, device=x.device)
                gemm_tn(_gather_plain(x.data_ptr(), M, K, K), dy, dWg, M, N, K)
                dW = dWg.t()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if ctx.sink_b is not None:
                colsum(dy, out=ctx.sink_b, accumulate=True)
            else:
                db = colsum(dy)
        return dx, dW, db, None, None


def linear_w(x2d, lin, wp):
    """nn.Linear `lin` applied to token-major x2d through the WeightPlanes arena `wp` when its
    shape is tensor-core eligible both ways; otherwise the round-1 path."""
    w = lin.weight
    N, K = w.shape
    wf = wp.get(w, "f") if wp is not None else None
    wb = wp.get(w, "b") if wp is not None else None
    M = x2d.shape[0]
    if wf is not None and wb is not None and _tc_fwd_ok(M, N, K) and _tc_fwd_ok(M, K, N):
        return _LinearWFn.apply(x2d, w, lin.bias, wf, wb)
    return linear(x2d, w.t().contiguous(), lin.bias)


class _HeadsFn(torch.autograd.Function):
    """(w_out(x), w_aux(x)) of architecture.py:81-84 as ONE GEMM on the stacked weight (N1 + N2
    outputs; 80 + 48 = 128 in the transduction model), so that forward, data gradient and weight
    gradient all run on the tensor cores.  hf: [N1+N2][K] planes, hb: [K][N1+N2] planes."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, hf, hb):
        _chk(x, "x")
        M, K = x.shape
        N1, N2 = w1.shape[0], w2.shape[0]
        N = N1 + N2
        bias = torch.cat([b1, b2])
        y = torch.empty((M, N), dtype=_f32, device=x.device)
        xp = planes_of(x)
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), hf, N, K, _epi(_scatter_plain(y.data_ptr(), M, N), bias=bias))
        ctx.save_for_backward(xp, hb)
        ctx.split = (N1, N2)
        return y[:, :N1].contiguous(), y[:, N1:].contiguous()

    @staticmethod
    def backward(ctx, d1, d2):
        xp, hb = ctx.saved_tensors
        N1, N2 = ctx.split
        N = N1 + N2
        M, K = xp.shape[1:]
        dy = torch.cat([d1, d2], dim=1)
        dyp = split_planes(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=_f32, device=dy.device)
            gemm_tc_kmajor(tc_operand_plain(dyp, M, N), hb, K, N, _epi(_scatter_plain(dx.data_ptr(), M, K)))
        dWg = torch.empty((K, N), dtype=_f32, device=dy.device)
        gemm_tc_wgrad(tc_operand_plain(xp, M, K), dyp, N, K, dWg)
        db = colsum(dy)
        return dx, dWg[:, :N1].t(), db[:N1], dWg[:, N1:].t(), db[N1:], None, None


class _QKVFn(torch.autograd.Function):
    """qkv = x @ [Wq | Wk | Wv] for the per-head weights w_j (H, D, dh) of transformer.py:71-74
    (no bias).  wf: [3D][D] planes, wb: [D][3D] planes of the fused matrix."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, wf, wb):
        _chk(x, "x")
        M, D = x.shape
        y = torch.empty((M, 3 * D), dtype=_f32, device=x.device)
        xp = planes_of(x)
        gemm_tc_kmajor(tc_operand_plain(xp, M, D), wf, 3 * D, D, _epi(_scatter_plain(y.data_ptr(), M, 3 * D)))
        ctx.save_for_backward(xp, wb)
        ctx.wshape = tuple(wq.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, wb = ctx.saved_tensors
        H, D, dh = ctx.wshape
        dy = dy.contiguous()
        M = dy.shape[0]
        dyp = planes_of(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, D), dtype=_f32, device=dy.device)
            gemm_tc_kmajor(tc_operand_plain(dyp, M, 3 * D), wb, D, 3 * D,
                           _epi(_scatter_plain(dx.data_ptr(), M, D)))
        dWg = torch.empty((D, 3 * D), dtype=_f32, device=dy.device)
        gemm_tc_wgrad(tc_operand_plain(xp, M, D), dyp, 3 * D, D, dWg)
        # column (j, h, a) of dWg is d w_j[h, :, a]
        g = [dWg[:, j * D:(j + 1) * D].view(D, H, dh).permute(1, 0, 2) for j in range(3)]
        return dx, g[0], g[1], g[2], None, None


class _OutProjFn(torch.autograd.Function):
    """y = o @ Wo with w_o (H, dh, D) = the GEMM-layout matrix [K][N] (transformer.py:111).
    wf: [N][K] planes, wb: [K][N] planes.  dWo comes out of the GEMM in the parameter's layout."""

    @staticmethod
    def forward(ctx, o, w_o, wf, wb):
        _chk(o, "o")
        M, K = o.shape
        N = w_o.shape[2]
        y = torch.empty((M, N), dtype=_f32, device=o.device)
        op = planes_of(o)
        gemm_tc_kmajor(tc_operand_plain(op, M, K), wf, N, K, _epi(_scatter_plain(y.data_ptr(), M, N)))
        ctx.save_for_backward(op, wb, w_o)
        ctx.sink = _sink(w_o)
        return y

    @staticmethod
    def backward(ctx, dy):
        op, wb, w_o = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        K = op.shape[2]
        dyp = planes_of(dy)
        do = None
        if ctx.needs_input_grad[0]:
            do = torch.empty((M, K), dtype=_f32, device=dy.device)
            gemm_tc_kmajor(tc_operand_plain(dyp, M, N), wb, K, N, _epi(_scatter_plain(do.data_ptr(), M, K)))
        if ctx.sink is not None:
            gemm_tc_wgrad(tc_operand_plain(op, M, K), dyp, N, K, ctx.sink.view(K, N), accumulate=True)
            return do, None, None, None
        dW = torch.empty((K, N), dtype=_f32, device=dy.device)
        gemm_tc_wgrad(tc_operand_plain(op, M, K), dyp, N, K, dW)
        return do, dW.view_as(w_o), None, None


# ------------------------------------------------------------------------------------------
# Conv1d, channels-last: x (B, L, Cin) -> (B, Lout, Cout); kernel 3 pad 1 (stride 1|2) or
# kernel 1 pad 0 stride 2.  Wg: (ksize*Cin, Cout), row = tap*Cin + ci.   architecture.py:18-24
# ------------------------------------------------------------------------------------------
def _conv_gather(x, L, Cin, Lout, ksize, stride):
    off = -1 if ksize == 3 else 0
    return Gather(x.data_ptr(), L * Cin, Lout, Cin, L, Cin, stride, 1, off)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wg, bias, ksize, stride):
        _chk(x, "x"), _chk(Wg, "Wg")
        B, L, Cin = x.shape
        Cout = Wg.shape[1]
        assert Wg.shape[0] == ksize * Cin and ksize in (1, 3) and stride in (1, 2)
        Lout = (L - 1) // stride + 1
        y = torch.empty((B, Lout, Cout), dtype=_f32, device=x.device)
        M = B * Lout
        off = -1 if ksize == 3 else 0
        xp = None
        if _tc_fwd_ok(M, Cout, ksize * Cin, Cin):
            xp = planes_of(x)
            wp = split_planes(Wg.t().contiguous())               # [Cout][(tap, ci)]
            gemm_tc_kmajor(tc_operand_conv(xp, B, L, Cin, Lout, stride, 1, off), wp, Cout,
                           ksize * Cin, _epi(Scatter(y.data_ptr(), Lout * Cout, Lout, Cout, 1, 0),
                                             bias=bias))
        else:
            gemm_nn(_conv_gather(x, L, Cin, Lout, ksize, stride), Wg,
                    _epi(_scatter_plain(y.data_ptr(), M, Cout), bias=bias), M, Cout, ksize * Cin)
        keep = xp is not None and ctx.needs_input_grad[1] and _tc_wgrad_ok(M, Cout, ksize * Cin, Cin)
        ctx.save_for_backward(x, Wg, xp if keep else None)
        ctx.cfg = (ksize, stride, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wg, xp_saved = ctx.saved_tensors
        ksize, stride, has_bias = ctx.cfg
        dy = dy.contiguous()
        B, L, Cin = x.shape
        _, Lout, Cout = dy.shape
        M = B * Lout
        K = ksize * Cin
        dy2 = dy.view(M, Cout)
        off = -1 if ksize == 3 else 0
        dx = dW = db = None
        dyp = None
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(Wg)
            if _tc_wgrad_ok(M, Cout, K, Cin):
                dyp = planes_of(dy)
                xpl = xp_saved if xp_saved is not None else planes_of(x)
                gemm_tc_wgrad(tc_operand_conv(xpl, B, L, Cin, Lout, stride, 1, off), dyp, Cout, K, dW)
            else:
                gemm_tn(_conv_gather(x, L, Cin, Lout, ksize, stride), dy2, dW, M, Cout, K)
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        if ctx.needs_input_grad[0]:
            dyptr = dy.data_ptr()
            W3 = Wg.view(ksize, Cin, Cout)
            use_tc = _tc_fwd_ok(B * L // max(stride, 1), Cin, Cout, Cout)
            if use_tc and dyp is None:
                dyp = planes_of(dy)

            def run(rows, taps, tap_off, tapmap, d_t, d_off, dst, accumulate=0):
                """dst rows (b, t*d_t + d_off) = sum_j dy[b, t + j + tap_off] . W_tapmap[j]^T"""
                out = Scatter(dst.data_ptr(), L * Cin, rows, Cin, d_t, d_off)
                if use_tc:
                    Bd = torch.stack([W3[t] for t in tapmap], dim=1).reshape(Cin, taps * Cout)
                    gemm_tc_kmajor(tc_operand_conv(dyp, B, Lout, Cout, rows, 1, 1, tap_off),
                                   split_planes(Bd.contiguous()), Cin, taps * Cout,
                                   _epi(out, accumulate=accumulate))
                else:
                    ga = Gather(dyptr, Lout * Cout, rows, Cout, Lout, Cout, 1, 1, tap_off)
                    gemm_nt(ga, Wg, Cout, tapmap, _epi(out, accumulate=accumulate), B * rows, Cin,
                            taps * Cout)

            if ksize == 3 and stride == 1:
                dx = torch.empty_like(x)
                run(L, 3, -1, (2, 1, 0), 1, 0, dx)
            elif ksize == 3 and stride == 2:
                dx = torch.empty_like(x)
                ne, no = (L + 1) // 2, L // 2
                run(ne, 1, 0, (1,), 2, 0, dx)            # even rows: dy[t] . W_1^T
                if no > 0:
                    run(no, 2, 0, (2, 0), 2, 1, dx)      # odd rows: dy[t] . W_2^T + dy[t+1] . W_0^T
            elif ksize == 1 and stride == 2:
                dx = torch.zeros_like(x)
                run((L + 1) // 2, 1, 0, (0,), 2, 0, dx)
            else:
                raise NotImplementedError("conv backward for this (ksize, stride)")
        return dx, dW, db, None, None


def conv1d_cl(x, Wg, bias, ksize, stride):
    return _ConvFn.apply(x, Wg, bias, ksize, stride)


class _ConvWFn(torch.autograd.Function):
    """_ConvFn on the nn.Conv1d parameter itself (Cout, Cin, k) with arena operands: wf =
    [Cout][(tap, ci)] planes; wd = per-tap-subset data-gradient planes [Cin][(j, co)] in the
    order the backward uses them (k3 s1: (2,1,0); k3 s2: (1,), (2,0); k1 s2: (0,))."""

    @staticmethod
    def forward(ctx, x, weight, bias, ksize, stride, wf, wd0, wd1, zero_bias_grad=False):
        _chk(x, "x")
        B, L, Cin = x.shape
        Cout = weight.shape[0]
        Lout = (L - 1) // stride + 1
        y = torch.empty((B, Lout, Cout), dtype=_f32, device=x.device)
        off = -1 if ksize == 3 else 0
        xp = planes_of(x)
        gemm_tc_kmajor(tc_operand_conv(xp, B, L, Cin, Lout, stride, 1, off), wf, Cout, ksize * Cin,
                       _epi(Scatter(y.data_ptr(), Lout * Cout, Lout, Cout, 1, 0), bias=bias))
        ctx.save_for_backward(x, xp if ctx.needs_input_grad[1] else None, wd0, wd1)
        # a bias in front of a TRAINING-mode BatchNorm has an exactly zero gradient (the batch
        # mean absorbs it): the caller says so and the column sum of dy is skipped
        ctx.cfg = (ksize, stride, bias is not None and not zero_bias_grad, tuple(weight.shape))
        ctx.sink_b = _sink(bias) if bias is not None else None
        ctx.zero_b = bias is not None and zero_bias_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x, xp, wd0, wd1 = ctx.saved_tensors
        ksize, stride, has_bias, wshape = ctx.cfg
        Cout, Cin, _ = wshape
        dy = dy.contiguous()
        B, L, _ = x.shape
        Lout = dy.shape[1]
        M = B * Lout
        K = ksize * Cin
        off = -1 if ksize == 3 else 0
        dyp = planes_of(dy)
        dx = dW = db = None
        if ctx.needs_input_grad[1]:
            dWg = torch.empty((K, Cout), dtype=_f32, device=x.device)
            xpl = xp if xp is not None else planes_of(x)
            gemm_tc_wgrad(tc_operand_conv(xpl, B, L, Cin, Lout, stride, 1, off), dyp, Cout, K, dWg)
            dW = dWg.view(ksize, Cin, Cout).permute(2, 1, 0)       # -> (Cout, Cin, k)
        if has_bias and ctx.needs_input_grad[2]:
            if ctx.sink_b is not None:
                colsum(dy.view(M, Cout), out=ctx.sink_b, accumulate=True)
            else:
                db = colsum(dy.view(M, Cout))
        elif ctx.zero_b and ctx.needs_input_grad[2] and ctx.sink_b is None:
            db = torch.zeros(Cout, dtype=_f32, device=x.device)   # exact zero; a bucket already holds it
        if ctx.needs_input_grad[0]:
            def run(rows, taps, tap_off, planes, d_t, d_off, dst):
                gemm_tc_kmajor(tc_operand_conv(dyp, B, Lout, Cout, rows, 1, 1, tap_off), planes, Cin,
                               taps * Cout, _epi(Scatter(dst.data_ptr(), L * Cin, rows, Cin, d_t, d_off)))
            if ksize == 3 and stride == 1:
                dx = torch.empty_like(x)
                run(L, 3, -1, wd0, 1, 0, dx)
            elif ksize == 3 and stride == 2:
                dx = torch.empty_like(x)
                ne, no = (L + 1) // 2, L // 2
                run(ne, 1, 0, wd0, 2, 0, dx)             # even rows: dy[t] . W_1^T
                if no > 0:
                    run(no, 2, 0, wd1, 2, 1, dx)         # odd rows: dy[t] . W_2^T + dy[t+1] . W_0^T
            else:
                dx = torch.zeros_like(x)
                run((L + 1) // 2, 1, 0, wd0, 2, 0, dx)
        return dx, dW, db, None, None, None, None, None, None


class _ConvPairWFn(torch.autograd.Function):
    """The two stride-2 convolutions that read a ResBlock's input, conv1 (k3, p1) and
    residual_path (k1) of architecture.py:18,24, as ONE autograd node: their input gradients land
    in one buffer (the 1x1 data-gradient GEMM accumulates into the even rows the k3 one wrote)
    instead of a zero-filled second tensor summed by the autograd engine (a fill and an add over
    the block input per ResBlock and step).  Operands as in _ConvWFn."""

    @staticmethod
    def forward(ctx, x, w1, b1, wr, br, wf1, wd1a, wd1b, wfr, wdr, zero_bias_grad):
        _chk(x, "x")
        B, L, Cin = x.shape
        Cout = w1.shape[0]
        Lout = (L - 1) // 2 + 1
        xp = planes_of(x)
        c1 = torch.empty((B, Lout, Cout), dtype=_f32, device=x.device)
        cr = torch.empty((B, Lout, Cout), dtype=_f32, device=x.device)
        gemm_tc_kmajor(tc_operand_conv(xp, B, L, Cin, Lout, 2, 1, -1), wf1, Cout, 3 * Cin,
                       _epi(Scatter(c1.data_ptr(), Lout * Cout, Lout, Cout, 1, 0), bias=b1))
        gemm_tc_kmajor(tc_operand_conv(xp, B, L, Cin, Lout, 2, 1, 0), wfr, Cout, Cin,
                       _epi(Scatter(cr.data_ptr(), Lout * Cout, Lout, Cout, 1, 0), bias=br))
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[3]
        ctx.save_for_backward(x, xp if need_w else None, wd1a, wd1b, wdr)
        ctx.cfg = (tuple(w1.shape), bool(zero_bias_grad))
        ctx.sinks = (_sink(b1), _sink(br))
        return c1, cr

    @staticmethod
    def backward(ctx, d1, dr):
        x, xp, wd1a, wd1b, wdr = ctx.saved_tensors
        (Cout, Cin, _), zero_b = ctx.cfg
        d1, dr = d1.contiguous(), dr.contiguous()
        B, L, _ = x.shape
        Lout = d1.shape[1]
        M = B * Lout
        dev = x.device
        d1p, drp = planes_of(d1), planes_of(dr)
        dx = dW1 = dWr = db1 = dbr = None
        xpl = xp if xp is not None else planes_of(x)
        if ctx.needs_input_grad[1]:
            g = torch.empty((3 * Cin, Cout), dtype=_f32, device=dev)
            gemm_tc_wgrad(tc_operand_conv(xpl, B, L, Cin, Lout, 2, 1, -1), d1p, Cout, 3 * Cin, g)
            dW1 = g.view(3, Cin, Cout).permute(2, 1, 0)
        if ctx.needs_input_grad[3]:
            g = torch.empty((Cin, Cout), dtype=_f32, device=dev)
            gemm_tc_wgrad(tc_operand_conv(xpl, B, L, Cin, Lout, 2, 1, 0), drp, Cout, Cin, g)
            dWr = g.view(1, Cin, Cout).permute(2, 1, 0)
        for need, d, sink, which in ((ctx.needs_input_grad[2], d1, ctx.sinks[0], 0),
                                     (ctx.needs_input_grad[4], dr, ctx.sinks[1], 1)):
            if not need:
                continue
            if zero_b:       # bias in front of a training-mode BatchNorm: exact zero gradient
                val = None if sink is not None else torch.zeros(Cout, dtype=_f32, device=dev)
            elif sink is not None:
                colsum(d.view(M, Cout), out=sink, accumulate=True)
                val = None
            else:
                val = colsum(d.view(M, Cout))
            if which == 0:
                db1 = val
            else:
                dbr = val
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)

            def run(dyp, rows, taps, planes, d_off, acc):
                gemm_tc_kmajor(tc_operand_conv(dyp, B, Lout, Cout, rows, 1, 1, 0), planes, Cin,
                               taps * Cout, _epi(Scatter(dx.data_ptr(), L * Cin, rows, Cin, 2, d_off),
                                                 accumulate=acc))
            ne, no = (L + 1) // 2, L // 2
            run(d1p, ne, 1, wd1a, 0, 0)          # even rows: d1[t] . W1_1^T
            if no > 0:
                run(d1p, no, 2, wd1b, 1, 0)      # odd rows: d1[t] . W1_2^T + d1[t+1] . W1_0^T
            run(drp, ne, 1, wdr, 0, 1)           # even rows += dr[t] . Wr^T
        return dx, dW1, db1, dWr, dbr, None, None, None, None, None, None


_CONV_TAPMAPS = {(3, 1): [(2, 1, 0)], (3, 2): [(1,), (2, 0)], (1, 2): [(0,)]}


def conv_pair_w(x, conv1, conv_r, wp, zero_bias_grad=False):
    """(conv1(x), residual_path(x)) for a stride-2 ResBlock through one autograd node, or None when
    the pair is not eligible (the caller then uses conv1d_w twice)."""
    if wp is None or not x.requires_grad or os.environ.get("SSB_CONVPAIR", "1") == "0":
        return None
    w1, wr = conv1.weight, conv_r.weight
    if (tuple(w1.shape[2:]) != (3,) or tuple(wr.shape[2:]) != (1,) or conv1.stride != (2,)
            or conv_r.stride != (2,) or conv1.bias is None or conv_r.bias is None):
        return None
    Cout, Cin, _ = w1.shape
    B, L, _ = x.shape
    Lout = (L - 1) // 2 + 1
    wf1, wfr = wp.get(w1, "conv_f"), wp.get(wr, "conv_f")
    wd1 = [wp.get(w1, ("conv_d", tm)) for tm in _CONV_TAPMAPS[(3, 2)]]
    wdr = wp.get(wr, ("conv_d", _CONV_TAPMAPS[(1, 2)][0]))
    ok = (all(t is not None for t in (wf1, wfr, wdr, *wd1))
          and _tc_fwd_ok(B * Lout, Cout, 3 * Cin, Cin) and _tc_wgrad_ok(B * Lout, Cout, 3 * Cin, Cin)
          and _tc_fwd_ok(B * Lout, Cout, Cin, Cin) and _tc_wgrad_ok(B * Lout, Cout, Cin, Cin)
          and _tc_fwd_ok(B * L // 2, Cin, Cout, Cout))
    if not ok:
        return None
    return _ConvPairWFn.apply(x, w1, conv1.bias, wr, conv_r.bias, wf1, wd1[0], wd1[1], wfr, wdr,
                              bool(zero_bias_grad))


def conv1d_w(x, conv, wp, ksize, stride, gemm_weight, zero_bias_grad=False):
    """nn.Conv1d `conv` on channels-last x through the WeightPlanes arena `wp` when eligible;
    `gemm_weight()` lazily builds the (k*Cin, Cout) matrix of the round-1 path otherwise.
    zero_bias_grad: the output feeds a training-mode BatchNorm and nothing else."""
    w = conv.weight
    Cout, Cin, _ = w.shape
    B, L, _ = x.shape
    Lout = (L - 1) // stride + 1
    wf = wp.get(w, "conv_f") if wp is not None else None
    tms = _CONV_TAPMAPS.get((ksize, stride), [])
    wd = [wp.get(w, ("conv_d", tm)) for tm in tms] if wp is not None else []
    ok = (wf is not None and wd and all(d is not None for d in wd)
          and _tc_fwd_ok(B * Lout, Cout, ksize * Cin, Cin) and _tc_wgrad_ok(B * Lout, Cout, ksize * Cin, Cin)
          and _tc_fwd_ok(B * L // max(stride, 1), Cin, Cout, Cout))
    if ok:
        return _ConvWFn.apply(x, w, conv.bias, ksize, stride, wf, wd[0], wd[1] if len(wd) > 1 else None,
                              bool(zero_bias_grad))
    return conv1d_cl(x, gemm_weight(), conv.bias, ksize, stride)


# ------------------------------------------------------------------------------------------
# BatchNorm1d (+ optional second normalised branch) (+ ReLU), channels-last (rows, C)
# ------------------------------------------------------------------------------------------
def _bn_stats(x2, gamma, beta, rm, rv, training, momentum, eps):
    lib = _lib.load()
    rows, C = x2.shape
    dev = x2.device
    buf = torch.empty((4, C), dtype=_f32, device=dev)  # mean, rstd, scale, shift
    ws = _ws(lib.ssb_col_partials_bytes(rows, C), dev)
    _lib.check(lib.ssb_bn_stats(x2.data_ptr(), rows, C, gamma.data_ptr(), beta.data_ptr(),
                                rm.data_ptr() if rm is not None else None,
                                rv.data_ptr() if rv is not None else None, momentum, eps,
                                int(training), buf[0].data_ptr(), buf[1].data_ptr(),
                                buf[2].data_ptr(), buf[3].data_ptr(), ws.data_ptr(), ws.numel(),
                                _stream()))
    return buf


def _bn_bwd(dy2, mask_src, x2, stats, gamma, training, sinks=(None, None)):
    """sinks: gradient-bucket views of (gamma, beta): the finalize kernel adds into them and None
    is returned in their place."""
    lib = _lib.load()
    rows, C = x2.shape
    dev = x2.device
    dx = torch.empty_like(x2)
    dxp = (torch.empty((2, rows, C), dtype=torch.bfloat16, device=dev)
           if _want_planes(rows, C) else None)
    sunk = sinks[0] is not None and sinks[1] is not None
    dg = sinks[0] if sunk else torch.empty(C, dtype=_f32, device=dev)
    db = sinks[1] if sunk else torch.empty(C, dtype=_f32, device=dev)
    ws = _ws(lib.ssb_col_partials_bytes(rows, C) + 12 * C, dev)
    _lib.check(lib.ssb_bn_bwd(dy2.data_ptr(), mask_src.data_ptr() if mask_src is not None else None,
                              x2.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                              gamma.data_ptr(), int(training), rows, C, dx.data_ptr(),
                              dxp.data_ptr() if dxp is not None else None,
                              dg.data_ptr(), db.data_ptr(), int(sunk), ws.data_ptr(), ws.numel(),
                              _stream()))
    return dx, (None if sunk else dg), (None if sunk else db), dxp


def _bn_bwd2(dy2, mask_src, xa2, sa, ga, xb2, sb, gb, training, sinks):
    """Both normalised branches of a ResBlock output in one pair of passes (ssb_bn_bwd2).
    sinks: bucket views of (gamma_a, beta_a, gamma_b, beta_b) or Nones."""
    lib = _lib.load()
    rows, C = xa2.shape
    dev = xa2.device
    dxa, dxb = torch.empty_like(xa2), torch.empty_like(xb2)
    pl = _want_planes(rows, C)
    dxap = torch.empty((2, rows, C), dtype=torch.bfloat16, device=dev) if pl else None
    dxbp = torch.empty((2, rows, C), dtype=torch.bfloat16, device=dev) if pl else None
    sunk = all(t is not None for t in sinks)
    pg = list(sinks) if sunk else [torch.empty(C, dtype=_f32, device=dev) for _ in range(4)]
    ws = _ws(lib.ssb_col_partials_bytes(rows, C) + 12 * C, dev)
    _lib.check(lib.ssb_bn_bwd2(
        dy2.data_ptr(), mask_src.data_ptr() if mask_src is not None else None,
        xa2.data_ptr(), sa[0].data_ptr(), sa[1].data_ptr(), ga.data_ptr(),
        xb2.data_ptr(), sb[0].data_ptr(), sb[1].data_ptr(), gb.data_ptr(), int(training), rows, C,
        dxa.data_ptr(), dxap.data_ptr() if pl else None, dxb.data_ptr(),
        dxbp.data_ptr() if pl else None, pg[0].data_ptr(), pg[1].data_ptr(), pg[2].data_ptr(),
        pg[3].data_ptr(), int(sunk), ws.data_ptr(), ws.numel(), _stream()))
    if sunk:
        pg = [None] * 4
    return dxa, dxap, dxb, dxbp, pg


class _BNActFn(torch.autograd.Function):
    """y = [relu]( BN_a(xa) [+ BN_b(xb)] );  running stats updated in place when training."""

    @staticmethod
    def forward(ctx, xa, ga, ba, rma, rva, xb, gb, bb, rmb, rvb, training, relu, momentum, eps):
        lib = _lib.load()
        _chk(xa, "xa")
        shape = xa.shape
        C = shape[-1]
        xa2 = xa.view(-1, C)
        rows = xa2.shape[0]
        sa = _bn_stats(xa2, ga, ba, rma, rva, training, momentum, eps)
        sb = None
        xb2 = None
        if xb is not None:
            _chk(xb, "xb")
            xb2 = xb.view(-1, C)
            sb = _bn_stats(xb2, gb, bb, rmb, rvb, training, momentum, eps)
        y = torch.empty_like(xa)
        yp = (torch.empty((2,) + tuple(shape), dtype=torch.bfloat16, device=xa.device)
              if _want_planes(rows, C) else None)
        _lib.check(lib.ssb_bn_apply(xa2.data_ptr(), sa[0].data_ptr(), sa[2].data_ptr(), ba.data_ptr(),
                                    xb2.data_ptr() if xb2 is not None else None,
                                    sb[0].data_ptr() if sb is not None else None,
                                    sb[2].data_ptr() if sb is not None else None,
                                    bb.data_ptr() if sb is not None else None, int(relu), rows,
                                    C, y.data_ptr(), yp.data_ptr() if yp is not None else None,
                                    _stream()))
        if yp is not None:
            attach_planes(y, yp)     # the next convolution's operand, written by this kernel
        ctx.training, ctx.relu, ctx.two = training, relu, xb is not None
        ctx.sinks = (_sink(ga), _sink(ba), _sink(gb) if xb is not None else None,
                     _sink(bb) if xb is not None else None)
        if ctx.two:
            ctx.save_for_backward(xa, ga, sa, y, xb, gb, sb)
        else:
            ctx.save_for_backward(xa, ga, sa, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        if ctx.two:
            xa, ga, sa, y, xb, gb, sb = ctx.saved_tensors
        else:
            xa, ga, sa, y = ctx.saved_tensors
        C = xa.shape[-1]
        dy2 = dy.view(-1, C)
        mask = y.view(-1, C) if ctx.relu else None
        def shaped(d, dp, like):     # gradient of a convolution output + its operand planes
            d = d.view_as(like)
            return attach_planes(d, dp.view((2,) + tuple(like.shape))) if dp is not None else d
        if ctx.two:
            dxa, dxap, dxb, dxbp, (dga, dba, dgb, dbb) = _bn_bwd2(
                dy2, mask, xa.view(-1, C), sa, ga, xb.view(-1, C), sb, gb, ctx.training, ctx.sinks)
            return (shaped(dxa, dxap, xa), dga, dba, None, None, shaped(dxb, dxbp, xb), dgb, dbb,
                    None, None, None, None, None, None)
        dxa, dga, dba, dxap = _bn_bwd(dy2, mask, xa.view(-1, C), sa, ga, ctx.training, ctx.sinks[:2])
        return (shaped(dxa, dxap, xa), dga, dba, None, None, None, None, None, None, None, None,
                None, None, None)


def bn_act(xa, ga, ba, rma, rva, training, relu, xb=None, gb=None, bb=None, rmb=None, rvb=None,
           momentum=0.1, eps=1e-5):
    return _BNActFn.apply(xa, ga, ba, rma, rva, xb, gb, bb, rmb, rvb, bool(training), bool(relu),
                          float(momentum), float(eps))


# ------------------------------------------------------------------------------------------
# z = res + dropout(branch); y = LayerNorm(z)          transformer.py:55-56, 58-59
# ------------------------------------------------------------------------------------------
class _AddDropLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, res, branch, gamma, beta, p, seed, site, eps):
        lib = _lib.load()
        _chk(res, "res"), _chk(branch, "branch")
        rows, D = res.shape
        dev = res.device
        need_bwd = any(ctx.needs_input_grad[:4])
        y = torch.empty_like(res)
        yp = (torch.empty((2, rows, D), dtype=torch.bfloat16, device=dev)
              if _want_planes(rows, D) else None)
        z = torch.empty_like(res) if need_bwd else None
        stat = torch.empty((2, rows), dtype=_f32, device=dev) if need_bwd else None
        _lib.check(lib.ssb_add_dropout_ln_fwd(
            res.data_ptr(), branch.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, D, eps, p,
            seed & 0xFFFFFFFFFFFFFFFF, site, z.data_ptr() if need_bwd else None, y.data_ptr(),
            stat[0].data_ptr() if need_bwd else None, stat[1].data_ptr() if need_bwd else None,
            yp.data_ptr() if yp is not None else None, _stream()))
        if yp is not None:
            attach_planes(y, yp)     # operand of the next QKV / FFN GEMM
        if need_bwd:
            ctx.save_for_backward(z, stat, gamma)
        ctx.cfg = (p, seed, site)
        ctx.sinks = (_sink(gamma), _sink(beta))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        z, stat, gamma = ctx.saved_tensors
        p, seed, site = ctx.cfg
        dy = dy.contiguous()
        rows, D = z.shape
        dev = z.device
        d_res = torch.empty_like(z)
        d_branch = torch.empty_like(z)
        dbp = (torch.empty((2, rows, D), dtype=torch.bfloat16, device=dev)
               if _want_planes(rows, D) else None)
        sunk = ctx.sinks[0] is not None and ctx.sinks[1] is not None
        dg = ctx.sinks[0] if sunk else torch.empty(D, dtype=_f32, device=dev)
        db = ctx.sinks[1] if sunk else torch.empty(D, dtype=_f32, device=dev)
        ws = _ws(lib.ssb_add_dropout_ln_bwd_workspace_bytes(rows, D), dev)
        _lib.check(lib.ssb_add_dropout_ln_bwd(
            dy.data_ptr(), z.data_ptr(), stat[0].data_ptr(), stat[1].data_ptr(), gamma.data_ptr(),
            rows, D, p, seed & 0xFFFFFFFFFFFFFFFF, site, d_res.data_ptr(), d_branch.data_ptr(),
            dbp.data_ptr() if dbp is not None else None,
            dg.data_ptr(), db.data_ptr(), int(sunk), ws.data_ptr(), ws.numel(), _stream()))
        if dbp is not None:
            attach_planes(d_branch, dbp)   # dy operand of the branch's backward GEMMs
        return (d_res, d_branch, None if sunk else dg, None if sunk else db, None, None, None, None)


def add_dropout_layernorm(res, branch, gamma, beta, p=0.0, seed=0, site=0, eps=1e-5):
    return _AddDropLNFn.apply(res, branch, gamma, beta, float(p), int(seed), int(site), float(eps))


# ------------------------------------------------------------------------------------------
# banded relative-position attention            transformer.py:99-110, 162-297
# qkv: (B*T, 3D) [q|k|v];  E: (H, RW, dh) relative-position table zero-padded to RW rows,
# never differentiated.  Returns O: (B*T, D).
# ------------------------------------------------------------------------------------------
class _BandAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, E, B, T, H, dh, W, p, seed, site):
        lib = _lib.load()
        _chk(qkv, "qkv"), _chk(E, "E")
        M, D3 = qkv.shape
        D = H * dh
        RW = E.shape[1]
        assert M == B * T and D3 == 3 * D and E.shape == (H, RW, dh)
        dev = qkv.device
        need_bwd = ctx.needs_input_grad[0]
        R = torch.empty((M, H, RW), dtype=_f32, device=dev)
        for h in range(H):  # R[:, h, :] = q_h @ E_h^T
            ga = Gather(qkv.data_ptr() + 4 * h * dh, 0, M, dh, M, D3, 1, 0, 0)
            out = Scatter(R.data_ptr() + 4 * h * RW, 0, M, H * RW, 1, 0)
            gemm_nt(ga, E[h], dh, (0,), _epi(out), M, RW, dh)
        P = torch.empty((M, H, RW), dtype=_f32, device=dev) if need_bwd else None
        O = torch.empty((M, D), dtype=_f32, device=dev)
        _lib.check(lib.ssb_band_attn_fwd(qkv.data_ptr(), R.data_ptr(), B, T, H, dh, W, RW, p,
                                         seed & 0xFFFFFFFFFFFFFFFF, site,
                                         P.data_ptr() if need_bwd else None, O.data_ptr(),
                                         _stream()))
        if need_bwd:
            ctx.save_for_backward(qkv, E, P)
        ctx.cfg = (B, T, H, dh, W, p, seed, site)
        return O

    @staticmethod
    def backward(ctx, dO):
        lib = _lib.load()
        qkv, E, P = ctx.saved_tensors
        B, T, H, dh, W, p, seed, site = ctx.cfg
        dO = dO.contiguous()
        M, D3 = qkv.shape
        RW = E.shape[1]
        dS = torch.empty_like(P)
        dqkv = torch.empty_like(qkv)
        _lib.check(lib.ssb_band_attn_bwd(qkv.data_ptr(), P.data_ptr(), dO.data_ptr(), B, T, H, dh,
                                         W, RW, p, seed & 0xFFFFFFFFFFFFFFFF, site, dS.data_ptr(),
                                         dqkv.data_ptr(), _stream()))
        for h in range(H):  # dq_h += dS_h @ E_h   (positional part; no gradient for E: F3)
            ga = Gather(dS.data_ptr() + 4 * h * RW, 0, M, RW, M, H * RW, 1, 0, 0)
            out = Scatter(dqkv.data_ptr() + 4 * h * dh, 0, M, D3, 1, 0)
            gemm_nn(ga, E[h], _epi(out, accumulate=1), M, dh, RW)
        return dqkv, None, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------
# Tensor-core attention: the same banded relative-position attention, scheduled as batched
# tcgen05 GEMMs (one batch item per (b, h)) around two fused element-wise kernels
# (csrc/attn_tc.cu).  Per (b, h) the logits are a dense (T x T) tile masked to the exact band;
# at T = 500 that is 2.5x the band's flops, on a pipe that is >20x faster than CUDA cores.
# ------------------------------------------------------------------------------------------
_HP = 128     # padded head dim of every plane tensor
_RWP = 256    # padded band width (K of the positional dQ GEMM)


def _op(planes, offset_elems, plane_stride, batches, rows, C, ld, s_lo, div=0, s_hi=0):
    return TcOperand(planes.data_ptr() + 2 * offset_elems, plane_stride, s_lo, batches, rows, rows,
                     C, ld, 1, 0, 0, div, 0, s_hi)


def _bscatter(t, offset_elems, rows, ld, s_lo, s_hi):
    return Scatter(t.data_ptr() + 4 * offset_elems, s_lo, rows, ld, 1, 0, s_hi)


def _tc_batched(A, Bo, b_mode, N, K, epi):
    lib = _lib.load()
    _lib.check(lib.ssb_gemm_tc_batched(ctypes.byref(A), ctypes.byref(Bo), b_mode, N, K,
                                       ctypes.byref(epi), _stream()))


def _tc_batched_tn(X, G, N, K, epi):
    lib = _lib.load()
    _lib.check(lib.ssb_gemm_tc_batched_tn(ctypes.byref(X), ctypes.byref(G), N, K, ctypes.byref(epi),
                                          _stream()))


class _DenseTCAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, E, B, T, H, dh, W, p, seed, site):
        lib = _lib.load()
        _chk(qkv, "qkv")
        M, D3 = qkv.shape
        D = H * dh
        BH = B * H
        Tp = (T + 63) // 64 * 64
        RW = (2 * W + 1 + 3) // 4 * 4
        dev = qkv.device
        bf = torch.bfloat16
        st = _stream()
        # split planes with heads zero-padded to 128: (2, B*T, 3H, 128) = [q heads | k heads | v heads]
        qkvp = torch.empty((2, M, 3 * H, _HP), dtype=bf, device=dev)
        _lib.check(lib.ssb_pad_split_heads(qkv.data_ptr(), M, D3, 0, 3 * H, dh, qkvp.data_ptr(), st))
        vt = torch.empty((2, BH, _HP, Tp), dtype=bf, device=dev)
        _lib.check(lib.ssb_transpose_split_heads(qkv.data_ptr(), D3, 2 * D, B, T, H, dh, Tp,
                                                 vt.data_ptr(), st))
        ep = _const_planes(E, ("fwd", W, dh, RW), lambda: torch.nn.functional.pad(
            E[:, :2 * W + 1, :dh], (0, _HP - dh, 0, RW - (2 * W + 1))).contiguous())
        ld_qkv = 3 * H * _HP
        pq = M * ld_qkv
        q_op = _op(qkvp, 0, pq, BH, T, _HP, ld_qkv, _HP, H, T * ld_qkv)
        k_op = _op(qkvp, H * _HP, pq, BH, T, _HP, ld_qkv, _HP, H, T * ld_qkv)
        # S = Q K^T (raw, unscaled)
        S = torch.empty((BH, T, Tp), dtype=_f32, device=dev)
        _tc_batched(q_op, k_op, 1, Tp, _HP, _epi(_bscatter(S, 0, T, Tp, T * Tp, H * T * Tp)))
        # R = Q E^T  (positional logits, band layout)
        R = torch.empty((BH, T, RW), dtype=_f32, device=dev)
        e_op = _op(ep, 0, H * RW * _HP, H, RW, _HP, _HP, RW * _HP)
        _tc_batched(q_op, e_op, 2, RW, _HP, _epi(_bscatter(R, 0, T, RW, T * RW, H * T * RW)))
        # P = softmax(scale*S + R) inside the band (in place), dropout(P) as planes
        pd = torch.empty((2, BH, T, Tp), dtype=bf, device=dev)
        _lib.check(lib.ssb_attn_softmax_fwd(S.data_ptr(), R.data_ptr(), B, H, T, Tp, W, RW, dh, p,
                                            seed & 0xFFFFFFFFFFFFFFFF, site, pd.data_ptr(), st))
        # O = dropout(P) V
        O = torch.empty((M, D), dtype=_f32, device=dev)
        pd_op = _op(pd, 0, BH * T * Tp, BH, T, Tp, Tp, T * Tp, H, H * T * Tp)
        vt_op = _op(vt, 0, BH * _HP * Tp, BH, _HP, Tp, Tp, _HP * Tp, H, H * _HP * Tp)
        _tc_batched(pd_op, vt_op, 1, dh, Tp, _epi(_bscatter(O, 0, T, D, dh, T * D)))
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(qkv, qkvp, S, pd, E)
        ctx.cfg = (B, T, H, dh, W, p, seed, site, Tp)
        return O

    @staticmethod
    def backward(ctx, dO):
        lib = _lib.load()
        qkv, qkvp, P, pd, E = ctx.saved_tensors
        B, T, H, dh, W, p, seed, site, Tp = ctx.cfg
        dO = dO.contiguous()
        M, D3 = qkv.shape
        D = H * dh
        BH = B * H
        dev = qkv.device
        bf = torch.bfloat16
        st = _stream()
        ld_qkv = 3 * H * _HP
        pq = M * ld_qkv
        q_op = _op(qkvp, 0, pq, BH, T, _HP, ld_qkv, _HP, H, T * ld_qkv)
        v_op = _op(qkvp, 2 * H * _HP, pq, BH, T, _HP, ld_qkv, _HP, H, T * ld_qkv)
        dop = torch.empty((2, M, H, _HP), dtype=bf, device=dev)
        _lib.check(lib.ssb_pad_split_heads(dO.data_ptr(), M, D, 0, H, dh, dop.data_ptr(), st))
        do_op = _op(dop, 0, M * H * _HP, BH, T, _HP, H * _HP, _HP, H, T * H * _HP)
        # dP = dO V^T
        dP = torch.empty((BH, T, Tp), dtype=_f32, device=dev)
        _tc_batched(do_op, v_op, 1, Tp, _HP, _epi(_bscatter(dP, 0, T, Tp, T * Tp, H * T * Tp)))
        # dS (scaled planes for dQ / dK, band planes for the positional dQ)
        dsp = torch.empty((2, BH, T, Tp), dtype=bf, device=dev)
        dsb = torch.empty((2, M, H, _RWP), dtype=bf, device=dev)
        _lib.check(lib.ssb_attn_ds_bwd(P.data_ptr(), dP.data_ptr(), B, H, T, Tp, W, _RWP, dh, p,
                                       seed & 0xFFFFFFFFFFFFFFFF, site, dsp.data_ptr(),
                                       dsb.data_ptr(), st))
        dqkv = torch.empty_like(qkv)
        ds_op = _op(dsp, 0, BH * T * Tp, BH, T, Tp, Tp, T * Tp, H, H * T * Tp)
        pd_op = _op(pd, 0, BH * T * Tp, BH, T, Tp, Tp, T * Tp, H, H * T * Tp)
        # dQ = scale * dS K   (+ positional part dS_band E)
        kt = torch.empty((2, BH, _HP, Tp), dtype=bf, device=dev)
        _lib.check(lib.ssb_transpose_split_heads(qkv.data_ptr(), D3, D, B, T, H, dh, Tp,
                                                 kt.data_ptr(), st))
        kt_op = _op(kt, 0, BH * _HP * Tp, BH, _HP, Tp, Tp, _HP * Tp, H, H * _HP * Tp)
        _tc_batched(ds_op, kt_op, 1, dh, Tp, _epi(_bscatter(dqkv, 0, T, D3, dh, T * D3)))
        etp = _const_planes(E, ("bwd", W, dh), lambda: torch.nn.functional.pad(    # (2, H, 128, RWP)
            E[:, :2 * W + 1, :dh], (0, _HP - dh, 0, _RWP - (2 * W + 1))).transpose(1, 2).contiguous())
        dsb_op = _op(dsb, 0, M * H * _RWP, BH, T, _RWP, H * _RWP, _RWP, H, T * H * _RWP)
        et_op = _op(etp, 0, H * _HP * _RWP, H, _HP, _RWP, _RWP, _HP * _RWP)
        _tc_batched(dsb_op, et_op, 2, dh, _RWP,
                    _epi(_bscatter(dqkv, 0, T, D3, dh, T * D3), accumulate=1))
        # dK = scale * dS^T Q ;  dV = dropout(P)^T dO     (reduction over queries: MN-major)
        _tc_batched_tn(ds_op, q_op, dh, T, _epi(_bscatter(dqkv, D, T, D3, dh, T * D3)))
        _tc_batched_tn(pd_op, do_op, dh, T, _epi(_bscatter(dqkv, 2 * D, T, D3, dh, T * D3)))
        return dqkv, None, None, None, None, None, None, None, None, None


_band_bufs = {}            # LRU: (B, T, H, W, rwp, dev, stream) -> zero-initialised buffer
_BAND_LRU = 3
_capture_keepalive = []    # buffers whose addresses were baked into the graph being captured
_exec_stream = None        # while capturing: the stream the graph will be REPLAYED on


def set_capture_execution_stream(stream_handle):
    """A CUDA graph is captured on torch's private capture stream but executes on the stream that
    replays it.  Per-stream scratch (below) must be keyed by the latter: the capturer passes the
    handle of its replay stream before capture and None afterwards."""
    global _exec_stream
    _exec_stream = stream_handle



def take_capture_keepalive():
    """Buffers from this module's caches that the CUDA graph captured since the last call holds by
    ADDRESS.  The capturer (training.GraphedTrainStep) stores the list next to the graph, so the
    caches below may evict freely without ever freeing memory a graph still replays into."""
    global _capture_keepalive
    out, _capture_keepalive = _capture_keepalive, []
    return out


def _band_scratch(key):
    """Zero-initialised (2, B*T, H, RWp) bf16 buffer for one attention geometry ON ONE STREAM.
    The fused backward overwrites exactly the in-band, in-sequence entries - the same set every
    call for a given geometry - and everything else must read 0, so the buffer is zeroed once and
    reused by every layer (their backward passes are ordered on the stream the key names; another
    stream gets its own buffer).  Bounded: real SizeAwareSampler batches change B almost every
    step, so only the _BAND_LRU most recent geometries are kept (131 MB each at cfg-1); a buffer
    captured into a CUDA graph is additionally kept alive by that graph's owner."""
    key = key + (_exec_stream if _exec_stream is not None
                 else torch.cuda.current_stream().cuda_stream,)
    buf = _band_bufs.pop(key, None)
    if buf is None:
        B, T, H, W, rwp, dev, _ = key
        if torch.cuda.is_current_stream_capturing():
            # first sight of this geometry under capture (GraphedTrainStep runs one eager step
            # first, so this is rare): graph-pool memory, zeroed by a captured memset per replay
            return torch.zeros((2, B * T, H, rwp), dtype=torch.bfloat16, device=dev)
        buf = torch.zeros((2, B * T, H, rwp), dtype=torch.bfloat16, device=dev)
    _band_bufs[key] = buf                      # most recently used last
    while len(_band_bufs) > _BAND_LRU:
        _band_bufs.pop(next(iter(_band_bufs)))
    if torch.cuda.is_current_stream_capturing():
        _capture_keepalive.append(buf)
    return buf


def _const_planes(E, key, make):
    """Split planes derived from the constant positional table E, cached on the tensor object
    (transformer.LearnedRelativePositionalEmbedding.padded_table keeps E alive while valid)."""
    cache = getattr(E, "_ssb_planes", None)
    if cache is not None and cache[0] == E._version and key in cache[1]:
        return cache[1][key]
    planes = split_planes(make())
    if torch.cuda.is_current_stream_capturing():
        return planes                   # graph-pool memory must not outlive the capture
    if cache is None or cache[0] != E._version:
        cache = (E._version, {})
        E._ssb_planes = cache
    cache[1][key] = planes
    return planes


def _fused_attn_fwd(qkv, E, B, T, H, dh, W, p, seed, site, need_bwd=True, packed=None, o_planes=None):
    """Fused band attention forward.  qkv: fp32 (M, 3D), re-laid here into head-padded planes; or
    packed = (2, M, 3D) split planes straight from the QKV GEMM's epilogue (head stride dh): the
    kernels' tensor maps only ever address the first dh columns of a head, and the positional
    GEMM's 128-wide reduction reads the next head's first columns against zero rows of the padded
    table.  Returns (O, saved-for-backward)."""
    lib = _lib.load()
    D = H * dh
    BH = B * H
    RW = (2 * W + 1 + 3) // 4 * 4
    bf = torch.bfloat16
    st = _stream()
    if packed is None:
        M, D3 = qkv.shape
        dev = qkv.device
        hs = _HP
        qkvp = torch.empty((2, M, 3 * H, _HP), dtype=bf, device=dev)
        _lib.check(lib.ssb_pad_split_heads(qkv.data_ptr(), M, D3, 0, 3 * H, dh, qkvp.data_ptr(), st))
    else:
        qkvp, hs = packed, dh
        M, dev = packed.shape[1], packed.device
    ep = _const_planes(E, ("fwd", W, dh, RW), lambda: torch.nn.functional.pad(
        E[:, :2 * W + 1, :dh], (0, _HP - dh, 0, RW - (2 * W + 1))).contiguous())
    ld_qkv = 3 * H * hs
    q_op = _op(qkvp, 0, M * ld_qkv, BH, T, _HP, ld_qkv, hs, H, T * ld_qkv)
    R = torch.empty((BH, T, RW), dtype=_f32, device=dev)
    e_op = _op(ep, 0, H * RW * _HP, H, RW, _HP, _HP, RW * _HP)
    _tc_batched(q_op, e_op, 2, RW, _HP, _epi(_bscatter(R, 0, T, RW, T * RW, H * T * RW)))
    O = torch.empty((M, D), dtype=_f32, device=dev)
    stats = torch.empty((2, BH, T), dtype=_f32, device=dev)
    _lib.check(lib.ssb_attn_fused_fwd(qkvp.data_ptr(), R.data_ptr(), B, T, H, dh, W, RW, p,
                                      seed & 0xFFFFFFFFFFFFFFFF, site, O.data_ptr(),
                                      o_planes.data_ptr() if o_planes is not None else None,
                                      stats[0].data_ptr(), stats[1].data_ptr(), hs, st))
    return O, ((qkvp, R, stats, O, E) if need_bwd else None)


def _fused_attn_bwd(saved, cfg, dO, dO_packed=None, planes_out=False):
    """-> dqkv (M, 3D) fp32 (content + positional parts; no gradient for E: SURVEY.md F3), or with
    planes_out its (2, M, 3D) bf16 split planes (dK / dV thirds written by the kernel's epilogue,
    the dQ third split from a (M, D) fp32 accumulator after the positional GEMM).
    dO_packed: (2, M, D) split planes of dO when its producer already wrote them."""
    lib = _lib.load()
    qkvp, R, stats, O, E = saved
    B, T, H, dh, W, p, seed, site, RW = cfg
    dO = dO.contiguous()
    M, D = O.shape
    D3 = 3 * D
    BH = B * H
    dev = O.device
    bf = torch.bfloat16
    st = _stream()
    hs = dh if qkvp.dim() == 3 else _HP          # (2, M, 3D) packed vs (2, M, 3H, 128) padded
    if dO_packed is not None:
        dop, dhs = dO_packed, dh
    else:
        dhs = _HP
        dop = torch.empty((2, M, H, _HP), dtype=bf, device=dev)
        _lib.check(lib.ssb_pad_split_heads(dO.data_ptr(), M, D, 0, H, dh, dop.data_ptr(), st))
    delta = torch.empty((BH, T), dtype=_f32, device=dev)
    ldq = D if planes_out else D3
    dqkv = torch.empty((M, ldq), dtype=_f32, device=dev)
    dqp = torch.empty((2, M, D3), dtype=bf, device=dev) if planes_out else None
    # content dQ arrives by red.global.add: its columns are cleared by the delta pass
    _lib.check(lib.ssb_attn_delta(O.data_ptr(), dO.data_ptr(), B, T, H, dh, delta.data_ptr(),
                                  dqkv.data_ptr(), ldq, st))
    # band-layout dS: the kernel overwrites exactly the in-band, in-sequence entries - the same
    # set every call for a given geometry - and everything else must read 0.  One persistent
    # buffer per geometry, zeroed once, replaces a 131 MB fill per layer and step (backward
    # passes of successive layers are stream-ordered, so they can share it).
    dsb = _band_scratch((B, T, H, W, _RWP, dev))
    _lib.check(lib.ssb_attn_fused_bwd(qkvp.data_ptr(), dop.data_ptr(), R.data_ptr(),
                                      stats[0].data_ptr(), stats[1].data_ptr(),
                                      delta.data_ptr(), B, T, H, dh, W, RW, p,
                                      seed & 0xFFFFFFFFFFFFFFFF, site, dqkv.data_ptr(), ldq,
                                      dqp.data_ptr() if planes_out else None,
                                      dsb.data_ptr(), _RWP, hs, dhs, st))
    # positional part: dQ += dS_band E
    etp = _const_planes(E, ("bwd", W, dh), lambda: torch.nn.functional.pad(    # (2, H, 128, RWP)
        E[:, :2 * W + 1, :dh], (0, _HP - dh, 0, _RWP - (2 * W + 1))).transpose(1, 2).contiguous())
    dsb_op = _op(dsb, 0, M * H * _RWP, BH, T, _RWP, H * _RWP, _RWP, H, T * H * _RWP)
    et_op = _op(etp, 0, H * _HP * _RWP, H, _HP, _RWP, _RWP, _HP * _RWP)
    _tc_batched(dsb_op, et_op, 2, dh, _RWP,
                _epi(_bscatter(dqkv, 0, T, ldq, dh, T * ldq), accumulate=1))
    if not planes_out:
        return dqkv
    _lib.check(lib.ssb_split_bf16_2d(dqkv.data_ptr(), M, D, D, dqp.data_ptr(), D3, M * D3, st))
    return dqp


class _FusedAttnFn(torch.autograd.Function):
    """Fused band attention (csrc/attn_fused.cu): S, P, dP and dS never leave the SM.  Around the
    two fused kernels only the positional GEMMs with E remain: R = Q E^T before the forward and
    dQ += dS_band E after the backward (no gradient flows to E: SURVEY.md F3)."""

    @staticmethod
    def forward(ctx, qkv, E, B, T, H, dh, W, p, seed, site):
        _chk(qkv, "qkv")
        O, saved = _fused_attn_fwd(qkv, E, B, T, H, dh, W, p, seed, site, ctx.needs_input_grad[0])
        if saved is not None:
            ctx.save_for_backward(*saved)
        ctx.cfg = (B, T, H, dh, W, p, seed, site, (2 * W + 1 + 3) // 4 * 4)
        return O

    @staticmethod
    def backward(ctx, dO):
        dqkv = _fused_attn_bwd(ctx.saved_tensors, ctx.cfg, dO)
        return dqkv, None, None, None, None, None, None, None, None, None


def _ln_fwd(res, branch, gamma, beta, p, seed, site, eps, need_bwd=True):
    """z = res + dropout(branch); y = LayerNorm(z).  -> y (planes attached), z, stat."""
    lib = _lib.load()
    rows, D = res.shape
    dev = res.device
    y = torch.empty_like(res)
    yp = torch.empty((2, rows, D), dtype=torch.bfloat16, device=dev) if _want_planes(rows, D) else None
    z = torch.empty_like(res) if need_bwd else None
    stat = torch.empty((2, rows), dtype=_f32, device=dev) if need_bwd else None
    _lib.check(lib.ssb_add_dropout_ln_fwd(
        res.data_ptr(), branch.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, D, eps, p,
        seed & 0xFFFFFFFFFFFFFFFF, site, z.data_ptr() if need_bwd else None, y.data_ptr(),
        stat[0].data_ptr() if need_bwd else None, stat[1].data_ptr() if need_bwd else None,
        yp.data_ptr() if yp is not None else None, _stream()))
    if yp is not None:
        attach_planes(y, yp)
    return y, z, stat


def _ln_bwd(dy, z, stat, gamma, p, seed, site, sinks, planes_only=False):
    """-> d_res, d_branch, d_branch planes (or None), dgamma, dbeta (None when sunk).
    planes_only: when the planes are produced, skip the fp32 d_branch (returned as None): its only
    readers are tensor-core GEMMs and a column sum that takes planes as well."""
    lib = _lib.load()
    dy = dy.contiguous()
    rows, D = z.shape
    dev = z.device
    d_res = torch.empty_like(z)
    dbp = torch.empty((2, rows, D), dtype=torch.bfloat16, device=dev) if _want_planes(rows, D) else None
    d_branch = None if (planes_only and dbp is not None) else torch.empty_like(z)
    sunk = sinks[0] is not None and sinks[1] is not None
    dg = sinks[0] if sunk else torch.empty(D, dtype=_f32, device=dev)
    db = sinks[1] if sunk else torch.empty(D, dtype=_f32, device=dev)
    ws = _ws(lib.ssb_add_dropout_ln_bwd_workspace_bytes(rows, D), dev)
    _lib.check(lib.ssb_add_dropout_ln_bwd(
        dy.data_ptr(), z.data_ptr(), stat[0].data_ptr(), stat[1].data_ptr(), gamma.data_ptr(),
        rows, D, p, seed & 0xFFFFFFFFFFFFFFFF, site, d_res.data_ptr(),
        d_branch.data_ptr() if d_branch is not None else None,
        dbp.data_ptr() if dbp is not None else None, dg.data_ptr(), db.data_ptr(), int(sunk),
        ws.data_ptr(), ws.numel(), _stream()))
    return d_res, d_branch, dbp, (None if sunk else dg), (None if sunk else db)


class _AttnBlockFn(torch.autograd.Function):
    """x1 = LayerNorm(x + dropout(Attention(x) Wo))   transformer.py:54-56 as ONE autograd node:
    fused QKV GEMM -> fused band attention -> out-projection -> add + dropout + LayerNorm.  The
    backward hands the residual gradient from the LayerNorm kernel to the QKV data-gradient GEMM's
    epilogue (accumulate), so the two gradients of x are never summed by a separate pass, and
    writes dWo and the LayerNorm parameter gradients straight into the gradient bucket."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, w_o, E, gamma, beta, B, T, W, p_attn, p_res, seed, site, eps,
                qf, qb, of, ob):
        _chk(x, "x")
        H, D, dh = wq.shape
        M = x.shape[0]
        need_bwd = any(ctx.needs_input_grad[:8])
        xp = planes_of(x)
        # the projection exists only as packed split planes: the GEMM epilogue writes the operand
        # format of the attention kernels directly (no fp32 qkv, no head re-layout pass)
        qkvp = torch.empty((2, M, 3 * D), dtype=torch.bfloat16, device=x.device)
        gemm_tc_kmajor(tc_operand_plain(xp, M, D), qf, 3 * D, D,
                       _epi(_scatter_plain(None, M, 3 * D), planes_out=qkvp))
        op = torch.empty((2, M, D), dtype=torch.bfloat16, device=x.device)   # written by the kernel's epilogue
        O, att = _fused_attn_fwd(None, E, B, T, H, dh, W, p_attn, seed, site, need_bwd, packed=qkvp,
                                 o_planes=op)
        a = torch.empty((M, D), dtype=_f32, device=x.device)
        gemm_tc_kmajor(tc_operand_plain(op, M, D), of, D, D, _epi(_scatter_plain(a.data_ptr(), M, D)))
        y, z, stat = _ln_fwd(x, a, gamma, beta, p_res, seed, site + 1, eps, need_bwd)
        if need_bwd:
            ctx.save_for_backward(xp, op, z, stat, gamma, qb, ob, w_o, *att)
        ctx.cfg = (B, T, H, dh, W, p_attn, seed, site, (2 * W + 1 + 3) // 4 * 4)
        ctx.p_res = p_res
        ctx.sinks = (_sink(gamma), _sink(beta), _sink(w_o))
        # w_q, w_k, w_v gradients as ONE target when their bucket views are adjacent and in order
        sq, sk, sv = _sink(wq), _sink(wk), _sink(wv)
        n = wq.numel()
        ctx.qkv_sink = (sq if (sq is not None and sk is not None and sv is not None
                               and sq.is_contiguous() and sk.is_contiguous() and sv.is_contiguous()
                               and sk.data_ptr() == sq.data_ptr() + 4 * n
                               and sv.data_ptr() == sk.data_ptr() + 4 * n) else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, op, z, stat, gamma, qb, ob, w_o = ctx.saved_tensors[:8]
        att = ctx.saved_tensors[8:]
        B, T, H, dh, W, p_attn, seed, site, RW = ctx.cfg
        D = H * dh
        M = z.shape[0]
        dev = z.device
        s_g, s_b, s_wo = ctx.sinks
        d_res, d_a, d_ap, dg, db = _ln_bwd(dy, z, stat, gamma, ctx.p_res, seed, site + 1, (s_g, s_b),
                                           planes_only=True)
        if d_ap is None:
            d_ap = split_planes(d_a)
        # out-projection: dO = d_a Wo^T, dWo = O^T d_a (parameter layout = GEMM layout)
        dO = torch.empty((M, D), dtype=_f32, device=dev)           # fp32 for delta = rowsum(O . dO)
        dOp = torch.empty((2, M, D), dtype=torch.bfloat16, device=dev)   # packed operand of the bwd
        gemm_tc_kmajor(tc_operand_plain(d_ap, M, D), ob, D, D,
                       _epi(_scatter_plain(dO.data_ptr(), M, D), planes_out=dOp))
        if s_wo is not None:
            gemm_tc_wgrad(tc_operand_plain(op, M, D), d_ap, D, D, s_wo.view(D, D), accumulate=True)
            dWo = None
        else:
            dWo = torch.empty((D, D), dtype=_f32, device=dev)
            gemm_tc_wgrad(tc_operand_plain(op, M, D), d_ap, D, D, dWo)
            dWo = dWo.view_as(w_o)
        del d_a, d_ap
        dqp = _fused_attn_bwd(att, ctx.cfg, dO, dO_packed=dOp, planes_out=True)
        # x receives d_res (through the LayerNorm) + dqkv Wqkv^T: accumulated by the GEMM epilogue
        gemm_tc_kmajor(tc_operand_plain(dqp, M, 3 * D), qb, D, 3 * D,
                       _epi(_scatter_plain(d_res.data_ptr(), M, D), accumulate=1))
        s_q = ctx.qkv_sink
        if s_q is not None:
            # element (d, j H dh + h dh + a) of the fused gradient is grad(w_{q,k,v})[h, d, a], and the
            # three (H, D, dh) gradients sit back to back in the bucket: one grouped-column scatter
            gemm_tc_wgrad(tc_operand_plain(xp, M, D), dqp, 3 * D, D, s_q, accumulate=True,
                          group=(dh, dh, D * dh))
            g = [None, None, None]
        else:
            dWg = torch.empty((D, 3 * D), dtype=_f32, device=dev)
            gemm_tc_wgrad(tc_operand_plain(xp, M, D), dqp, 3 * D, D, dWg)
            g = [dWg[:, j * D:(j + 1) * D].view(D, H, dh).permute(1, 0, 2) for j in range(3)]
        return (d_res, g[0], g[1], g[2], dWo, None, dg, db) + (None,) * 12


class _FFNBlockFn(torch.autograd.Function):
    """y = LayerNorm(x + dropout(FFN(x)))   transformer.py:57-59 as ONE autograd node.  The hidden
    activation and its gradient exist only as operand planes (see _FFNNativeFn); in the backward
    the residual gradient from the LayerNorm kernel is the accumulator of the last data-gradient
    GEMM, and every parameter gradient lands in the gradient bucket when one is registered."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, p_ffn, p_res, seed, site, eps, w1f, w1t, w2f, w2t):
        _chk(x, "x")
        M, K = x.shape
        F_ = w1.shape[0]
        dev = x.device
        need_bwd = any(ctx.needs_input_grad[:7])
        xp = planes_of(x)
        hp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
        # 1 bit per hidden unit (h > 0: through the ReLU and kept by dropout) for the data gradient,
        # written by the same epilogue: 6 MB instead of re-reading the 98 MB hi plane
        hbits = (torch.empty((M, F_ // 8), dtype=torch.uint8, device=dev)
                 if need_bwd and F_ % 8 == 0 and os.environ.get("SSB_MASKBITS", "1") != "0" else None)
        gemm_tc_kmajor(tc_operand_plain(xp, M, K), w1f, F_, K,
                       _epi(_scatter_plain(None, M, F_), bias=b1, relu=1, drop_p=p_ffn, seed=seed,
                            site=site, planes_out=hp, mask_bits_out=hbits))
        f = torch.empty((M, K), dtype=_f32, device=dev)
        gemm_tc_kmajor(tc_operand_plain(hp, M, F_), w2f, K, F_, _epi(_scatter_plain(f.data_ptr(), M, K), bias=b2))
        y, z, stat = _ln_fwd(x, f, gamma, beta, p_res, seed, site + 1, eps, need_bwd)
        if need_bwd:
            ctx.save_for_backward(xp, hp, z, stat, gamma, w1t, w2t, w1, w2, hbits)
        ctx.cfg = (p_ffn, p_res, seed, site)
        ctx.sinks = tuple(_sink(t) for t in (w1, b1, w2, b2, gamma, beta))
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, hp, z, stat, gamma, w1t, w2t, w1, w2, hbits = ctx.saved_tensors
        p_ffn, p_res, seed, site = ctx.cfg
        s_w1, s_b1, s_w2, s_b2, s_g, s_b = ctx.sinks
        M, K = z.shape
        F_ = hp.shape[2]
        dev = z.device
        scale = 1.0 / (1.0 - p_ffn) if p_ffn > 0 else 1.0
        d_res, d_f, d_fp, dg, db = _ln_bwd(dy, z, stat, gamma, p_res, seed, site + 1, (s_g, s_b),
                                           planes_only=True)
        if d_fp is None:
            d_fp = split_planes(d_f)
        dW2 = s_w2 if s_w2 is not None else torch.empty_like(w2)          # (K, F) = d_f^T h
        gemm_tc_wgrad(tc_operand_plain(d_fp, M, K), hp, F_, K, dW2, accumulate=s_w2 is not None)
        db2 = (colsum(d_f, out=s_b2, accumulate=s_b2 is not None) if d_f is not None
               else colsum_planes(d_fp, out=s_b2, accumulate=s_b2 is not None))
        dhp = torch.empty((2, M, F_), dtype=torch.bfloat16, device=dev)
        gemm_tc_kmajor(tc_operand_plain(d_fp, M, K), w2t, F_, K,
                       _epi(_scatter_plain(None, M, F_), mask_scale=scale, planes_out=dhp,
                            mask_planes=hp[0] if hbits is None else None, mask_bits=hbits))
        dW1 = s_w1 if s_w1 is not None else torch.empty_like(w1)          # (F, K) = dh^T x
        gemm_tc_wgrad(tc_operand_plain(dhp, M, F_), xp, K, F_, dW1, accumulate=s_w1 is not None)
        db1 = colsum_planes(dhp, out=s_b1, accumulate=s_b1 is not None)
        # x receives d_res (through the LayerNorm) + dh W1: accumulated by the GEMM epilogue
        gemm_tc_kmajor(tc_operand_plain(dhp, M, F_), w1t, K, F_,
                       _epi(_scatter_plain(d_res.data_ptr(), M, K), accumulate=1))
        return (d_res, None if s_w1 is not None else dW1, None if s_b1 is not None else db1,
                None if s_w2 is not None else dW2, None if s_b2 is not None else db2, dg, db) + \
            (None,) * 9


def attn_block_ok(M, B, T, H, dh, W, D):
    """Shapes for which _AttnBlockFn's kernels all apply (else the composition of single ops)."""
    return (_fused_attn_ok(B, T, H, dh, W) and _tc_fwd_ok(M, 3 * D, D) and _tc_fwd_ok(M, D, 3 * D)
            and _tc_wgrad_ok(M, 3 * D, D) and _tc_wgrad_ok(M, D, D) and _want_planes(M, D)
            and os.environ.get("SSB_BLOCKS", "1") != "0")


def ffn_block_ok(M, K, F_):
    return (_tc_fwd_ok(M, F_, K) and _tc_fwd_ok(M, K, F_) and _tc_wgrad_ok(M, F_, K)
            and _tc_wgrad_ok(M, K, F_) and K % 8 == 0 and _want_planes(M, K)
            and os.environ.get("SSB_BLOCKS", "1") != "0")


def _fused_attn_ok(B, T, H, dh, W):
    return _tc_enabled() and os.environ.get("SSB_ATTN", "fused") == "fused" and \
        dh in (32, 64, 96) and W <= 99 and 64 <= T <= 1024 and B * H <= 65535


def _tc_attn_ok(T, dh, W):
    return _tc_enabled() and os.environ.get("SSB_ATTN", "fused") != "simt" and dh % 4 == 0 and \
        dh <= _HP and 64 <= T <= 1024 and W <= 127


def band_attention(qkv, E_pad, B, T, H, dh, W, p=0.0, seed=0, site=0):
    """Banded relative-position attention: the fused tcgen05 kernels (csrc/attn_fused.cu) when the
    shape allows, else the multi-kernel tensor-core schedule (csrc/attn_tc.cu), else the CUDA-core
    band kernels (csrc/attn.cu).  SSB_ATTN=fused|tc|simt caps the choice (A/B testing)."""
    if _fused_attn_ok(B, T, H, dh, W):
        return _FusedAttnFn.apply(qkv, E_pad, int(B), int(T), int(H), int(dh), int(W), float(p),
                                  int(seed), int(site))
    if _tc_attn_ok(T, dh, W):
        return _DenseTCAttnFn.apply(qkv, E_pad, int(B), int(T), int(H), int(dh), int(W), float(p),
                                    int(seed), int(site))
    return _BandAttnFn.apply(qkv, E_pad, int(B), int(T), int(H), int(dh), int(W), float(p),
                             int(seed), int(site))

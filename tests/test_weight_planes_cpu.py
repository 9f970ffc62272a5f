"""CPU: the table that drives csrc/prep.cu (silent_speech_b200/weights.py).  The kernel's indexing
rule — view(r, c) = src[(r / RL) * s_rhi + (r % RL) * s_rlo + (c / CL) * s_chi + (c % CL) * s_clo],
written to dst_n[r * ld_n + c] and dst_t[c * ld_t + r] as hi / lo bf16 planes — is emulated here
with numpy over the REAL table, and every layout is compared with the torch expression the
round-1 code derived it with (`.t().contiguous()`, cat of permuted views, stacked conv taps)."""
import numpy as np
import torch
from absl import flags


def _emulate(wp):
    arena = np.zeros(wp.arena.numel(), dtype=np.float32)      # hi + lo summed, per plane slot
    hi = torch.zeros(wp.arena.numel(), dtype=torch.bfloat16)
    base = wp.arena.data_ptr()
    params = {p.data_ptr(): p for p in wp.model.parameters()}
    for i in range(wp.n_entries):
        e = wp.table_ctypes[i]
        src = None
        for ptr, p in params.items():
            if ptr <= e.src < ptr + 4 * p.numel():
                src, off0 = p.detach().reshape(-1), (e.src - ptr) // 4
        assert src is not None
        r = np.arange(e.rows)[:, None]
        c = np.arange(e.cols)[None, :]
        idx = off0 + (r // e.RL) * e.s_rhi + (r % e.RL) * e.s_rlo + (c // e.CL) * e.s_chi + (c % e.CL) * e.s_clo
        assert idx.min() >= 0 and idx.max() < src.numel()
        v = src[torch.from_numpy(idx.reshape(-1))].reshape(e.rows, e.cols)
        h = v.bfloat16()
        lo = (v - h.float()).bfloat16()
        if e.dst_n:
            o = (e.dst_n - base) // 2
            d = (r * e.ld_n + c).reshape(-1)
            hi[torch.from_numpy(o + d)] = h.reshape(-1)
            hi[torch.from_numpy(o + e.plane_n + d)] = lo.reshape(-1)
        if e.dst_t:
            o = (e.dst_t - base) // 2
            d = (c * e.ld_t + r).reshape(-1)
            hi[torch.from_numpy(o + d)] = h.reshape(-1)
            hi[torch.from_numpy(o + e.plane_t + d)] = lo.reshape(-1)
    wp.arena.copy_(hi)


def _value(planes):
    return planes[0].float() + planes[1].float()


def test_every_layout_matches_the_torch_derivation():
    from silent_speech_b200 import architecture as A
    from silent_speech_b200.weights import WeightPlanes
    F = flags.FLAGS
    if not F.is_parsed():
        F(["t"])
    F.model_size, F.num_layers, F.dropout = 128, 2, 0.0
    torch.manual_seed(0)
    m = A.Model(112, 80, 48)
    wp = WeightPlanes(m)
    wp._build()
    assert wp.n_entries > 20 and wp.tiles > 0
    _emulate(wp)
    tol = dict(rtol=0, atol=2e-5)          # hi + lo reproduces fp32 to ~2^-17 relative

    def check(owner, kind, want):
        got = wp.get(owner, kind)
        assert got is not None, kind
        assert tuple(got.shape[1:]) == tuple(want.shape), (kind, got.shape, want.shape)
        assert torch.allclose(_value(got), want.detach(), **tol), kind

    lins = [m.w_raw_in] + [l.linear1 for l in m.transformer.layers]
    lins += [l.linear2 for l in m.transformer.layers]
    for lin in lins:
        check(lin.weight, "f", lin.weight)
        check(lin.weight, "b", lin.weight.t())
    check(m.w_out.weight, "f", m.w_out.weight)          # 80-wide head: forward operand only
    assert wp.get(m.w_out.weight, "b") is None
    stacked = torch.cat([m.w_out.weight, m.w_aux.weight], 0)     # both heads as one 128-wide matrix
    check(m, "heads_f", stacked)
    check(m, "heads_b", stacked.t())
    D = 128
    for layer in m.transformer.layers:
        at = layer.self_attn
        Wg = at.qkv_weight()                              # (D, 3D) as round 1 built it
        check(at, "qkv_b", Wg)
        check(at, "qkv_f", Wg.t())
        check(at.w_o, "b", at.w_o.reshape(D, D))
        check(at.w_o, "f", at.w_o.reshape(D, D).t())
    for bi, blk in enumerate(m.conv_blocks):
        for conv in (blk.conv1, blk.conv2, blk.residual_path):
            Cout, Cin, k = conv.weight.shape
            if Cin == 8:
                assert wp.get(conv.weight, "conv_f") is None
                continue
            Wg = A._conv_weight(conv)                     # (k*Cin, Cout)
            check(conv.weight, "conv_f", Wg.t())
            W3 = Wg.view(k, Cin, Cout)
            tms = {(3, 1): [(2, 1, 0)], (3, 2): [(1,), (2, 0)], (1, 2): [(0,)]}[(k, conv.stride[0])]
            for tm in tms:
                Bd = torch.stack([W3[t] for t in tm], dim=1).reshape(Cin, len(tm) * Cout)
                check(conv.weight, ("conv_d", tm), Bd)
    # nothing overlaps: two destinations never share arena elements
    spans = sorted((v.data_ptr(), v.data_ptr() + 2 * v.numel()) for v in wp.views.values())
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0

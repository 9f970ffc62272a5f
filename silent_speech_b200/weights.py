"""Operand planes of a Model's weights, all layouts, refreshed by ONE kernel launch per step.

The tensor-core GEMMs (csrc/gemm_tc.cu) consume weights as bf16 hi/lo split planes in a layout
that depends on the use: `[N][K]` for a forward / data-gradient GEMM whose reduction runs over the
weight's K, the transpose for the other direction, per-tap-subset matrices for the strided
convolutions' data gradients, one fused `[3D][D]` matrix for the three attention projections.
Round 1 derived each of them at every use with torch copies and a split pass.  `WeightPlanes`
allocates one arena holding every layout of every weight, describes each as a strided view of
the parameter (ssb_prep_entry_t, csrc/prep.cu) and fills the whole arena with a single launch;
the parameters themselves keep the reference's names and shapes (checkpoint contract).
"""
import ctypes

import torch

from . import _lib

_EPOCH = [0]     # bumped by optimisers that write parameters through raw pointers (FlatAdamW)


def bump_epoch():
    _EPOCH[0] += 1


def _ok(n, k):
    """tcgen05 eligibility of a GEMM with output width n and reduction k (functional._tc_fwd_ok)."""
    return k % 64 == 0 and n % 8 == 0


class WeightPlanes:
    def __init__(self, model):
        self.model = model
        self._ptrs = None
        self._stamp = None
        self.views = {}          # (id(param), kind) -> (2, rows, cols) bf16 view into the arena
        self.n_entries = 0

    # ---- table construction -------------------------------------------------------------------
    def _specs(self):
        """[(param, kind_n, kind_t, view spec)] for every weight the tcgen05 engine can take."""
        m = self.model
        out = []

        def lin(w):            # nn.Linear weight (N, K): "f" = [N][K], "b" = [K][N]
            N, K = w.shape
            kn = "f" if _ok(N, K) else None       # forward: reduction over K
            kt = "b" if _ok(K, N) else None       # data gradient: reduction over N
            if kn or kt:
                out.append(dict(p=w, rows=N, cols=K, RL=N, CL=K, s=(0, K, 0, 1), kn=kn, kt=kt))

        lin(m.w_raw_in.weight)
        lin(m.w_out.weight)
        if getattr(m, "has_aux_out", False):
            lin(m.w_aux.weight)
            # both heads as ONE matrix (rows: w_out, then w_aux): 80 + 48 = 128 outputs make the
            # forward, data-gradient and weight-gradient GEMMs tensor-core eligible together
            n1, n2, K = m.w_out.weight.shape[0], m.w_aux.weight.shape[0], m.w_out.weight.shape[1]
            if _ok(n1 + n2, K) and _ok(K, n1 + n2) and n1 % 8 == 0:
                for w, r0 in ((m.w_out.weight, 0), (m.w_aux.weight, n1)):
                    out.append(dict(p=w, rows=w.shape[0], cols=K, RL=w.shape[0], CL=K, s=(0, K, 0, 1),
                                    kn="heads_f", kt="heads_b", group=m, stack=(r0, n1 + n2)))
        for layer in m.transformer.layers:
            lin(layer.linear1.weight)
            lin(layer.linear2.weight)
            at = layer.self_attn
            H, D, dh = at.w_q.shape
            if _ok(3 * D, D) and D % 128 == 0:
                for j, w in enumerate((at.w_q, at.w_k, at.w_v)):
                    # rows (h, a), cols k: "qkv_f" = [3D][D] rows j*D.., "qkv_b" = [D][3D] cols j*D..
                    out.append(dict(p=w, rows=D, cols=D, RL=dh, CL=D, s=(D * dh, 1, 0, dh),
                                    kn="qkv_f", kt="qkv_b", group=at, part=j, parts=3))
                # w_o (H, dh, D) is the GEMM-layout matrix [K = (h, a)][N = D]
                out.append(dict(p=at.w_o, rows=D, cols=D, RL=D, CL=D, s=(0, D, 0, 1), kn="b", kt="f"))
        for blk in m.conv_blocks:
            for conv in (blk.conv1, blk.conv2, blk.residual_path):
                if conv is None:
                    continue
                w = conv.weight
                Cout, Cin, k = w.shape
                if not (Cin % 64 == 0 and Cout % 8 == 0):
                    continue                      # the 8-channel first convolution: CUDA-core path
                # forward operand [Cout][(tap, ci)]
                out.append(dict(p=w, rows=Cout, cols=k * Cin, RL=Cout, CL=Cin, s=(0, Cin * k, 1, k),
                                kn="conv_f", kt=None))
                stride = conv.stride[0]
                tapmaps = ({(3, 1): [(2, 1, 0)], (3, 2): [(1,), (2, 0)], (1, 2): [(0,)]}
                           .get((k, stride), []))
                for tm in tapmaps:                # data-gradient operands [Cin][(j, co)]
                    t0 = tm[0]
                    dt = (tm[1] - tm[0]) if len(tm) > 1 else 0
                    out.append(dict(p=w, off=t0, rows=Cin, cols=len(tm) * Cout, RL=Cin, CL=Cout,
                                    s=(0, k, dt, Cin * k), kn=("conv_d", tm), kt=None))
        return out

    def _build(self):
        lib = _lib.load()
        specs = self._specs()
        dev = next(self.model.parameters()).device
        # arena layout: one (2, R, C) block per destination; fused QKV blocks are shared by 3 entries
        blocks, total = {}, 0

        def block(key, R, C):
            nonlocal total
            if key not in blocks:
                blocks[key] = (total, R, C)
                total += 2 * R * C
                total = (total + 63) // 64 * 64          # 128 B aligned blocks
            return blocks[key]

        plan = []
        for sp in specs:
            p, R, C = sp["p"], sp["rows"], sp["cols"]
            parts, part = sp.get("parts", 1), sp.get("part", 0)
            r_off, r_tot = sp.get("stack", (part * R, R * parts))   # row offset / rows of the stack
            owner = id(sp.get("group", p))
            dn = dt_ = None
            if sp["kn"] is not None:     # as read: parts stack along rows
                dn = (block((owner, sp["kn"]), r_tot, C), r_off * C, C)
            if sp["kt"] is not None:     # transposed (C x R): parts stack along columns
                dt_ = (block((owner, sp["kt"]), C, r_tot), r_off, r_tot)
            plan.append((sp, dn, dt_))
        self.arena = torch.zeros(max(total, 64), dtype=torch.bfloat16, device=dev)
        base = self.arena.data_ptr()
        table = (_lib.PrepEntry * max(len(plan), 1))()
        self.views = {}
        for i, (sp, dn, dt_) in enumerate(plan):
            e = table[i]
            p = sp["p"]
            e.src = p.data_ptr() + 4 * sp.get("off", 0)
            e.rows, e.cols, e.RL, e.CL = sp["rows"], sp["cols"], sp["RL"], sp["CL"]
            e.s_rhi, e.s_rlo, e.s_chi, e.s_clo = sp["s"]
            for which, d in (("n", dn), ("t", dt_)):
                if d is None:
                    continue
                (off, R, C), sub, ld = d
                ptr = base + 2 * (off + sub)
                if which == "n":
                    e.dst_n, e.plane_n, e.ld_n = ptr, R * C, ld
                else:
                    e.dst_t, e.plane_t, e.ld_t = ptr, R * C, ld
                kind = sp["kn"] if which == "n" else sp["kt"]
                owner = sp.get("group", p)
                self.views[(id(owner), kind)] = self.arena[off:off + 2 * R * C].view(2, R, C)
        tiles = lib.ssb_prep_plan(table, len(plan))
        if tiles < 0:
            raise _lib.SSBError(tiles, _lib.last_error())
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self.table_ctypes = table
        self.table_host = raw.pin_memory() if dev.type == "cuda" else raw
        self.table = self.table_host.to(dev, non_blocking=True)
        self.n_entries, self.tiles = len(plan), int(tiles)
        self._ptrs = self._param_ptrs()
        self._stamp = None

    def _param_ptrs(self):
        return tuple(p.data_ptr() for p in self.model.parameters())

    def _versions(self):
        return (sum(p._version for p in self.model.parameters()), _EPOCH[0])

    # ---- per-forward refresh ------------------------------------------------------------------
    def refresh(self):
        """Make every plane current.  Eagerly this is skipped while no parameter changed (in-place
        torch updates bump tensor versions, FlatAdamW bumps the module epoch); under CUDA-graph
        capture it is always recorded, so every replay re-derives the planes from the weights the
        optimiser just wrote."""
        capturing = torch.cuda.is_current_stream_capturing()
        if self._ptrs != self._param_ptrs():
            if capturing:
                raise RuntimeError("WeightPlanes: parameter storage changed; run one eager forward "
                                   "before capturing a CUDA graph")
            self._build()
        stamp = self._versions()
        if not capturing and stamp == self._stamp:
            return
        if self.n_entries:
            lib = _lib.load()
            _lib.check(lib.ssb_prep_planes(self.table.data_ptr(), self.n_entries, self.tiles,
                                           _lib.current_stream()))
        self._stamp = None if capturing else stamp

    def get(self, owner, kind):
        """(2, rows, cols) planes of `owner` (a parameter, or the attention module for the fused
        QKV matrices) in layout `kind`, or None when that weight is not tensor-core eligible."""
        return self.views.get((id(owner), kind))

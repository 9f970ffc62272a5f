"""Drop-in mirror of the reference's align.py (align.py:5-34) on the B200 DTW kernels.

Same names, same argument meaning, same return types:
  time_warp(costs) -> np.ndarray              (align.py:5-14)
  align_from_distances(distance_matrix, debug=False) -> list[int]   (align.py:16-34)
plus the batched device-side entry points the reference lacks
  align_batch(cost) / time_warp_batch(cost)   torch CUDA tensors, no host round trip.

All arithmetic runs in csrc/dtw.cu through the C ABI (ssb_dtw_*); there is no CPU path.
"""
import numpy as np
import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise _lib.SSBError(-3, "align: a CUDA device is required (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _prepare(cost):
    """-> (tensor, npairs, pair_stride, N, M, stride_i, stride_j) for the C ABI."""
    _lib.require_cuda(cost, "cost")
    if cost.dtype not in (torch.float32, torch.float64):
        raise TypeError("DTW accumulates in the caller's floating dtype like align.py:6 "
                        "(float32 or float64); got %s" % cost.dtype)
    t = cost.unsqueeze(0) if cost.dim() == 2 else cost
    if t.dim() != 3:
        raise ValueError("cost must be (N, M) or (P, N, M)")
    P, N, M = t.shape
    sp, si, sj = t.stride()
    ok = (si == 1 or sj == 1) and N > 1 and M > 1 and (P <= 1 or sp >= N * M)
    if not ok:  # degenerate vectors or exotic views: one contiguous copy
        t = t.contiguous()
        sp, si, sj = N * M, M, 1
    return t, P, sp, N, M, si, sj


def align_batch(cost, return_dtw=False):
    """DTW-align a batch of cost matrices on the GPU.

    cost: CUDA fp32 tensor (P, N, M) or (N, M), any view with one unit stride among the last
    two dims (e.g. `costs.transpose(-1, -2)` exactly as transduction_model.py:126 builds it).
    Returns int32 CUDA tensor (P, N): path[p, i] = matched column for row i
    (semantics of align.py:16-34), and optionally the accumulated-cost matrices.
    """
    lib = _lib.load()
    t, P, sp, N, M, si, sj = _prepare(cost)
    path = torch.empty((P, N), dtype=torch.int32, device=t.device)
    if P == 0:
        return (path, torch.empty_like(t)) if return_dtw else path
    if t.dtype == torch.float64:     # interface parity (align.py:6 follows the input dtype)
        dtw = torch.empty_strided(t.shape, t.stride(), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.ssb_dtw_time_warp_batch_f64(t.data_ptr(), P, sp, N, M, si, sj,
                                                       dtw.data_ptr(), path.data_ptr(),
                                                       _lib.current_stream()))
        return (path, dtw) if return_dtw else path
    ws_bytes = lib.ssb_dtw_workspace_bytes(P, N, M, si, sj)
    if ws_bytes < 0:
        raise _lib.SSBError(ws_bytes, _lib.last_error())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        if return_dtw:
            dtw = torch.empty_strided(t.shape, t.stride(), dtype=torch.float32, device=t.device)
            _lib.check(lib.ssb_dtw_time_warp_batch(t.data_ptr(), P, sp, N, M, si, sj,
                                                   dtw.data_ptr(), path.data_ptr(), ws.data_ptr(),
                                                   ws_bytes, _lib.current_stream()))
            return path, dtw
        _lib.check(lib.ssb_dtw_align_batch(t.data_ptr(), P, sp, N, M, si, sj, path.data_ptr(),
                                           ws.data_ptr(), ws_bytes, _lib.current_stream()))
    return path


class RaggedPlan:
    """Host-side plan of ONE launch over pairs of different shapes (ssb_dtw_ragged_plan): matrix
    p is (M_p = T_pred rows) x (N_p = T_target columns) row-major with pitch_p, at float offset
    cost_off_p of one buffer, aligned as its `.T` view like transduction_model.py:126 does."""

    def __init__(self, N, M, cost_off, pitch, device):
        import ctypes
        lib = _lib.load()
        n = len(N)
        arr = lambda v: (ctypes.c_int64 * max(n, 1))(*[int(x) for x in v])
        table = (_lib.DtwPair * max(n, 1))()
        dims = (ctypes.c_int64 * 2)()
        ws = lib.ssb_dtw_ragged_plan(n, arr(N), arr(M), arr(cost_off), arr(pitch), table, dims)
        if ws < 0:
            raise _lib.SSBError(ws, _lib.last_error())
        self.npairs, self.workspace_bytes = n, int(ws)
        self.max_N, self.max_M = int(dims[0]), int(dims[1])
        self.vectorized = all(int(o) % 4 == 0 for o in cost_off) and all(int(q) % 4 == 0 for q in pitch)
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self.table_host = raw.pin_memory() if torch.cuda.is_available() else raw   # kept alive: an
        self.table = self.table_host.to(device, non_blocking=True)   # async H2D may be captured

    def run(self, cost_base):
        """cost_base: 1-D CUDA fp32 buffer holding every matrix.  -> (npairs, max_N) int32 paths
        (rows >= N_p are 0)."""
        lib = _lib.load()
        _lib.require_cuda(cost_base, "cost")
        path = torch.empty((self.npairs, max(self.max_N, 1)), dtype=torch.int32, device=cost_base.device)
        if self.npairs == 0:
            return path
        ws = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=cost_base.device)
        with torch.cuda.device(cost_base.device):
            _lib.check(lib.ssb_dtw_align_ragged(cost_base.data_ptr(), self.npairs, self.table.data_ptr(),
                                                self.max_N, self.max_M,
                                                int(self.vectorized and cost_base.data_ptr() % 16 == 0),
                                                path.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.current_stream()))
        return path


def align_ragged(costs):
    """Align a list of cost matrices of DIFFERENT shapes in one launch.  costs[p]: CUDA fp32
    (T_pred_p, T_target_p) tensor exactly as transduction_model.py:116-124 builds `costs`; the
    DTW runs on its transpose (:126).  Returns a list of int32 CUDA tensors, path_p[t] = predicted
    frame aligned to target frame t (align.py:16-34 semantics)."""
    if not costs:
        return []
    dev = costs[0].device
    N, M, off, pitch, flat = [], [], [], [], []
    cur = 0
    for c in costs:
        _lib.require_cuda(c, "cost")
        if c.dtype != torch.float32 or c.dim() != 2:
            raise TypeError("align_ragged: 2-D float32 matrices")
        Tp, Tg = c.shape
        q = (Tg + 3) // 4 * 4
        buf = torch.zeros((Tp, q), dtype=torch.float32, device=dev)
        buf[:, :Tg] = c
        N.append(Tg), M.append(Tp), off.append(cur), pitch.append(q)
        flat.append(buf.view(-1))
        cur += Tp * q
    plan = RaggedPlan(N, M, off, pitch, dev)
    path = plan.run(torch.cat(flat))
    return [path[i, :N[i]] for i in range(len(costs))]


def _to_device(a):
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("expected a 2-D cost matrix")
    if a.dtype not in (np.float32, np.float64):
        if a.dtype.kind not in "fiu":
            raise TypeError("expected a real-valued cost matrix; got %s" % a.dtype)
        a = a.astype(np.float64)      # numba's zeros_like would keep an integer dtype: not supported
    dev = _device()
    if a.flags.c_contiguous or a.size == 0:
        return torch.from_numpy(a).to(dev, non_blocking=False)
    if a.flags.f_contiguous:  # the reference's costs.T view: ship the underlying C block
        return torch.from_numpy(a.T).to(dev, non_blocking=False).t()
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)


def time_warp(costs):
    """Accumulated-cost matrix, as align.py:5-14 returns it (numpy in, numpy out)."""
    t = _to_device(costs)
    _, dtw = align_batch(t, return_dtw=True)
    out = dtw[0].cpu().numpy() if dtw.dim() == 3 else dtw.cpu().numpy()
    return out


def align_from_distances(distance_matrix, debug=False):
    """For each row of the (N, M) matrix, the matched column under monotonic alignment.

    Mirrors align.py:16-34: numpy matrix in, Python list of N ints out; `debug=True`
    additionally plots the path when matplotlib is importable.
    """
    t = _to_device(distance_matrix)
    path = align_batch(t)[0].cpu().tolist()
    if debug:
        try:
            import matplotlib.pyplot as plt
            visual = np.zeros(np.asarray(distance_matrix).shape, dtype=np.float32)
            visual[range(len(path)), path] = 1
            plt.matshow(visual)
            plt.show()
        except ImportError:
            pass
    return path

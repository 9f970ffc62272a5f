"""Generate tests/golden/dtw_golden.npz by running the reference's align.py (numba) itself.

Run here (needs /root/reference):  python tests/golden/make_golden_dtw.py
Inputs are stored (small cases) or re-derivable from a stored seed (large cases), together
with the reference's outputs: the path list and, for small cases, the full dtw matrix.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _reference_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dtw_golden.npz")


def make_case(kind, seed, N, M):
    """Deterministic input for case (kind, seed, N, M): returns the (N, M) matrix AS PASSED to
    align_from_distances (possibly an F-ordered view, like transduction_model.py:126)."""
    rs = np.random.RandomState(seed)
    if kind == "randn_T":      # |randn| costs, F-ordered view of a (M, N) block  (the real call)
        return np.abs(rs.randn(M, N)).astype(np.float32).T
    if kind == "randn_C":      # C-contiguous
        return np.abs(rs.randn(N, M)).astype(np.float32)
    if kind == "ties_T":       # integer-valued costs: heavy ties, exercises first-wins order
        return rs.randint(0, 4, size=(M, N)).astype(np.float32).T
    if kind == "ties_C":
        return rs.randint(0, 4, size=(N, M)).astype(np.float32)
    if kind == "ones_C":
        return np.ones((N, M), dtype=np.float32)
    if kind == "signed_T":     # negative costs too
        return rs.randn(M, N).astype(np.float32).T
    if kind == "inf_T":        # some +inf entries
        a = np.abs(rs.randn(M, N)).astype(np.float32)
        a[rs.rand(M, N) < 0.05] = np.inf
        return a.T
    if kind == "cdist_T":      # what dtw_loss really feeds: cdist(pred, tgt).T
        import torch
        g = torch.Generator().manual_seed(seed)
        pred = torch.randn(M, 80, generator=g)
        tgt = torch.randn(N, 80, generator=g)
        return torch.cdist(pred[None], tgt[None])[0].numpy().T
    raise ValueError(kind)


CASES = [
    # (kind, seed, N, M, store_input_and_dtw)
    ("randn_T", 1, 6, 5, True), ("randn_C", 2, 5, 6, True), ("ones_C", 0, 6, 5, True),
    ("ones_C", 0, 5, 6, True), ("ties_T", 3, 17, 9, True), ("ties_C", 4, 9, 17, True),
    ("randn_T", 5, 1, 7, True), ("randn_T", 6, 7, 1, True), ("randn_C", 7, 1, 1, True),
    ("randn_T", 8, 2, 2, True), ("inf_T", 9, 33, 40, True), ("signed_T", 10, 64, 50, True),
    ("randn_T", 11, 129, 33, True), ("randn_C", 12, 33, 129, True),
    ("ties_T", 13, 130, 131, True), ("ties_C", 14, 257, 70, True),
    ("randn_T", 20, 600, 500, False), ("ties_T", 21, 600, 500, False),
    ("cdist_T", 22, 600, 500, False), ("randn_C", 23, 500, 600, False),
    ("randn_T", 24, 601, 503, False), ("ties_C", 25, 513, 259, False),
    ("cdist_T", 26, 750, 625, False),
]


def main():
    (align,) = import_reference("align")
    out = {}
    meta = []
    for idx, (kind, seed, N, M, store) in enumerate(CASES):
        a = make_case(kind, seed, N, M)
        assert a.shape == (N, M) and a.dtype == np.float32
        path = np.asarray(align.align_from_distances(a), dtype=np.int32)
        dtw = align.time_warp(a)
        out[f"path_{idx}"] = path
        # order-independent fingerprint of the full table (inf-safe)
        finite = np.isfinite(dtw)
        out[f"dtwsum_{idx}"] = np.array([np.sum(dtw[finite], dtype=np.float64), finite.sum()])
        out[f"dtwlast_{idx}"] = np.array(dtw[N - 1, M - 1], dtype=np.float32)
        if store:
            out[f"input_{idx}"] = np.ascontiguousarray(a)
            out[f"fortran_{idx}"] = np.array(a.flags.f_contiguous and not a.flags.c_contiguous)
            out[f"dtw_{idx}"] = dtw.astype(np.float32)
        meta.append(f"{kind},{seed},{N},{M},{int(store)}")
    out["meta"] = np.array(meta)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()

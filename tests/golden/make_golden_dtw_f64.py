"""Generate tests/golden/dtw_golden_f64.npz by running the reference's align.py (numba) on
FLOAT64 inputs (align.py:6 `zeros_like(costs)`: the accumulated-cost table follows the caller's
dtype).  Inputs are stored; outputs: path and full table.
Run here (needs /root/reference):  python tests/golden/make_golden_dtw_f64.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _reference_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dtw_golden_f64.npz")


def cases():
    rs = np.random.RandomState(77)
    yield "rand_C", np.abs(rs.randn(37, 41))
    yield "rand_T", np.abs(rs.randn(29, 53)).T
    yield "ties_T", rs.randint(0, 3, size=(40, 33)).astype(np.float64).T
    # costs that differ only below fp32 resolution: a float32 accumulation would tie, float64 does not
    a = np.ones((24, 24)) + rs.rand(24, 24) * 1e-9
    yield "sub_fp32_eps", a
    yield "row", np.abs(rs.randn(1, 9))
    yield "col", np.abs(rs.randn(9, 1))
    b = np.abs(rs.randn(30, 20))
    b[rs.rand(30, 20) < 0.1] = np.inf
    yield "inf", b


def main():
    (align,) = import_reference("align")
    out = {}
    for name, a in cases():
        assert a.dtype == np.float64
        dtw = align.time_warp(a)
        assert dtw.dtype == np.float64
        out[f"{name}_input"] = np.ascontiguousarray(a)
        out[f"{name}_fortran"] = np.array(int(a.flags.f_contiguous and not a.flags.c_contiguous))
        out[f"{name}_dtw"] = dtw
        out[f"{name}_path"] = np.array(align.align_from_distances(a), dtype=np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()

"""CPU: the C oracle (oracle/dtw_oracle.c) against golden vectors produced by the executed
reference align.py (tests/golden/make_golden_dtw.py).  Bit-exact on paths and dtw tables."""
import os

import numpy as np
import pytest

from make_golden_dtw import CASES, make_case
from oracle import dtw as odtw


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw_golden.npz"))


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_matches_reference_golden(golden, idx):
    kind, seed, N, M, store = CASES[idx]
    assert golden["meta"][idx] == f"{kind},{seed},{N},{M},{int(store)}"
    a = make_case(kind, seed, N, M)
    if store:  # regenerated input must equal the stored one (guards RNG drift)
        np.testing.assert_array_equal(np.ascontiguousarray(a), golden[f"input_{idx}"])
    path = odtw.align_from_distances(a)
    np.testing.assert_array_equal(np.asarray(path, np.int32), golden[f"path_{idx}"])
    dtw = odtw.time_warp(a)
    assert dtw.dtype == np.float32 and dtw.shape == (N, M)
    if store:
        np.testing.assert_array_equal(dtw, golden[f"dtw_{idx}"])
    fin = np.isfinite(dtw)
    assert fin.sum() == golden[f"dtwsum_{idx}"][1]
    assert np.sum(dtw[fin], dtype=np.float64) == golden[f"dtwsum_{idx}"][0]
    assert np.array_equal(dtw[N - 1, M - 1], golden[f"dtwlast_{idx}"]) or (N == 1 or M == 1)


def test_probed_examples():
    # SURVEY.md §8 a10 (probed on the reference): 6x5 ones -> [0,1,2,3,4,4]; 5x6 ones -> [0,1,2,3,4]
    assert odtw.align_from_distances(np.ones((6, 5), np.float32)) == [0, 1, 2, 3, 4, 4]
    assert odtw.align_from_distances(np.ones((5, 6), np.float32)) == [0, 1, 2, 3, 4]


def test_fp64_follows_input_dtype():
    rs = np.random.RandomState(0)
    a = np.abs(rs.randn(20, 30))
    d = odtw.time_warp(a)
    assert d.dtype == np.float64
    p = odtw.align_from_distances(a)
    assert p[0] == 0 and all(p[i] <= p[i + 1] for i in range(len(p) - 1))


def test_batch_matches_single():
    rs = np.random.RandomState(3)
    c = np.abs(rs.randn(7, 40, 50)).astype(np.float32)   # (P, M, N) blocks
    view = c.transpose(0, 2, 1)                            # (P, N=50, M=40) F-ordered pairs
    got = odtw.align_batch(view, threads=2)
    for p in range(7):
        assert got[p].tolist() == odtw.align_from_distances(view[p])


F64_CASES = ["rand_C", "rand_T", "ties_T", "sub_fp32_eps", "row", "col", "inf"]


def f64_case(g, name):
    a = g[f"{name}_input"]
    return a.T.copy().T if int(g[f"{name}_fortran"]) else a     # restore the F-ordered view


@pytest.mark.parametrize("name", F64_CASES)
def test_fp64_oracle_matches_reference_golden(golden_dir, name):
    """float64 in -> float64 accumulation (align.py:6): table and path bit-exact with the
    reference's numba run (tests/golden/make_golden_dtw_f64.py)."""
    g = np.load(os.path.join(golden_dir, "dtw_golden_f64.npz"))
    a = f64_case(g, name)
    assert a.dtype == np.float64
    d = odtw.time_warp(a)
    np.testing.assert_array_equal(d, g[f"{name}_dtw"])
    assert odtw.align_from_distances(a) == g[f"{name}_path"].tolist()

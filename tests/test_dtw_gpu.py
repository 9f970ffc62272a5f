"""GPU parity: csrc/dtw.cu through the C ABI vs the C oracle and the reference-generated
golden vectors.  Integer paths and fp32 accumulated-cost tables must be BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from make_golden_dtw import CASES, make_case
from oracle import dtw as odtw
from silent_speech_b200 import align

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw_golden.npz"))


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_golden_cases_dropin_api(golden, idx):
    kind, seed, N, M, store = CASES[idx]
    a = make_case(kind, seed, N, M)
    path = align.align_from_distances(a)
    assert isinstance(path, list) and all(isinstance(v, int) for v in path)
    np.testing.assert_array_equal(np.asarray(path, np.int32), golden[f"path_{idx}"])
    dtw = align.time_warp(a)
    assert dtw.shape == (N, M) and dtw.dtype == np.float32
    np.testing.assert_array_equal(dtw, odtw.time_warp(a))


SHAPES = [(2, 2), (3, 5), (5, 3), (4, 33), (33, 4), (127, 40), (128, 40), (129, 40), (40, 127),
          (40, 129), (16, 16), (17, 15), (31, 32), (32, 31), (47, 48), (255, 257), (300, 64),
          (130, 9), (9, 130), (1, 9), (9, 1), (1, 1)]


@pytest.mark.parametrize("layout", ["T", "C", "T_odd_pitch", "C_odd_pitch"])
def test_batched_random_and_ties_vs_oracle(layout):
    rs = np.random.RandomState(1234)
    for (N, M) in SHAPES:
        for ties in (False, True):
            P = 5
            if layout.startswith("T"):
                pitch = N + (3 if "odd" in layout else 0)
                block = np.zeros((P, M, pitch), np.float32)
                vals = rs.randint(0, 3, (P, M, N)) if ties else np.abs(rs.randn(P, M, N))
                block[:, :, :N] = vals
                host = block[:, :, :N].transpose(0, 2, 1)        # (P, N, M), stride_i == 1
                dev = torch.from_numpy(block).cuda()[:, :, :N].transpose(1, 2)
            else:
                pitch = M + (3 if "odd" in layout else 0)
                block = np.zeros((P, N, pitch), np.float32)
                vals = rs.randint(0, 3, (P, N, M)) if ties else np.abs(rs.randn(P, N, M))
                block[:, :, :M] = vals
                host = block[:, :, :M]
                dev = torch.from_numpy(block).cuda()[:, :, :M]
            want = odtw.align_batch(host, threads=2)
            got, dtw = align.align_batch(dev, return_dtw=True)
            np.testing.assert_array_equal(got.cpu().numpy(), want, err_msg=f"{layout} {N}x{M}")
            got2 = align.align_batch(dev)
            np.testing.assert_array_equal(got2.cpu().numpy(), want)
            d = dtw.cpu().numpy()
            for p in range(P):
                np.testing.assert_array_equal(d[p], odtw.time_warp(host[p]))


def test_full_size_properties_and_sampled_parity():
    """cfg-2 geometry (500x600 frames, DTW on the 600x500 transposed view) on 1024 pairs."""
    P, Tp, Tg = 1024, 500, 600
    g = torch.Generator(device="cpu").manual_seed(1234)
    pred = torch.randn(P, Tp, 80, generator=g)
    tgt = torch.randn(P, Tg, 80, generator=g)
    cost = torch.cdist(pred.cuda(), tgt.cuda())          # (P, 500, 600) as dtw_loss builds it
    view = cost.transpose(1, 2)                          # (P, 600, 500), what align sees
    path = align.align_batch(view)
    assert path.shape == (P, Tg) and path.dtype == torch.int32
    p = path.cpu().numpy()
    assert (p[:, 0] == 0).all()
    assert (np.diff(p, axis=1) >= 0).all(), "paths must be monotone"
    assert p.min() >= 0 and p.max() == Tp - 1 and (p[:, -1] == Tp - 1).all()
    # determinism / idempotence
    np.testing.assert_array_equal(align.align_batch(view).cpu().numpy(), p)
    # sampled bit-exact parity with the oracle on the identical matrices
    host = cost[:48].cpu().numpy().transpose(0, 2, 1)
    np.testing.assert_array_equal(p[:48], odtw.align_batch(host))
    # path cost equals the oracle's dtw[N-1, M-1] (sum along the path, same add order)
    d_last = odtw.time_warp(host[0])[-1, -1]
    _, dtw = align.align_batch(view[:1], return_dtw=True)
    assert dtw[0, -1, -1].item() == d_last


def test_inf_costs_and_errors():
    rs = np.random.RandomState(5)
    a = np.abs(rs.randn(60, 70)).astype(np.float32)
    a[rs.rand(60, 70) < 0.1] = np.inf
    assert align.align_from_distances(a) == odtw.align_from_distances(a)
    assert align.align_from_distances(a.T) == odtw.align_from_distances(a.T)
    with pytest.raises(TypeError):
        align.align_from_distances(a.astype(np.complex64))
    with pytest.raises(Exception):
        align.align_batch(torch.zeros(3, 4))  # CPU tensor: no CPU path


@pytest.mark.parametrize("name", ["rand_C", "rand_T", "ties_T", "sub_fp32_eps", "row", "col", "inf"])
def test_fp64_inputs_follow_the_reference_dtype(golden_dir, name):
    """align.py:6: float64 costs are accumulated and compared in float64.  Paths and tables
    bit-exact with the reference's own float64 run (dtw_golden_f64.npz) through the drop-in API."""
    import os
    g = np.load(os.path.join(golden_dir, "dtw_golden_f64.npz"))
    a = g[f"{name}_input"]
    if int(g[f"{name}_fortran"]):
        a = a.T.copy().T
    assert align.align_from_distances(a) == g[f"{name}_path"].tolist()
    d = align.time_warp(a)
    assert d.dtype == np.float64
    np.testing.assert_array_equal(d, g[f"{name}_dtw"])
    if name == "sub_fp32_eps":      # the float32 kernel on the rounded costs ties differently
        assert align.align_from_distances(a.astype(np.float32)) == odtw.align_from_distances(
            a.astype(np.float32))


def test_fp64_batch_on_device():
    rs = np.random.RandomState(12)
    c = torch.from_numpy(np.abs(rs.randn(5, 44, 38))).cuda()         # (P, M, N) float64 blocks
    p = align.align_batch(c.transpose(1, 2)).cpu().numpy()           # F-ordered pairs (38 x 44)
    for i in range(5):
        assert p[i].tolist() == odtw.align_from_distances(c[i].cpu().numpy().T)

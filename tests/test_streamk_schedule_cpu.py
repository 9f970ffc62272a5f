"""CPU: the stream-K remainder schedule of gemm_tc.cu (plan_streamk on the host, sk_range / get_item and
the owner's contributor walk in the kernel), restated line by line in Python and checked for the
properties the kernel relies on: every k-block of every tile is computed exactly once, every tile has
exactly one epilogue (full or owner) item, a worker writes at most one partial (one workspace slot per
worker), its partial piece comes before its owner piece (no wait chains), an owner finds exactly the
workers that hold the other pieces of its tile, and never more than the 6 its list can hold."""
import itertools

import pytest

IT_FULL, IT_PARTIAL, IT_OWNER = 0, 1, 2


def plan_streamk(total, workers, num_kb, have_ws=True):
    """gemm_tc.cu plan_streamk: -> (grid_workers, sk_full, sk_rem, sk_workers)"""
    grid_workers = min(total, workers)
    rem = total % workers
    if not have_ws or rem == 0 or rem * 4 > workers * 3 or num_kb < 16:
        return grid_workers, total, 0, 0
    sk_workers = min(rem * 4, workers)
    return max(grid_workers, sk_workers), total - rem, rem, sk_workers


def sk_range(w, sk_rem, sk_workers, num_kb):
    U = sk_rem * num_kb
    return w * U // sk_workers, (w + 1) * U // sk_workers


def items(worker, num_workers, total, sk_full, sk_rem, sk_workers, num_kb):
    """gemm_tc.cu get_item for idx = 0, 1, ...: list of (tile, kb0, nk, mode)"""
    out = []
    full = sk_full if sk_rem else total
    for tile in range(worker, full, num_workers):
        out.append((tile, 0, num_kb, IT_FULL))
    if not sk_rem or worker >= sk_workers:
        return out
    u0, u1 = sk_range(worker, sk_rem, sk_workers, num_kb)
    if u0 == u1:
        return out
    ta = u0 // num_kb
    a_last = (ta + 1) * num_kb
    has_b = u1 > a_last
    if has_b:
        out.append((sk_full + ta + 1, 0, u1 - a_last, IT_PARTIAL))
    a_end = a_last if has_b else u1
    kb0 = u0 - ta * num_kb
    mode = (IT_FULL if kb0 == 0 else IT_OWNER) if a_end == a_last else IT_PARTIAL
    out.append((sk_full + ta, kb0, a_end - u0, mode))
    return out


def contributors(worker, tile, sk_full, sk_rem, sk_workers, num_kb):
    t_first = (tile - sk_full) * num_kb
    out = []
    w = worker - 1
    while w >= 0 and len(out) < 6:
        u0, u1 = sk_range(w, sk_rem, sk_workers, num_kb)
        if u1 <= t_first:
            break
        if u0 != u1:
            out.append(w)
        w -= 1
    return out


SHAPES = [  # (rows, N, K) of the cfg-1 / cfg-5 / vocoder GEMMs
    (16000, 3072, 768), (16000, 768, 3072), (16000, 2304, 768), (16000, 768, 2304), (16000, 768, 768),
    (64000, 768, 2304), (32000, 768, 2304), (32000, 768, 768), (24000, 768, 3072), (24000, 2304, 768),
    (4800, 256, 768), (4800, 256, 2816), (38400, 128, 896), (601, 2048, 1024), (4801, 1024, 512),
]


def _check(total, workers, num_kb):
    grid, sk_full, sk_rem, sk_workers = plan_streamk(total, workers, num_kb)
    cover = {}
    epilogues = {}
    partial_owner = {}
    for w in range(grid):
        its = items(w, grid, total, sk_full, sk_rem, sk_workers, num_kb)
        modes = [m for *_, m in its]
        assert modes.count(IT_PARTIAL) <= 1
        if IT_PARTIAL in modes and IT_OWNER in modes:
            assert modes.index(IT_PARTIAL) < modes.index(IT_OWNER)
        assert all(m == IT_FULL for m in modes[:len(modes) - 2])
        for tile, kb0, nk, mode in its:
            assert nk > 0 and 0 <= kb0 and kb0 + nk <= num_kb and 0 <= tile < total
            for kb in range(kb0, kb0 + nk):
                assert (tile, kb) not in cover
                cover[(tile, kb)] = w
            if mode == IT_PARTIAL:
                partial_owner.setdefault(tile, []).append(w)
            else:
                assert tile not in epilogues
                epilogues[tile] = (w, mode, kb0, nk)
    assert len(cover) == total * num_kb
    assert set(epilogues) == set(range(total))
    for tile, (w, mode, kb0, nk) in epilogues.items():
        if mode == IT_FULL:
            assert kb0 == 0 and nk == num_kb and tile not in partial_owner
        else:
            assert kb0 + nk == num_kb
            got = contributors(w, tile, sk_full, sk_rem, sk_workers, num_kb)
            assert sorted(got) == sorted(partial_owner[tile]) and 1 <= len(got) <= 6
    return sk_rem


@pytest.mark.parametrize("rows,N,K", SHAPES)
def test_model_shapes(rows, N, K):
    for ctas, workers in ((2, 74), (1, 148), (2, 66)):
        total = -(-rows // (128 * ctas)) * -(-N // 256)
        _check(total, workers, K // 32)


def test_sweep():
    used = 0
    for total, workers, num_kb in itertools.product(list(range(1, 200)) + [567, 750, 756, 1125], (74, 148, 7),
                                                     (16, 17, 19, 24, 25, 72, 96)):
        used += 1 if _check(total, workers, num_kb) else 0
    assert used > 1000
    for num_kb in (1, 3, 15):          # short K: classic schedule
        assert plan_streamk(189, 74, num_kb)[2] == 0
    assert plan_streamk(189, 74, 96, have_ws=False)[2] == 0
    assert plan_streamk(74 * 3 + 70, 74, 96)[2] == 0     # nearly full last wave: not worth it

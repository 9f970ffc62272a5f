"""ORACLE — test infrastructure, not product code.

numpy-facing wrappers over oracle/dtw_oracle.c (the C restatement of align.py:5-34).
"""
import ctypes
import os

import numpy as np

from . import lib as _lib

_i64 = ctypes.c_int64


def _elem_strides(a):
    s0, s1 = a.strides
    it = a.itemsize
    assert s0 % it == 0 and s1 % it == 0
    return s0 // it, s1 // it


def time_warp(costs):
    """align.py:5-14 — accumulated-cost matrix, same dtype as the input (fp32 / fp64)."""
    a = np.asarray(costs)
    assert a.ndim == 2 and a.dtype in (np.float32, np.float64)
    N, M = a.shape
    si, sj = _elem_strides(a)
    out = np.empty((N, M), dtype=a.dtype)
    fn = _lib().ssb_oracle_time_warp_f32 if a.dtype == np.float32 else _lib().ssb_oracle_time_warp_f64
    fn(a.ctypes.data_as(ctypes.c_void_p), _i64(N), _i64(M), _i64(si), _i64(sj),
       out.ctypes.data_as(ctypes.c_void_p))
    return out


def align_from_distances(distance_matrix):
    """align.py:16-34 — list of N ints."""
    a = np.asarray(distance_matrix)
    assert a.ndim == 2 and a.dtype in (np.float32, np.float64)
    N, M = a.shape
    si, sj = _elem_strides(a)
    res = np.empty(N, dtype=np.int32)
    scratch = np.empty((N, M), dtype=a.dtype)
    fn = _lib().ssb_oracle_align_f32 if a.dtype == np.float32 else _lib().ssb_oracle_align_f64
    fn(a.ctypes.data_as(ctypes.c_void_p), _i64(N), _i64(M), _i64(si), _i64(sj),
       res.ctypes.data_as(ctypes.c_void_p), scratch.ctypes.data_as(ctypes.c_void_p))
    return res.tolist()


def align_batch(cost, threads=None):
    """Batched fp32 alignment of a (P, N, M) array (any view with element strides); all cores."""
    a = np.asarray(cost)
    assert a.ndim == 3 and a.dtype == np.float32
    P, N, M = a.shape
    it = a.itemsize
    sp, si, sj = (s // it for s in a.strides)
    path = np.empty((P, N), dtype=np.int32)
    if threads is None:
        threads = os.cpu_count() or 1
    rc = _lib().ssb_oracle_align_batch_f32(a.ctypes.data_as(ctypes.c_void_p), _i64(P), _i64(sp),
                                           _i64(N), _i64(M), _i64(si), _i64(sj),
                                           path.ctypes.data_as(ctypes.c_void_p),
                                           ctypes.c_int(int(threads)))
    if rc:
        raise MemoryError("oracle align_batch: allocation failed")
    return path

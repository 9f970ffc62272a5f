"""ctypes binding of libssb.so (the C ABI declared in include/ssb.h).

There is NO fallback: if the shared library is missing it is built with nvcc, and if that
is impossible the import of any kernel-backed function raises.  Nothing in this package
computes the hot path on the CPU.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssb.so")

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_u64 = ctypes.c_uint64
c_ptr = ctypes.c_void_p

c_int = ctypes.c_int
c_u32 = ctypes.c_uint32


class Gather(ctypes.Structure):
    """ssb_gather_t"""
    _fields_ = [("base", c_ptr), ("batch_stride", c_i64), ("rows_per_batch", c_i32), ("C", c_i32),
                ("L_src", c_i32), ("ld", c_i32), ("s_t", c_i32), ("s_tap", c_i32), ("off", c_i32)]


class Scatter(ctypes.Structure):
    """ssb_scatter_t"""
    _fields_ = [("base", c_ptr), ("batch_stride", c_i64), ("rows_per_batch", c_i32), ("ld", c_i32),
                ("d_t", c_i32), ("d_off", c_i32), ("batch_stride_hi", c_i64)]


class Epilogue(ctypes.Structure):
    """ssb_epilogue_t"""
    _fields_ = [("out", Scatter), ("bias", c_ptr), ("mask_src", c_ptr), ("mask_scale", c_f32),
                ("relu", c_i32), ("accumulate", c_i32), ("drop_p", c_f32), ("seed", c_u64),
                ("site", c_u32), ("planes_out", c_ptr), ("planes_stride", c_i64),
                ("mask_planes", c_ptr), ("mask_bits", c_ptr), ("mask_bits_out", c_ptr),
                ("planes_lrelu", c_i32), ("planes_neg_slope", c_f32)]


class TcOperand(ctypes.Structure):
    """ssb_tc_operand_t"""
    _fields_ = [("planes", c_ptr), ("plane_stride", c_i64), ("batch_stride", c_i64),
                ("batches", c_i32), ("rows_out", c_i32), ("L_src", c_i32), ("C", c_i32),
                ("ld", c_i32), ("s_t", c_i32), ("s_tap", c_i32), ("off", c_i32),
                ("batch_div", c_i32), ("reserved_", c_i32), ("batch_stride_hi", c_i64)]


class DtwPair(ctypes.Structure):
    """ssb_dtw_pair_t"""
    _fields_ = [("cost_off", c_i64), ("dirs_off", c_i64), ("pitch", c_i64), ("N", c_i32),
                ("M", c_i32), ("nbands", c_i32), ("nch", c_i32)]


class Utt(ctypes.Structure):
    """ssb_utt_t"""
    _fields_ = [("pred_row", c_i64), ("tgt_row", c_i64), ("cost_off", c_i64), ("Tp", c_i32),
                ("Tg", c_i32), ("pitch", c_i32), ("silent", c_i32), ("pair", c_i32),
                ("reserved_", c_i32)]


class PrepEntry(ctypes.Structure):
    """ssb_prep_entry_t"""
    _fields_ = [("src", c_ptr), ("dst_n", c_ptr), ("dst_t", c_ptr), ("plane_n", c_i64),
                ("plane_t", c_i64), ("s_rhi", c_i64), ("s_rlo", c_i64), ("s_chi", c_i64),
                ("s_clo", c_i64), ("rows", c_i32), ("cols", c_i32), ("RL", c_i32), ("CL", c_i32),
                ("ld_n", c_i32), ("ld_t", c_i32), ("tile0", c_i32), ("tiles_c", c_i32)]


class EmgRec(ctypes.Structure):
    """ssb_emg_rec_t"""
    _fields_ = [("off", c_i64), ("out_off", c_i64), ("n", c_i32), ("n_out", c_i32)]


_PT = ctypes.POINTER(TcOperand)
_PG = ctypes.POINTER(Gather)
_PE = ctypes.POINTER(Epilogue)

# name -> (restype, argtypes); mirrors include/ssb.h one to one (tests check the symbol list)
_SIGNATURES = {
    "ssb_version": (ctypes.c_int, []),
    "ssb_sizeof": (c_i64, [ctypes.c_int]),
    "ssb_last_error": (ctypes.c_char_p, []),
    "ssb_device_sm_count": (ctypes.c_int, []),
    "ssb_set_seed_source": (ctypes.c_int, [c_ptr]),
    "ssb_dtw_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_i64]),
    "ssb_dtw_align_batch": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr,
                                           c_ptr, c_i64, c_ptr]),
    "ssb_dtw_time_warp_batch": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                               c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ssb_dtw_time_warp_batch_f64": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                                   c_ptr, c_ptr, c_ptr]),
    "ssb_dtw_ragged_plan": (c_i64, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ssb_dtw_align_ragged": (ctypes.c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, ctypes.c_int, c_ptr,
                                            c_ptr, c_i64, c_ptr]),
    "ssb_dtw_cost_batch": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64,
                                          c_i64, c_i64, c_f32, c_ptr, c_ptr]),
    "ssb_dtw_loss_rows": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64,
                                         c_i64, c_i64, c_i64, c_f32, c_f32, c_ptr, c_ptr, c_ptr,
                                         c_ptr]),
    "ssb_prep_plan": (c_i64, [c_ptr, c_i64]),
    "ssb_prep_planes": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_ptr]),
    "ssb_mel_num_frames": (c_i64, [c_i64, ctypes.c_int, ctypes.c_int]),
    "ssb_mel_fwd": (ctypes.c_int, [c_ptr, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_f32, c_ptr,
                                   c_ptr, c_ptr]),
    "ssb_gemm_nn": (c_int, [_PG, c_ptr, c_i64, _PE, c_i64, c_i64, c_i64, c_ptr]),
    "ssb_gemm_nt": (c_int, [_PG, c_ptr, c_i64, c_i64, c_int, c_int, c_int, _PE, c_i64, c_i64,
                            c_i64, c_ptr]),
    "ssb_gemm_tn": (c_int, [_PG, c_ptr, c_i64, c_ptr, c_i64, c_int, c_i64, c_i64, c_i64, c_ptr]),
    "ssb_split_bf16": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "ssb_gemm_tc_kmajor": (c_int, [_PT, c_ptr, c_i64, c_i64, _PE, c_ptr]),
    "ssb_gemm_tc_streamk_workspace_bytes": (c_i64, []),
    "ssb_gemm_tc_set_streamk_workspace": (c_int, [c_ptr, c_i64]),
    "ssb_gemm_tc_wgrad": (c_int, [_PT, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_i64, c_i64, c_int,
                                  c_ptr]),
    "ssb_gemm_tc_batched": (c_int, [_PT, _PT, c_int, c_i64, c_i64, _PE, c_ptr]),
    "ssb_gemm_tc_batched_tn": (c_int, [_PT, _PT, c_i64, c_i64, _PE, c_ptr]),
    "ssb_pad_split_heads": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "ssb_transpose_split_heads": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                          c_ptr, c_ptr]),
    "ssb_attn_softmax_fwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                     c_f32, c_u64, c_u32, c_ptr, c_ptr]),
    "ssb_attn_ds_bwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32,
                                c_u64, c_u32, c_ptr, c_ptr, c_ptr]),
    "ssb_attn_fused_fwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32,
                                   c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ssb_attn_delta": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    "ssb_attn_fused_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64,
                                   c_i64, c_i64, c_i64, c_f32, c_u64, c_u32, c_ptr, c_i64, c_ptr,
                                   c_ptr, c_i64, c_i64, c_i64, c_ptr]),
    "ssb_col_partials_bytes": (c_i64, [c_i64, c_i64]),
    "ssb_colsum": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "ssb_split_bf16_t": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "ssb_split_bf16_2d": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr]),
    "ssb_ctc_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "ssb_ctc_loss_fused": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_i64,
                                   c_int, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ssb_adamw_flat": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_f32, c_f32, c_f32,
                               c_f32, c_f32, c_ptr]),
    "ssb_colsum_planes": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "ssb_bn_stats": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_int,
                             c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ssb_bn_apply": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64,
                             c_i64, c_ptr, c_ptr, c_ptr]),
    "ssb_bn_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_i64, c_ptr,
                           c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "ssb_bn_bwd2": (c_int, [c_ptr] * 10 + [c_int, c_i64, c_i64] + [c_ptr] * 8 + [c_int, c_ptr, c_i64,
                                                                                c_ptr]),
    "ssb_add_dropout_ln_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_f32, c_f32,
                                       c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ssb_add_dropout_ln_bwd_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "ssb_add_dropout_ln_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_f32,
                                       c_u64, c_u32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                       c_ptr, c_i64, c_ptr]),
    "ssb_band_attn_fwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32,
                                  c_u64, c_u32, c_ptr, c_ptr, c_ptr]),
    "ssb_band_attn_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                  c_f32, c_u64, c_u32, c_ptr, c_ptr, c_ptr]),
    "ssb_emg_filtfilt_workspace_bytes": (c_i64, [c_i64, c_i64, c_int]),
    "ssb_emg_filtfilt_chain": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr,
                                       c_int, c_ptr, c_i64, c_ptr]),
    "ssb_emg_subsample": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_int, ctypes.c_double, ctypes.c_double,
                                  c_ptr, c_int, c_ptr]),
    "ssb_voc_mix": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_f32, c_int, c_f32, c_ptr, c_ptr, c_i64,
                            c_ptr]),
    "ssb_voc_post": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_f32, c_f32, c_ptr, c_f32,
                             c_ptr, c_ptr]),
}

# kernels launched by one call of each entry point (used by bench.py's gpu_launches count)
_KERNELS_PER_CALL = {
    "ssb_dtw_align_batch": 2, "ssb_dtw_time_warp_batch": 2, "ssb_dtw_time_warp_batch_f64": 1, "ssb_dtw_align_ragged": 2, "ssb_dtw_cost_batch": 1,
    "ssb_dtw_loss_rows": 1, "ssb_prep_planes": 1, "ssb_mel_fwd": 1, "ssb_gemm_nn": 1,
    "ssb_gemm_nt": 1, "ssb_gemm_tn": 1, "ssb_colsum": 2, "ssb_colsum_planes": 2, "ssb_adamw_flat": 2, "ssb_ctc_loss_fused": 3, "ssb_bn_stats": 2, "ssb_bn_apply": 1,
    "ssb_bn_bwd": 3, "ssb_bn_bwd2": 3, "ssb_add_dropout_ln_fwd": 1, "ssb_add_dropout_ln_bwd": 2,
    "ssb_band_attn_fwd": 1, "ssb_band_attn_bwd": 2,
    "ssb_split_bf16": 1, "ssb_split_bf16_t": 1, "ssb_split_bf16_2d": 1, "ssb_gemm_tc_kmajor": 1, "ssb_gemm_tc_wgrad": 1, "ssb_gemm_tc_batched": 1,
    "ssb_gemm_tc_batched_tn": 1, "ssb_pad_split_heads": 1, "ssb_transpose_split_heads": 1,
    "ssb_attn_softmax_fwd": 1, "ssb_attn_ds_bwd": 1,
    "ssb_attn_fused_fwd": 1, "ssb_attn_delta": 1, "ssb_attn_fused_bwd": 1,
    "ssb_emg_filtfilt_chain": 1, "ssb_emg_subsample": 1, "ssb_voc_mix": 1, "ssb_voc_post": 1,
}
launch_count = 0   # running total of libssb kernel launches issued by this process

_lock = threading.Lock()
_lib = None


def _counted(fn, n):
    def call(*args):
        global launch_count
        launch_count += n
        return fn(*args)
    call.__name__ = fn.__name__
    return call


class SSBError(RuntimeError):
    """A libssb entry point returned a non-zero status."""

    def __init__(self, code, message):
        super().__init__(f"libssb error {code}: {message}")
        self.code = code


def signatures():
    return dict(_SIGNATURES)


def load():
    """Load (building first if stale/missing) libssb.so and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if os.environ.get("SSB_REBUILD") == "1" or _build.is_stale():
            _build.build(force=os.environ.get("SSB_REBUILD") == "1")  # needs nvcc; raises loudly
        lib = ctypes.CDLL(LIB_PATH)
        _check_abi(lib)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError => header/library mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
            if name in _KERNELS_PER_CALL:
                setattr(lib, name, _counted(fn, _KERNELS_PER_CALL[name]))
        _lib = lib
    return _lib


ABI_VERSION = 212      # include/ssb.h SSB_ABI_VERSION: bumped whenever a struct or signature changes


def _check_abi(lib):
    """A stale or foreign libssb.so must not load: same ABI version, and the ctypes mirrors of the
    descriptor structs have the size the C compiler gave them."""
    lib.ssb_version.restype = ctypes.c_int
    got = lib.ssb_version()
    if got != ABI_VERSION:
        raise SSBError(-4, f"libssb.so reports ABI {got}, this package binds ABI {ABI_VERSION}: "
                           f"rebuild (python -m silent_speech_b200.build --force)")
    lib.ssb_sizeof.restype = c_i64
    lib.ssb_sizeof.argtypes = [ctypes.c_int]
    for which, cls in enumerate((Gather, Scatter, Epilogue, TcOperand, DtwPair, Utt, PrepEntry,
                                 EmgRec)):
        if lib.ssb_sizeof(which) != ctypes.sizeof(cls):
            raise SSBError(-4, f"struct layout mismatch for {cls.__name__}: library "
                               f"{lib.ssb_sizeof(which)} B, ctypes {ctypes.sizeof(cls)} B")


def last_error():
    return load().ssb_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise SSBError(rc, last_error())


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise SSBError(-1, f"{name} must live on a CUDA device; libssb has no CPU path")
    return t

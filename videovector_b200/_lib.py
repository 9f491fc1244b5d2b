"""ctypes binding of libvv_b200.so (the C-ABI declared in include/vv_b200.h).

The product path has no fallback: if the shared library is missing this module
raises, and every wrapper raises VVError on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvv_b200.so")

VV_MAX_CONTEXT = 64
PREC = {"fp32_simt": 0, "tf32x3": 1, "tf32": 2, "bf16": 3, "f16x3": 4}
F16X3_HEADER_BYTES = 128
RECORD_VIDEO_SHOTS, RECORD_TEST_WINDOWS = 0, 1
CONTEXT = {"pairwise": 0, "window": 1, "past": 2, "past_continuous": 3, "past_continuous_fixed": 4}
DROPOUT_NONE, DROPOUT_MASK01, DROPOUT_MASK_U32, DROPOUT_PHILOX, DROPOUT_HASH = 0, 1, 2, 3, 4


class VVError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p)]


class Act(C.Structure):
    _fields_ = [("relu", C.c_int), ("negative_slope", C.c_float), ("dropout_mode", C.c_int),
                ("dropout_ratio", C.c_float), ("mask", C.c_void_p), ("mask_out", C.c_void_p),
                ("seed", C.c_uint64), ("step", C.c_uint64)]


class RankCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("C", C.c_int), ("Nn", C.c_int), ("N", C.c_int),
                ("coeff", C.c_float * VV_MAX_CONTEXT), ("margin", C.c_float), ("norm", C.c_int),
                ("eps", C.c_float)]


class TrainerCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("C", C.c_int), ("Nn", C.c_int), ("K", C.c_int), ("N", C.c_int),
                ("coeff", C.c_float * VV_MAX_CONTEXT), ("margin", C.c_float), ("norm", C.c_int),
                ("dropout_ratio", C.c_float), ("dropout_mode", C.c_int), ("dropout_seed", C.c_uint64),
                ("loss_weight", C.c_float), ("regularization", C.c_float),
                ("lr_policy", C.c_char * 16), ("base_lr", C.c_float), ("gamma", C.c_float),
                ("power", C.c_float), ("stepsize", C.c_int),
                ("momentum", C.c_float), ("weight_decay", C.c_float), ("reg_type", C.c_int),
                ("lr_mult", C.c_float * 2), ("decay_mult", C.c_float * 2),
                ("prec", C.c_int), ("world_size", C.c_int), ("rank", C.c_int),
                ("compute_dgrad", C.c_int), ("keep_blobs", C.c_int), ("split_rank_loss", C.c_int)]


_P = C.c_void_p
_i, _i64, _f, _u64, _sz = C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t

# name -> (restype, argtypes)
SIGNATURES = {
    "vv_last_error": (C.c_char_p, []),
    "vv_version": (_i, []),
    "vv_device_check": (_i, []),
    "vv_gather_rows": (_i, [_P, _i64, _i, _P, _P, _i, _i, _P, _P, _P, _i, _P, _P]),
    "vv_prepare_operand": (_i, [_P, _i64, _i, _P, _P, _P]),
    "vv_prepare_bank_operand": (_i, [_P, _i64, _i, _i, _P, _P, _P]),
    "vv_operand_bytes": (C.c_size_t, [_i64, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "vv_operand_set_scale": (_i, [_P, _i, _i, _P]),
    "vv_operand_rescale": (_i, [_P, _i, _i, _P]),
    "vv_operand_measure": (_i, [_P, _i, _P, _i64, _P]),
    "vv_ip_forward": (_i, [Operand, Operand, _P, _i, _i, _i, _i, C.POINTER(Act), _P, _P, _P]),
    "vv_ip_wgrad": (_i, [Operand, Operand, _i, _i, _i, _i, _f, _P, _i, _P, _sz, _P]),
    "vv_ip_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "vv_ip_wgrad_auto_nsplit": (_i, [_i, _i, _i, _i]),
    "vv_ip_bias_grad": (_i, [_P, _i, _i, _P, _P]),
    "vv_ip_dgrad": (_i, [Operand, Operand, _i, _i, _i, _i, _P, _P]),
    "vv_reduce_parts": (_i, [_P, _i, _i64, _i64, _P, _P]),
    "vv_gather_plan": (_i, [_P, _i, _P, _P, _i, _i, _P, _P, _P]),
    "vv_gather_plan_checked": (_i, [_P, _i64, _i, _P, _P, _i, _i, _P, _P, _P, _P]),
    "vv_ip_forward_gathered": (_i, [Operand, _i64, _P, _P, _P, Operand, _P, _i, _i, _i, _i, C.POINTER(Act), _P, _P, _P]),
    "vv_ip_wgrad_gathered": (_i, [Operand, Operand, _i64, _P, _i, _i, _i, _i, _f, _P, _i, _P]),
    "vv_ip_wgrad_gathered_part": (_i, [Operand, Operand, _i64, _P, _i, _i, _i, _i, _f, _P, _i, _i, _i, _P]),
    "vv_add_column": (_i, [_P, _i64, _i, _P, _i, _P]),
    "vv_rank_loss_backward_ex": (_i, [_P, C.POINTER(RankCfg), _P, _f, _i, _f, _P, _P, _P, _i, _P, _P, _P, _P]),
    "vv_trainer_set_bank": (_i, [_P, _P, _i64]),
    "vv_rank_loss_fused_supported": (_i, [C.POINTER(RankCfg)]),
    "vv_rank_loss_fused": (_i, [_P, C.POINTER(RankCfg), _f, _i, _f, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _i, _P, _P, _P, _P]),
    "vv_rank_loss_forward": (_i, [_P, C.POINTER(RankCfg), _P, _P, _P, _P, _P, _P, _P, _P]),
    "vv_rank_loss_backward": (_i, [_P, C.POINTER(RankCfg), _P, _f, _i, _f, _P, _P, _P, _i, _P, _P]),
    "vv_sgd_update": (_i, [_P, _P, _i, _i64, _P, _P, _i64, _f, _f, _f, _i, _f, _P, _P, _i, _P]),
    "vv_learning_rate": (_f, [C.c_char_p, _f, _f, _f, _i, _i]),
    "vv_relu_forward": (_i, [_P, _i64, _f, _P, _P]),
    "vv_relu_backward": (_i, [_P, _P, _i64, _f, _P, _P]),
    "vv_dropout_forward": (_i, [_P, _P, _i, _i64, _f, _P, _P]),
    "vv_dropout_backward": (_i, [_P, _P, _i, _i64, _f, _P, _P]),
    "vv_dropout_make_mask": (_i, [_P, _i, _i, _f, _u64, _u64, _P]),
    "vv_dropout_make_mask_mode": (_i, [_P, _i, _i, _f, _u64, _u64, _i, _P]),
    "vv_eltwise_sum_forward": (_i, [C.POINTER(_P), C.POINTER(_f), _i, _i64, _P, _P]),
    "vv_eltwise_prod_forward": (_i, [_P, _P, _i64, _P, _P]),
    "vv_axpby": (_i, [_i64, _f, _P, _f, _P, _P]),
    "vv_mul": (_i, [_i64, _P, _P, _P, _P]),
    "vv_sign_axpy": (_i, [_i64, _f, _P, _P, _P]),
    "vv_l2norm_forward": (_i, [_P, _i, _i, _P, _P]),
    "vv_l2norm_backward": (_i, [_P, _P, _i, _i, _P, _P]),
    "vv_rowsum_forward": (_i, [_P, _i, _i, _i, _P, _P]),
    "vv_rowsum_backward": (_i, [_P, _i, _i, _i, _P, _P]),
    "vv_copy_strided": (_i, [_P, _i64, _P, _i64, _i64, _i64, _P]),
    "vv_max_margin_forward": (_i, [_P, _P, _i, _f, _i, _P, _P, _P, _P]),
    "vv_max_margin_backward": (_i, [_P, _P, _i, _f, _i, _f, _P, _P, _P]),
    "vv_max_margin_forward_w": (_i, [_P, _P, _P, _i, _f, _i, _P, _P, _P, _P]),
    "vv_max_margin_backward_w": (_i, [_P, _P, _P, _i, _f, _i, _f, _P, _P, _P]),
    "vv_id_to_weight": (_i, [_P, _i, _P, _P, _i, _P, _P]),
    "vv_fill_bank": (_i, [_P, _i64, _i, _u64, _P]),
    "vv_bank_value_host": (_f, [_u64, _i64, _i, _i]),
    "vv_id_lookup_forward": (_i, [_P, _i, _i, _P, _i, _P, _P]),
    "vv_id_lookup_backward": (_i, [_P, _P, _i, _i, _i, _P, _P]),
    "vv_gather_mean_rows": (_i, [_P, _i64, _i, _P, _i, _i, _P, _P, _P]),
    "vv_retrieval_stats_workspace_bytes": (C.c_size_t, [_i]),
    "vv_retrieval_stats": (_i, [_P, _i, _i, _P, _P, _i, _P, _P, C.c_size_t, _P, _P, _P]),
    "vv_retrieval_stats_ex": (_i, [_P, _i, _i, _P, _P, _i, _P, _P, C.c_size_t, _P, _P, _P, _P]),
    "vv_video_mean_rows": (_i, [_P, _i, _i, _P, _i, _P, _P]),
    "vv_record_set_create": (_P, [_i, _i, _i]),
    "vv_record_set_destroy": (None, [_P]),
    "vv_record_set_add": (_i, [_P, C.c_char_p, C.c_size_t]),
    "vv_record_set_load_file": (_i, [_P, C.c_char_p]),
    "vv_record_set_info": (_i, [_P, _P, _P, _P, _P]),
    "vv_record_set_tables": (_i, [_P, _P, _P, _P]),
    "vv_record_set_bank": (_P, [_P]),
    "vv_record_set_upload": (_i, [_P, _P, _P]),
    "vv_sampler_create": (_P, [_i, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, C.c_uint]),
    "vv_sampler_create_ex": (_P, [_i, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, C.c_uint, _i]),
    "vv_sampler_create_ex2": (_P, [_i, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, C.c_uint, _i, _i, _i, _P, _P, _P, _i]),
    "vv_sampler_destroy": (None, [_P]),
    "vv_sampler_next": (_i, [_P, _P, _P]),
    "vv_sampler_cursor": (_i, [_P]),
    "vv_sampler_prefetch": (_i, [_P, _i]),
    "vv_sampler_prefetch_ready": (_i, [_P]),
    "vv_sampler_set_row_base": (_i, [_P, C.c_int32]),
    "vv_glibc_rand_create": (_P, [C.c_uint]),
    "vv_glibc_rand_next": (_i, [_P]),
    "vv_glibc_rand_destroy": (None, [_P]),
    "vv_trainer_create": (_P, [C.POINTER(TrainerCfg), _P]),
    "vv_trainer_destroy": (None, [_P]),
    "vv_trainer_weight": (_P, [_P]),
    "vv_trainer_bias": (_P, [_P]),
    "vv_trainer_weight_hist": (_P, [_P]),
    "vv_trainer_bias_hist": (_P, [_P]),
    "vv_trainer_weight_diff": (_P, [_P]),
    "vv_trainer_bias_diff": (_P, [_P]),
    "vv_trainer_blob": (_P, [_P, C.c_char_p]),
    "vv_trainer_sync_weights": (_i, [_P]),
    "vv_trainer_step": (_i, [_P, _P, _i64, _P, _P, _P, _i, _i]),
    "vv_trainer_last_launches": (_i, [_P]),
    "vv_trainer_set_timing": (_i, [_P, _i]),
    "vv_trainer_phase_ms": (_i, [_P, _P, _P]),
    "vv_trainer_extract": (_i, [_P, _P, _i64, _P]),
    "vv_dp_unique_id": (_i, [_P]),
    "vv_dp_init": (_i, [_P, _P]),
    "vv_dp_allreduce_inplace": (_i, [_P, _P, _i64, _P]),
    "vv_dp_mode": (_i, [_P]),
    "vv_dp_mode_reason": (C.c_char_p, [_P]),
    "vv_dp_gather_state": (_i, [_P]),
}

_lib = None


def load():
    """Load libvv_b200.so (raises if it was not built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VVError("libvv_b200.so not found at %s -- run `make` (or __graft_entry__.build())" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VVError("vv_b200 error %d: %s" % (rc, load().vv_last_error().decode()))


def last_error():
    return load().vv_last_error().decode()

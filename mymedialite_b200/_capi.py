"""ctypes binding of libmmlb200.so -- the same entry points the C# classes bind through P/Invoke
(include/mmlb200.h). There is no fallback: a missing library or a failing call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libmmlb200.so")

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
vp = C.c_void_p

LOSS_RMSE, LOSS_MAE, LOSS_LOGISTIC = 0, 1, 2
SCHEDULE_SERIAL, SCHEDULE_DSGD, SCHEDULE_NAIVE = 0, 1, 2
GROUPS_PERM_MOD, GROUPS_BALANCED = 0, 1
INTRA_ROUNDS, INTRA_ASYNC = 0, 1
FILE_RATINGS, FILE_RATINGS_NO_VALUE, FILE_FEEDBACK = 0, 1, 2
MAP_IDENTITY, MAP_FIRST_SEEN = 0, 1
ERR_CUDA, ERR_ARG, ERR_STATE, ERR_NCCL, ERR_UNSUPPORTED, ERR_FORMAT, ERR_IO = 1, 2, 3, 4, 5, 6, 7


class MmlError(RuntimeError):
    """Non-zero status from libmmlb200 (the C# wrapper throws InvalidOperationException here)."""

    def __init__(self, status, message):
        super().__init__("libmmlb200 status %d: %s" % (status, message))
        self.status = status


class MFParams(C.Structure):
    _fields_ = [
        ("biased", C.c_int32), ("num_factors", C.c_int32), ("learn_rate", C.c_float), ("decay", C.c_float),
        ("regularization", C.c_float), ("bias_learn_rate", C.c_float), ("bias_reg", C.c_float),
        ("reg_u", C.c_float), ("reg_i", C.c_float), ("frequency_regularization", C.c_int32),
        ("loss", C.c_int32), ("bold_driver", C.c_int32), ("max_threads", C.c_int32),
        ("schedule", C.c_int32), ("num_groups", C.c_int32), ("num_subgroups", C.c_int32),
        ("group_rule", C.c_int32), ("persistent", C.c_int32), ("hot_item_factor", C.c_float), ("hot_copies", C.c_int32), ("intra_block", C.c_int32), ("hot_merge_average", C.c_int32), ("async_workers", C.c_int32), ("ctas_per_group", C.c_int32),
    ]


class WrmfParams(C.Structure):
    _fields_ = [("num_factors", C.c_int32), ("alpha", C.c_double), ("regularization", C.c_double)]


def _opt(ptr_type):
    """ndpointer that also accepts None (NULL)."""
    base = ptr_type

    class _Opt(base):
        @classmethod
        def from_param(cls, obj):
            if obj is None:
                return None
            return base.from_param(obj)
    return _Opt


oi32p, oi64p, of32p = _opt(i32p), _opt(i64p), _opt(f32p)
PP = C.POINTER(vp)

SIGNATURES = {
    "mml_last_error": (C.c_char_p, []),
    "mml_version": (C.c_char_p, []),
    "mml_ctx_create": (C.c_int32, [C.c_int32, oi32p, PP]),
    "mml_dist_unique_id": (C.c_int32, [np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")]),
    "mml_ctx_create_dist": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), PP]),
    "mml_ctx_destroy": (C.c_int32, [vp]),
    "mml_ctx_synchronize": (C.c_int32, [vp]),
    "mml_ctx_flush_l2": (C.c_int32, [vp]),
    "mml_ctx_probe_l2": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "mml_ctx_sm_count": (C.c_int32, [vp, C.POINTER(C.c_int32)]),
    "mml_ingest_file": (C.c_int32, [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, PP]),
    "mml_ingest_text": (C.c_int32, [C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, PP]),
    "mml_ingest_destroy": (C.c_int32, [vp]),
    "mml_ingest_info": (C.c_int32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mml_ingest_copy": (C.c_int32, [vp, oi32p, oi32p, of32p]),
    "mml_ingest_original_ids": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int64, oi64p,
                                            C.POINTER(C.c_int64)]),
    "mml_ingest_to_ratings": (C.c_int32, [vp, vp, PP]),
    "mml_ingest_to_feedback": (C.c_int32, [vp, vp, PP]),
    "mml_ratings_create": (C.c_int32, [vp, oi32p, oi32p, of32p, C.c_int64, C.c_int32, C.c_int32, PP]),
    "mml_ratings_destroy": (C.c_int32, [vp]),
    "mml_ratings_counts": (C.c_int32, [vp, C.c_int32, i32p]),
    "mml_ratings_csr": (C.c_int32, [vp, C.c_int32, i64p, i32p]),
    "mml_ratings_stats": (C.c_int32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mml_shuffle_apply": (C.c_int32, [vp, i32p, i32p, C.c_int64]),
    "mml_partition_blocks": (C.c_int32, [vp, i32p, i32p, C.c_int32, i64p, i32p]),
    "mml_partition_indices": (C.c_int32, [vp, i32p, C.c_int64, C.c_int32, i64p, i32p]),
    "mml_mf_params_default": (None, [C.POINTER(MFParams)]),
    "mml_sgd_create": (C.c_int32, [vp, vp, C.POINTER(MFParams), oi32p, oi32p, PP]),
    "mml_sgd_destroy": (C.c_int32, [vp]),
    "mml_sgd_set_model": (C.c_int32, [vp, f32p, f32p, of32p, of32p]),
    "mml_sgd_init_model": (C.c_int32, [vp, C.c_uint64, C.c_double, C.c_double]),
    "mml_sgd_get_model": (C.c_int32, [vp, of32p, of32p, of32p, of32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mml_sgd_set_learnrate": (C.c_int32, [vp, C.c_float]),
    "mml_sgd_set_scale": (C.c_int32, [vp, C.c_float, C.c_float, C.c_float]),
    "mml_sgd_iterate": (C.c_int32, [vp, oi32p, oi32p, C.c_int64]),
    "mml_sgd_invalidate_index": (C.c_int32, [vp]),
    "mml_sgd_iterate_indices": (C.c_int32, [vp, oi32p, C.c_int64, C.c_int32, C.c_int32]),
    "mml_sgd_learn_factors": (C.c_int32, [vp, oi32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "mml_sgd_predict": (C.c_int32, [vp, oi32p, oi32p, C.c_int64, of32p]),
    "mml_sgd_fold_in": (C.c_int32, [vp, oi64p, oi32p, of32p, C.c_int64, of32p, C.c_int32, of32p]),
    "mml_sgd_score_items": (C.c_int32, [vp, of32p, C.c_int64, oi32p, C.c_int64, of32p]),
    "mml_sgd_set_rows": (C.c_int32, [vp, C.c_int32, oi32p, C.c_int64, of32p, of32p]),
    "mml_sgd_evaluate": (C.c_int32, [vp, oi32p, oi32p, of32p, C.c_int64, f32p]),
    "mml_sgd_evaluate_train": (C.c_int32, [vp, f32p]),
    "mml_sgd_objective": (C.c_int32, [vp, C.POINTER(C.c_double)]),
    "mml_sgd_stats": (C.c_int32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "mml_sgd_strata_info": (C.c_int32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mml_sgd_hot_items": (C.c_int32, [vp, C.POINTER(C.c_int64)]),
    "mml_sgd_grid": (C.c_int32, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mml_topn_mf": (C.c_int32, [vp, f32p, C.c_int32, f32p, C.c_int32, C.c_int32, oi32p, C.c_int64, C.c_int32,
                                oi32p, C.c_int64, oi64p, oi32p, i32p, f32p, i32p]),
    "mml_items_evaluate_mf": (C.c_int32, [vp, f32p, C.c_int32, f32p, C.c_int32, C.c_int32, oi32p, C.c_int64, oi32p, C.c_int64,
                                          oi64p, oi32p, oi64p, oi32p, C.c_int32, of32p, oi32p]),
    "mml_wrmf_evaluate": (C.c_int32, [vp, oi32p, C.c_int64, oi32p, C.c_int64, oi64p, oi32p, oi64p, oi32p, C.c_int32,
                                      of32p, oi32p]),
    "mml_topn_set_mode": (C.c_int32, [C.c_int32]),
    "mml_topn_set_filter": (C.c_int32, [C.c_int32]),
    "mml_topn_last_stats": (C.c_int32, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "mml_feedback_create": (C.c_int32, [vp, oi32p, oi32p, C.c_int64, C.c_int32, C.c_int32, PP]),
    "mml_feedback_destroy": (C.c_int32, [vp]),
    "mml_feedback_nnz": (C.c_int32, [vp, C.POINTER(C.c_int64)]),
    "mml_feedback_csr": (C.c_int32, [vp, C.c_int32, i64p, i32p]),
    "mml_wrmf_create": (C.c_int32, [vp, vp, C.POINTER(WrmfParams), PP]),
    "mml_wrmf_destroy": (C.c_int32, [vp]),
    "mml_wrmf_set_model": (C.c_int32, [vp, f32p, f32p]),
    "mml_wrmf_init_model": (C.c_int32, [vp, C.c_uint64, C.c_double, C.c_double]),
    "mml_wrmf_get_model": (C.c_int32, [vp, of32p, of32p]),
    "mml_wrmf_iterate": (C.c_int32, [vp]),
    "mml_wrmf_retrain": (C.c_int32, [vp, C.c_int32, oi32p, C.c_int64]),
    "mml_wrmf_stats": (C.c_int32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "mml_wrmf_shard": (C.c_int32, [vp, C.c_int32, C.POINTER(C.c_int32)]),
    "mml_wrmf_set_mode": (C.c_int32, [C.c_int32]),
    "mml_wrmf_debug_gram": (C.c_int32, [vp, f32p, C.POINTER(C.c_int32)]),
    "mml_wrmf_recommend": (C.c_int32, [vp, oi32p, C.c_int64, C.c_int32, oi32p, C.c_int64, oi64p, oi32p, i32p, f32p, i32p]),
    "mml_sgd_schedule_dump": (C.c_int32, [vp, oi32p, i32p, oi32p, oi32p, oi32p]),
}

TOPN_AUTO, TOPN_EXACT, TOPN_TENSOR = 0, 1, 2
TOPN_FILTER_BF16, TOPN_FILTER_TF32 = 0, 1
WRMF_AUTO, WRMF_FP64, WRMF_TENSOR, WRMF_TENSOR_F64, WRMF_TENSOR_PCG = 0, 1, 2, 3, 4

_lib = None


def load():
    """Loads libmmlb200.so; raises if it has not been built (python -m mymedialite_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("libmmlb200.so is missing at %s: build it with `python -m mymedialite_b200.build` "
                          "(there is no CPU fallback)" % SO_PATH)
    # libmmlb200 needs libnccl.so.2. PyTorch bundles a newer NCCL under the same soname than the system one; whichever
    # is loaded first serves both, and torch does not start on the older system build -- so let torch load its own first.
    try:
        import torch  # noqa: F401
    except ImportError:
        pass
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status):
    if status != 0:
        raise MmlError(status, load().mml_last_error().decode("utf-8", "replace"))

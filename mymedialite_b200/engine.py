"""Thin object layer over the C ABI (handle lifetimes, numpy marshalling). The recommender classes in
recommenders.py -- the mirror of the reference's IRecommender surface -- are built on these."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MFParams, check


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Context:
    """One GPU of a one-process-per-GPU job (rank / world / unique_id), or -- n_gpus > 1 -- the GPUs device ..
    device + n_gpus - 1 driven from this one process (mml_ctx_create with n_gpus > 1: the NumGpus property of the host
    classes). Raises MmlError without a CUDA device: there is no CPU path."""

    def __init__(self, device=0, rank=0, world=1, unique_id=None, n_gpus=1):
        self.lib = _capi.load()
        self.rank, self.world, self.n_gpus = rank, world, int(n_gpus)
        h = C.c_void_p()
        if world > 1:
            assert n_gpus == 1, "one process per GPU and several GPUs per process do not mix"
            uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
            assert uid.size == 128
            check(self.lib.mml_ctx_create_dist(rank, world, device, uid, C.byref(h)))
        else:
            dev = np.arange(device, device + self.n_gpus, dtype=np.int32)
            check(self.lib.mml_ctx_create(self.n_gpus, dev, C.byref(h)))
        self.h = h

    @staticmethod
    def unique_id():
        """ncclUniqueId bytes: rank 0 creates them and hands them to the other ranks."""
        out = np.zeros(128, np.uint8)
        check(_capi.load().mml_dist_unique_id(out))
        return out

    def sm_count(self):
        v = C.c_int32()
        check(self.lib.mml_ctx_sm_count(self.h, C.byref(v)))
        return v.value

    def flush_l2(self):
        check(self.lib.mml_ctx_flush_l2(self.h))

    def probe_l2(self, mode, n_rows, row_floats, reps=5):
        """Rows per second the L2 serves under the SGD kernel's item-row access pattern (0 read, 1 red.add, 2 read + red.add)."""
        out = C.c_double(0.0)
        check(self.lib.mml_ctx_probe_l2(self.h, mode, n_rows, row_floats, reps, C.byref(out)))
        return out.value

    def synchronize(self):
        check(self.lib.mml_ctx_synchronize(self.h))

    def close(self):
        if self.h:
            self.lib.mml_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceRatings:
    """COO rating set resident in HBM (StaticRatings / DataSet of the reference)."""

    def __init__(self, ctx, users, items, values, max_user=None, max_item=None):
        self.ctx = ctx
        self.lib = ctx.lib
        users, items, values = _i32(users), _i32(items), _f32(values)
        self.n = int(users.shape[0])
        self.max_user = int(users.max()) if max_user is None and self.n else (-1 if max_user is None else int(max_user))
        self.max_item = int(items.max()) if max_item is None and self.n else (-1 if max_item is None else int(max_item))
        h = C.c_void_p()
        check(self.lib.mml_ratings_create(ctx.h, users, items, values, self.n, self.max_user, self.max_item, C.byref(h)))
        self.h = h

    def counts(self, by_item=False):
        out = np.zeros((self.max_item if by_item else self.max_user) + 1, dtype=np.int32)
        check(self.lib.mml_ratings_counts(self.h, 1 if by_item else 0, out))
        return out

    def csr(self, by_item=False):
        rows = (self.max_item if by_item else self.max_user) + 1
        ptr = np.zeros(rows + 1, dtype=np.int64)
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        check(self.lib.mml_ratings_csr(self.h, 1 if by_item else 0, ptr, idx))
        return ptr, idx[:self.n]

    def stats(self):
        a, mn, mx = C.c_float(), C.c_float(), C.c_float()
        check(self.lib.mml_ratings_stats(self.h, C.byref(a), C.byref(mn), C.byref(mx)))
        return a.value, mn.value, mx.value

    def partition_blocks(self, user_perm, item_perm, g):
        ptr = np.zeros(g * g + 1, dtype=np.int64)
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        check(self.lib.mml_partition_blocks(self.h, _i32(user_perm), _i32(item_perm), int(g), ptr, idx))
        return ptr, idx[:self.n]

    def close(self):
        if self.h:
            if self.ctx.h:                      # the context owns the device state: nothing to release once it is gone
                self.lib.mml_ratings_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition_indices(ctx, random_index, num_groups):
    """MultiCore.PartitionIndices (MultiCore.cs:79-92) on the device: (list_ptr int64[num_groups + 1], idx int32[n])."""
    ri = _i32(random_index)
    ptr = np.zeros(int(num_groups) + 1, np.int64)
    idx = np.zeros(max(ri.shape[0], 1), np.int32)
    check(ctx.lib.mml_partition_indices(ctx.h, ri, ri.shape[0], int(num_groups), ptr, idx))
    return ptr, idx[:ri.shape[0]]


def default_params(**kw):
    p = MFParams()
    _capi.load().mml_mf_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError("mml_mf_params has no field %r" % k)
        setattr(p, k, v)
    return p


class SgdModel:
    """MatrixFactorization / BiasedMatrixFactorization state on the device."""

    def __init__(self, ctx, ratings, params, user_perm=None, item_perm=None):
        self.ctx, self.ratings, self.lib = ctx, ratings, ctx.lib
        self.params = params
        self.k = params.num_factors
        h = C.c_void_p()
        check(self.lib.mml_sgd_create(ctx.h, ratings.h, C.byref(params), _i32(user_perm), _i32(item_perm), C.byref(h)))
        self.h = h
        self.n_users, self.n_items = ratings.max_user + 1, ratings.max_item + 1

    def set_model(self, U, V, bu=None, bi=None):
        U, V = _f32(U), _f32(V)
        assert U.shape == (self.n_users, self.k) and V.shape == (self.n_items, self.k)
        check(self.lib.mml_sgd_set_model(self.h, U, V, _f32(bu), _f32(bi)))

    def init_model(self, seed, mean=0.0, stddev=0.1):
        check(self.lib.mml_sgd_init_model(self.h, int(seed), float(mean), float(stddev)))

    def get_model(self, factors=True, biases=True):
        U = np.zeros((self.n_users, self.k), np.float32) if factors else None
        V = np.zeros((self.n_items, self.k), np.float32) if factors else None
        bu = np.zeros(self.n_users, np.float32) if biases else None
        bi = np.zeros(self.n_items, np.float32) if biases else None
        gb, lr = C.c_float(), C.c_float()
        check(self.lib.mml_sgd_get_model(self.h, U, V, bu, bi, C.byref(gb), C.byref(lr)))
        return dict(U=U, V=V, bu=bu, bi=bi, global_bias=gb.value, learnrate=lr.value)

    @property
    def learnrate(self):
        return self.get_model(False, False)["learnrate"]

    def set_learnrate(self, lr):
        check(self.lib.mml_sgd_set_learnrate(self.h, float(lr)))

    def iterate(self, subepoch_sequence=None, random_index=None):
        ri = _i32(random_index)
        check(self.lib.mml_sgd_iterate(self.h, _i32(subepoch_sequence), ri, 0 if ri is None else ri.shape[0]))

    def iterate_indices(self, indices, update_user=True, update_item=True):
        idx = _i32(indices)
        check(self.lib.mml_sgd_iterate_indices(self.h, idx, idx.shape[0], int(update_user), int(update_item)))

    def learn_factors(self, indices, num_iter, update_user=True, update_item=True):
        """LearnFactors (MatrixFactorization.cs:198-202): num_iter passes of Iterate(indices, update_user, update_item)."""
        idx = _i32(indices)
        check(self.lib.mml_sgd_learn_factors(self.h, idx, idx.shape[0], int(update_user), int(update_item), int(num_iter)))

    def fold_in(self, rated_items, rated_values, init_factors, num_iter):
        """Batch FoldIn: rated_items / rated_values are per-user sequences (already shuffled), init_factors [n, k]."""
        n = len(rated_items)
        ptr = np.zeros(n + 1, np.int64)
        ptr[1:] = np.cumsum([len(x) for x in rated_items])
        it = _i32(np.concatenate([np.asarray(x, np.int32) for x in rated_items]) if ptr[-1] else np.zeros(1, np.int32))
        va = _f32(np.concatenate([np.asarray(x, np.float32) for x in rated_values]) if ptr[-1] else np.zeros(1, np.float32))
        init = _f32(np.asarray(init_factors, np.float32).reshape(max(n, 1), -1))
        stride = self.k + 1 if self.params.biased else self.k
        out = np.zeros((max(n, 1), stride), np.float32)
        check(self.lib.mml_sgd_fold_in(self.h, ptr, it, va, n, init, int(num_iter), out))
        return out[:n]

    def score_items(self, user_vectors, candidates):
        v = _f32(np.atleast_2d(np.asarray(user_vectors, np.float32)))
        cand = _i32(candidates)
        out = np.zeros((max(v.shape[0], 1), max(cand.shape[0], 1)), np.float32)
        check(self.lib.mml_sgd_score_items(self.h, v, v.shape[0], cand, cand.shape[0], out))
        return out[:v.shape[0], :cand.shape[0]]

    def set_rows(self, ids, factors=None, biases=None, by_item=False):
        ids = _i32(np.atleast_1d(ids))
        f = None if factors is None else _f32(np.asarray(factors, np.float32).reshape(ids.shape[0], -1))
        b = None if biases is None else _f32(np.atleast_1d(biases))
        check(self.lib.mml_sgd_set_rows(self.h, 1 if by_item else 0, ids, ids.shape[0], f, b))

    def predict(self, users, items):
        users, items = _i32(users), _i32(items)
        out = np.zeros(users.shape[0], np.float32)
        check(self.lib.mml_sgd_predict(self.h, users, items, users.shape[0], out))
        return out

    def evaluate(self, users, items, values):
        users, items, values = _i32(users), _i32(items), _f32(values)
        out = np.zeros(4, np.float32)
        check(self.lib.mml_sgd_evaluate(self.h, users, items, values, users.shape[0], out))
        return dict(RMSE=float(out[0]), MAE=float(out[1]), NMAE=float(out[2]), CBD=float(out[3]))

    def evaluate_train(self):
        out = np.zeros(4, np.float32)
        check(self.lib.mml_sgd_evaluate_train(self.h, out))
        return dict(RMSE=float(out[0]), MAE=float(out[1]), NMAE=float(out[2]), CBD=float(out[3]))

    def objective(self):
        v = C.c_double()
        check(self.lib.mml_sgd_objective(self.h, C.byref(v)))
        return v.value

    def stats(self):
        n, ms = C.c_int64(), C.c_float()
        check(self.lib.mml_sgd_stats(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def strata_info(self):
        G, W, ns, sb = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        check(self.lib.mml_sgd_strata_info(self.h, C.byref(G), C.byref(W), C.byref(ns), C.byref(sb)))
        cpg = C.c_int32()
        check(self.lib.mml_sgd_grid(self.h, C.byref(G), C.byref(cpg)))
        return dict(G=G.value, W=W.value, n_rounds=ns.value, staged_bytes=sb.value, cpg=cpg.value)

    def schedule(self, subepoch_sequence=None, detail=False, rounds=False):
        n = self.ratings.n
        order = np.zeros(max(n, 1), np.int32)
        block = np.zeros(max(n, 1), np.int32) if detail else None
        copy = np.zeros(max(n, 1), np.int32) if detail else None
        rnd = np.zeros(max(n, 1), np.int32) if rounds else None
        check(self.lib.mml_sgd_schedule_dump(self.h, _i32(subepoch_sequence), order, block, copy, rnd))
        if rounds:
            return rnd[:n]
        return (order[:n], block[:n], copy[:n]) if detail else order[:n]

    def round_sizes(self, subepoch_sequence=None):
        """Sizes of the rounds in schedule order."""
        rnd = self.schedule(subepoch_sequence, rounds=True)
        if rnd.size == 0:
            return np.zeros(0, np.int64)
        cuts = np.flatnonzero(np.diff(rnd)) + 1
        return np.diff(np.concatenate([[0], cuts, [rnd.size]]))

    def hot_items(self):
        v = C.c_int64()
        check(self.lib.mml_sgd_hot_items(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            if self.ctx.h:
                self.lib.mml_sgd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ignore_csr(ignore_lists, n_users):
    """list of per-user id arrays -> (ptr int64, idx int32) or (None, None)."""
    if ignore_lists is None:
        return None, None
    ptr = np.zeros(n_users + 1, np.int64)
    for b, lst in enumerate(ignore_lists):
        ptr[b + 1] = ptr[b] + len(lst)
    idx = np.concatenate([np.asarray(l, np.int32) for l in ignore_lists]) if ptr[-1] > 0 else np.zeros(1, np.int32)
    return ptr, np.ascontiguousarray(idx, np.int32)


def _topn_outputs(n_users, n, n_cand):
    n_out = n_cand if n < 0 else min(n, n_cand)
    return (np.zeros((n_users, max(n_out, 1)), np.int32), np.zeros((n_users, max(n_out, 1)), np.float32),
            np.zeros(max(n_users, 1), np.int32), n_out)


def topn_set_mode(mode):
    """_capi.TOPN_AUTO (default) / TOPN_EXACT (CUDA cores only) / TOPN_TENSOR (tcgen05 path or error)."""
    check(_capi.load().mml_topn_set_mode(int(mode)))


def topn_set_filter(kind):
    """_capi.TOPN_FILTER_TF32 (default) / TOPN_FILTER_BF16: operand precision of the tensor-core filter GEMM."""
    check(_capi.load().mml_topn_set_filter(int(kind)))


def topn_last_stats():
    a, b, ms = C.c_int64(), C.c_int64(), C.c_float()
    check(_capi.load().mml_topn_last_stats(C.byref(a), C.byref(b), C.byref(ms)))
    return dict(users_tensor_path=a.value, users_exact_path=b.value, tensor_path_ms=ms.value)


def topn_mf(ctx, U, V, users, n=-1, candidates=None, ignore_lists=None):
    """Recommender.Recommend for a batch of users on item-MF factors. Returns a list of (items, scores) per user."""
    U, V = _f32(U), _f32(V)
    users = _i32(users)
    cand = _i32(candidates)
    n_cand = V.shape[0] if cand is None else cand.shape[0]
    oi, os_, oc, n_out = _topn_outputs(users.shape[0], n, n_cand)
    ptr, idx = _ignore_csr(ignore_lists, users.shape[0])
    check(ctx.lib.mml_topn_mf(ctx.h, U, U.shape[0], V, V.shape[0], U.shape[1], users, users.shape[0], int(n),
                              cand, n_cand, ptr, idx, oi, os_, oc))
    return [(oi[b, :oc[b]].copy(), os_[b, :oc[b]].copy()) for b in range(users.shape[0])]


ITEM_MEASURES = ["AUC", "MAP", "NDCG", "MRR", "prec@5", "prec@10", "recall@5", "recall@10"]   # Eval/Items.cs:37-49


def _rows_csr(rows, n_users):
    """Rows given either as a list of per-user id sequences or as a ready (ptr, idx) pair."""
    if rows is None:
        return None, None
    if isinstance(rows, tuple) and len(rows) == 2 and isinstance(rows[0], np.ndarray):
        ptr, idx = np.ascontiguousarray(rows[0], np.int64), _i32(rows[1])
        return ptr, (idx if idx.size else np.zeros(1, np.int32))
    return _ignore_csr(rows, n_users)


def items_evaluate_mf(ctx, U, V, test_users, candidates, test_rows, ignore_rows=None, n=-1):
    """Per-user ranking measures (mml_items_evaluate_mf). Returns (measures [n_users, 8] float32, used [n_users] int32)."""
    U, V = _f32(U), _f32(V)
    users, cand = _i32(test_users), _i32(candidates)
    tp, ti = _rows_csr(test_rows, users.shape[0])
    ip, ii = _rows_csr(ignore_rows, users.shape[0])
    out = np.zeros((max(users.shape[0], 1), 8), np.float32)
    used = np.zeros(max(users.shape[0], 1), np.int32)
    check(ctx.lib.mml_items_evaluate_mf(ctx.h, U, U.shape[0], V, V.shape[0], U.shape[1], users, users.shape[0], cand,
                                        cand.shape[0], tp, ti, ip, ii, int(n), out, used))
    return out[:users.shape[0]], used[:users.shape[0]]


class DeviceFeedback:
    """PosOnlyFeedback resident in HBM: user and item matrices as CSR, duplicate events collapsed."""

    def __init__(self, ctx, users, items, max_user=None, max_item=None):
        self.ctx, self.lib = ctx, ctx.lib
        users, items = _i32(users), _i32(items)
        n = int(users.shape[0])
        self.max_user = (int(users.max()) if n else -1) if max_user is None else int(max_user)
        self.max_item = (int(items.max()) if n else -1) if max_item is None else int(max_item)
        h = C.c_void_p()
        check(self.lib.mml_feedback_create(ctx.h, users, items, n, self.max_user, self.max_item, C.byref(h)))
        self.h = h

    @property
    def nnz(self):
        v = C.c_int64()
        check(self.lib.mml_feedback_nnz(self.h, C.byref(v)))
        return v.value

    def csr(self, by_item=False):
        rows = (self.max_item if by_item else self.max_user) + 1
        ptr = np.zeros(rows + 1, np.int64)
        cols = np.zeros(max(self.nnz, 1), np.int32)
        check(self.lib.mml_feedback_csr(self.h, int(by_item), ptr, cols))
        return ptr, cols[:self.nnz]

    def close(self):
        if self.h:
            if self.ctx.h:
                self.lib.mml_feedback_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def wrmf_set_mode(mode):
    """_capi.WRMF_AUTO (default) / WRMF_FP64 (all-double CUDA-core kernels) / WRMF_TENSOR (tcgen05 Gram sums or error) /
    WRMF_TENSOR_F64 (the same with the double-precision Cholesky factor instead of the single-precision preconditioner)."""
    check(_capi.load().mml_wrmf_set_mode(int(mode)))


class WrmfModel:
    def __init__(self, ctx, feedback, num_factors=10, alpha=1.0, regularization=0.015):
        self.ctx, self.fb, self.lib = ctx, feedback, ctx.lib
        self.k = int(num_factors)
        p = _capi.WrmfParams(self.k, float(alpha), float(regularization))
        h = C.c_void_p()
        check(self.lib.mml_wrmf_create(ctx.h, feedback.h, C.byref(p), C.byref(h)))
        self.h = h
        self.n_users, self.n_items = feedback.max_user + 1, feedback.max_item + 1

    def set_model(self, U, V):
        U, V = _f32(U), _f32(V)
        assert U.shape == (self.n_users, self.k) and V.shape == (self.n_items, self.k)
        check(self.lib.mml_wrmf_set_model(self.h, U, V))

    def init_model(self, seed, mean=0.0, stddev=0.1):
        check(self.lib.mml_wrmf_init_model(self.h, int(seed), float(mean), float(stddev)))

    def get_model(self):
        U = np.zeros((self.n_users, self.k), np.float32)
        V = np.zeros((self.n_items, self.k), np.float32)
        check(self.lib.mml_wrmf_get_model(self.h, U, V))
        return U, V

    def iterate(self):
        check(self.lib.mml_wrmf_iterate(self.h))

    def retrain(self, ids, by_item=False):
        """RetrainUser / RetrainItem for a batch of rows (mml_wrmf_retrain)."""
        ids = _i32(np.atleast_1d(ids))
        check(self.lib.mml_wrmf_retrain(self.h, 1 if by_item else 0, ids, ids.shape[0]))

    def shard(self, by_item=False):
        """Multi-GPU: row ranges [ranges[r], ranges[r + 1]) each rank solves in the user (item) half-sweep."""
        r = (C.c_int32 * (self.ctx.world + 1))()
        check(self.lib.mml_wrmf_shard(self.h, int(by_item), r))
        return np.array(list(r), np.int32)

    def debug_gram(self):
        """(user, G) with G = sum of h_i h_i^T over that user's items as the tensor-core kernel computed it."""
        G = np.zeros((128, 128), np.float32)
        u = C.c_int32()
        check(self.lib.mml_wrmf_debug_gram(self.h, G, C.byref(u)))
        return u.value, G

    def stats(self):
        n, ms = C.c_int64(), C.c_float()
        check(self.lib.mml_wrmf_stats(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def recommend(self, users, n=-1, candidates=None, ignore_lists=None, raw=False):
        """ignore_lists: per-user id sequences or a ready (ptr, idx) CSR; raw: return the (items, scores, counts) arrays."""
        users = _i32(users)
        cand = _i32(candidates)
        n_cand = self.n_items if cand is None else cand.shape[0]
        oi, os_, oc, n_out = _topn_outputs(users.shape[0], n, n_cand)
        ptr, idx = _rows_csr(ignore_lists, users.shape[0])
        check(self.lib.mml_wrmf_recommend(self.h, users, users.shape[0], int(n), cand, n_cand, ptr, idx, oi, os_, oc))
        if raw:
            return oi, os_, oc
        return [(oi[b, :oc[b]].copy(), os_[b, :oc[b]].copy()) for b in range(users.shape[0])]

    def evaluate(self, test_users, candidates, test_rows, ignore_rows=None, n=-1):
        """mml_wrmf_evaluate on the device-resident model; see items_evaluate_mf."""
        users, cand = _i32(test_users), _i32(candidates)
        tp, ti = _rows_csr(test_rows, users.shape[0])
        ip, ii = _rows_csr(ignore_rows, users.shape[0])
        out = np.zeros((max(users.shape[0], 1), 8), np.float32)
        used = np.zeros(max(users.shape[0], 1), np.int32)
        check(self.lib.mml_wrmf_evaluate(self.h, users, users.shape[0], cand, cand.shape[0], tp, ti, ip, ii, int(n), out, used))
        return out[:users.shape[0]], used[:users.shape[0]]

    def close(self):
        if self.h:
            if self.ctx.h:
                self.lib.mml_wrmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

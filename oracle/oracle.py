"""ctypes view of the CPU oracle (oracle/mml_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from mymedialite_b200/ (the product).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmmloracle.so")

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

LOSS_RMSE, LOSS_MAE, LOSS_LOGISTIC = 0, 1, 2


class RngState(C.Structure):
    _fields_ = [("seed_array", C.c_int32 * 56), ("inext", C.c_int32), ("inextp", C.c_int32)]


class MFParams(C.Structure):
    _fields_ = [
        ("num_factors", C.c_int32), ("learn_rate", C.c_float), ("decay", C.c_float),
        ("regularization", C.c_float), ("num_iter", C.c_int32),
        ("init_mean", C.c_double), ("init_stddev", C.c_double),
        ("bias_learn_rate", C.c_float), ("bias_reg", C.c_float),
        ("reg_u", C.c_float), ("reg_i", C.c_float),
        ("frequency_regularization", C.c_int32), ("loss", C.c_int32),
        ("max_threads", C.c_int32), ("bold_driver", C.c_int32),
        ("naive_parallelization", C.c_int32), ("omp_threads", C.c_int32),
    ]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "mml_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    R = C.POINTER(RngState)
    P = C.POINTER(MFParams)
    vp = C.c_void_p
    sig = {
        "mo_rng_init": (None, [R, C.c_int32]),
        "mo_rng_next": (C.c_int32, [R]),
        "mo_rng_next_double": (C.c_double, [R]),
        "mo_rng_next_max": (C.c_int32, [R, C.c_int32]),
        "mo_shuffle_i32": (None, [R, i32p, C.c_int64]),
        "mo_shuffle_targets": (None, [R, i32p, C.c_int64]),
        "mo_shuffle_apply": (None, [i32p, i32p, C.c_int64]),
        "mo_normal_sample": (C.c_double, [R, C.c_double, C.c_double]),
        "mo_init_normal": (None, [R, f32p, C.c_int64, C.c_double, C.c_double]),
        "mo_count_by": (None, [i32p, C.c_int64, C.c_int32, i32p]),
        "mo_build_index": (None, [i32p, C.c_int64, C.c_int32, i64p, i32p]),
        "mo_average": (C.c_float, [f32p, C.c_int64]),
        "mo_scale": (None, [f32p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "mo_random_index": (None, [R, i32p, C.c_int64]),
        "mo_partition_users_and_items": (C.c_int32, [R, i32p, i32p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, i64p, i32p, vp, vp]),
        "mo_partition_blocks_given": (None, [i32p, i32p, C.c_int64, i32p, i32p, C.c_int32, i64p, i32p]),
        "mo_partition_indices": (C.c_int32, [i32p, C.c_int64, C.c_int32, i64p, i32p]),
        "mo_mf_params_default": (None, [P]),
        "mo_model_create": (vp, [C.c_int, P, i32p, i32p, f32p, C.c_int64, C.c_int32, C.c_int32]),
        "mo_model_destroy": (None, [vp]),
        "mo_model_init": (None, [vp, R]),
        "mo_model_train": (None, [vp, R]),
        "mo_model_iterate": (None, [vp, R]),
        "mo_model_predict": (C.c_float, [vp, C.c_int32, C.c_int32]),
        "mo_model_predict_many": (None, [vp, i32p, i32p, C.c_int64, f32p]),
        "mo_model_evaluate": (None, [vp, i32p, i32p, f32p, C.c_int64, f32p]),
        "mo_model_objective": (C.c_float, [vp]),
        "mo_model_learnrate": (C.c_float, [vp]),
        "mo_model_global_bias": (C.c_float, [vp]),
        "mo_model_user_factors": (C.POINTER(C.c_float), [vp]),
        "mo_model_item_factors": (C.POINTER(C.c_float), [vp]),
        "mo_model_user_bias": (C.POINTER(C.c_float), [vp]),
        "mo_model_item_bias": (C.POINTER(C.c_float), [vp]),
        "mo_model_random_index": (C.POINTER(C.c_int32), [vp]),
        "mo_model_iterate_indices": (None, [vp, i32p, C.c_int64, C.c_int, C.c_int]),
        "mo_model_fold_in": (None, [vp, i32p, f32p, C.c_int64, f32p, f32p]),
        "mo_model_predict_vector": (C.c_float, [vp, f32p, C.c_int32]),
        "mo_bmf_replay_runs": (None, [vp, i32p, i64p, C.c_int64, i32p, f32p, C.c_int32]),
        "mo_wrmf_optimize": (None, [i64p, i32p, C.c_int32, f32p, f32p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int]),
        "mo_wrmf_gram": (None, [f32p, C.c_int32, C.c_int32, f64p]),
        "mo_feedback_csr": (C.c_int64, [i32p, i32p, C.c_int64, C.c_int32, i64p, i32p]),
        "mo_recommend_mf": (C.c_int64, [f32p, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int64, i32p, C.c_int64, i32p, f32p]),
        "mo_recommend_model": (C.c_int64, [vp, C.c_int32, C.c_int32, i32p, C.c_int64, i32p, C.c_int64, i32p, f32p]),
        "mo_row_scalar_product": (C.c_float, [f32p, f32p, C.c_int32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class Random:
    """System.Random as used through MyMediaLite.Random (Random.cs:23-64)."""

    def __init__(self, seed):
        self.state = RngState()
        lib().mo_rng_init(C.byref(self.state), int(seed))

    @property
    def ref(self):
        return C.byref(self.state)

    def next(self):
        return lib().mo_rng_next(self.ref)

    def next_max(self, m):
        return lib().mo_rng_next_max(self.ref, int(m))

    def next_double(self):
        return lib().mo_rng_next_double(self.ref)

    def shuffle(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        lib().mo_shuffle_i32(self.ref, a, a.size)
        return a

    def shuffle_targets(self, n):
        h = np.empty(n, dtype=np.int32)
        lib().mo_shuffle_targets(self.ref, h, n)
        return h

    def init_normal(self, n, mean=0.0, stddev=0.1):
        d = np.empty(n, dtype=np.float32)
        lib().mo_init_normal(self.ref, d, n, mean, stddev)
        return d


def default_params(**kw):
    p = MFParams()
    lib().mo_mf_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    if "regularization" in kw:           # BiasedMatrixFactorization.cs:97-104: setter fans out
        if "reg_u" not in kw:
            p.reg_u = kw["regularization"]
        if "reg_i" not in kw:
            p.reg_i = kw["regularization"]
    return p


class Model:
    """MatrixFactorization / BiasedMatrixFactorization restated (see mml_oracle.c)."""

    def __init__(self, users, items, values, biased=True, max_user=None, max_item=None, **params):
        self.users = np.ascontiguousarray(users, dtype=np.int32)
        self.items = np.ascontiguousarray(items, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.float32)
        self.max_user = int(self.users.max()) if max_user is None else int(max_user)
        self.max_item = int(self.items.max()) if max_item is None else int(max_item)
        self.params = default_params(**params)
        self.k = self.params.num_factors
        self.biased = biased
        self.h = lib().mo_model_create(int(biased), C.byref(self.params), self.users, self.items, self.values,
                                       self.users.size, self.max_user, self.max_item)

    def __del__(self):
        if getattr(self, "h", None):
            lib().mo_model_destroy(self.h)
            self.h = None

    def init(self, rng):
        lib().mo_model_init(self.h, rng.ref)

    def train(self, rng):
        lib().mo_model_train(self.h, rng.ref)

    def iterate(self, rng):
        lib().mo_model_iterate(self.h, rng.ref)

    def iterate_indices(self, idx, update_user=True, update_item=True):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        lib().mo_model_iterate_indices(self.h, idx, idx.size, int(update_user), int(update_item))

    def replay_runs(self, run_item, run_ptr, ent_user, ent_value, batch):
        run_item = np.ascontiguousarray(run_item, dtype=np.int32)
        run_ptr = np.ascontiguousarray(run_ptr, dtype=np.int64)
        ent_user = np.ascontiguousarray(ent_user, dtype=np.int32)
        ent_value = np.ascontiguousarray(ent_value, dtype=np.float32)
        lib().mo_bmf_replay_runs(self.h, run_item, run_ptr, run_item.size, ent_user, ent_value, int(batch))

    def fold_in(self, items, values, init):
        """FoldIn on (item, rating) pairs already shuffled; init = the InitNormal vector. Returns the user vector."""
        items = np.ascontiguousarray(items, np.int32); values = np.ascontiguousarray(values, np.float32)
        out = np.zeros(self.k + (1 if self.biased else 0), np.float32)
        lib().mo_model_fold_in(self.h, items if items.size else np.zeros(1, np.int32),
                               values if values.size else np.zeros(1, np.float32), items.size,
                               np.ascontiguousarray(init, np.float32), out)
        return out

    def predict_vector(self, vector, item):
        return float(lib().mo_model_predict_vector(self.h, np.ascontiguousarray(vector, np.float32), int(item)))

    def predict(self, u, i):
        return lib().mo_model_predict(self.h, int(u), int(i))

    def predict_many(self, u, i):
        u = np.ascontiguousarray(u, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        out = np.empty(u.size, dtype=np.float32)
        lib().mo_model_predict_many(self.h, u, i, u.size, out)
        return out

    def evaluate(self, u, i, v):
        u = np.ascontiguousarray(u, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        out = np.empty(4, dtype=np.float32)
        lib().mo_model_evaluate(self.h, u, i, v, u.size, out)
        return {"RMSE": float(out[0]), "MAE": float(out[1]), "NMAE": float(out[2]), "CBD": float(out[3])}

    def objective(self):
        return lib().mo_model_objective(self.h)

    @property
    def learnrate(self):
        return lib().mo_model_learnrate(self.h)

    @property
    def global_bias(self):
        return lib().mo_model_global_bias(self.h)

    def _arr(self, ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def user_factors(self):
        return self._arr(lib().mo_model_user_factors(self.h), (self.max_user + 1, self.k))

    @property
    def item_factors(self):
        return self._arr(lib().mo_model_item_factors(self.h), (self.max_item + 1, self.k))

    @property
    def user_bias(self):
        return self._arr(lib().mo_model_user_bias(self.h), (self.max_user + 1,))

    @property
    def item_bias(self):
        return self._arr(lib().mo_model_item_bias(self.h), (self.max_item + 1,))

    @property
    def random_index(self):
        p = lib().mo_model_random_index(self.h)
        if not p:
            return None
        return self._arr(p, (self.users.size,))


def count_by(ids, max_id):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.empty(max_id + 1, dtype=np.int32)
    lib().mo_count_by(ids, ids.size, max_id, out)
    return out


def build_index(ids, max_id):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    ptr = np.empty(max_id + 2, dtype=np.int64)
    idx = np.empty(ids.size, dtype=np.int32)
    lib().mo_build_index(ids, ids.size, max_id, ptr, idx)
    return ptr, idx


def partition_users_and_items(rng, users, items, max_user, max_item, num_groups):
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    g = min(num_groups, max_user + 1, max_item + 1)
    ptr = np.empty(g * g + 1, dtype=np.int64)
    idx = np.empty(users.size, dtype=np.int32)
    up = np.empty(max_user + 1, dtype=np.int32)
    ip = np.empty(max_item + 1, dtype=np.int32)
    g2 = lib().mo_partition_users_and_items(rng.ref, users, items, users.size, max_user, max_item, num_groups,
                                            ptr, idx, up.ctypes.data, ip.ctypes.data)
    assert g2 == g
    return g, ptr, idx, up, ip


def partition_blocks_given(users, items, user_perm, item_perm, g):
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    ptr = np.empty(g * g + 1, dtype=np.int64)
    idx = np.empty(users.size, dtype=np.int32)
    lib().mo_partition_blocks_given(users, items, users.size, np.ascontiguousarray(user_perm, dtype=np.int32),
                                    np.ascontiguousarray(item_perm, dtype=np.int32), g, ptr, idx)
    return ptr, idx


def partition_indices(random_index, num_groups):
    ri = np.ascontiguousarray(random_index, dtype=np.int32)
    g = min(num_groups, ri.size)
    ptr = np.empty(g + 1, dtype=np.int64)
    idx = np.empty(ri.size, dtype=np.int32)
    g2 = lib().mo_partition_indices(ri, ri.size, num_groups, ptr, idx)
    assert g2 == g
    return g, ptr, idx


def feedback_csr(rows, cols, max_row):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    ptr = np.empty(max_row + 2, dtype=np.int64)
    out = np.empty(rows.size, dtype=np.int32)
    nnz = lib().mo_feedback_csr(rows, cols, rows.size, max_row, ptr, out)
    return ptr, out[:nnz].copy()


def wrmf_optimize(row_ptr, cols, W, H, alpha=1.0, regularization=0.015, omp_threads=1):
    """In-place ALS half-sweep on W (WRMF.cs:79-156)."""
    assert W.dtype == np.float32 and H.dtype == np.float32
    lib().mo_wrmf_optimize(np.ascontiguousarray(row_ptr, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int32),
                           W.shape[0], W, H, H.shape[0], W.shape[1], float(alpha), float(regularization), int(omp_threads))


def wrmf_gram(H):
    k = H.shape[1]
    out = np.empty((k, k), dtype=np.float64)
    lib().mo_wrmf_gram(np.ascontiguousarray(H, dtype=np.float32), H.shape[0], k, out)
    return out


def recommend_mf(U, V, user, n=-1, candidates=None, ignore=None):
    U = np.ascontiguousarray(U, dtype=np.float32)
    V = np.ascontiguousarray(V, dtype=np.float32)
    if candidates is None:
        candidates = np.arange(V.shape[0], dtype=np.int32)
    candidates = np.ascontiguousarray(candidates, dtype=np.int32)
    ignore = np.ascontiguousarray(ignore if ignore is not None else [], dtype=np.int32)
    oi = np.empty(max(candidates.size, 1), dtype=np.int32)
    os_ = np.empty(max(candidates.size, 1), dtype=np.float32)
    cnt = lib().mo_recommend_mf(U, V, U.shape[1], V.shape[0], U.shape[0], int(user), int(n), candidates, candidates.size,
                                ignore, ignore.size, oi, os_)
    return oi[:cnt].copy(), os_[:cnt].copy()


def recommend_model(model, user, n=-1, candidates=None, ignore=None):
    if candidates is None:
        candidates = np.arange(model.max_item + 1, dtype=np.int32)
    candidates = np.ascontiguousarray(candidates, dtype=np.int32)
    ignore = np.ascontiguousarray(ignore if ignore is not None else [], dtype=np.int32)
    oi = np.empty(max(candidates.size, 1), dtype=np.int32)
    os_ = np.empty(max(candidates.size, 1), dtype=np.float32)
    cnt = lib().mo_recommend_model(model.h, int(user), int(n), candidates, candidates.size, ignore, ignore.size, oi, os_)
    return oi[:cnt].copy(), os_[:cnt].copy()


# ---- file readers (test infrastructure for csrc/ingest.cu) ----------------------------------------------------------
_INT_RE = re.compile(r"^[\t-\r ]*[+-]?[0-9]+[\t-\r ]*$")
_FLOAT_RE = re.compile(r"^[\t-\r ]*[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?[\t-\r ]*$")


class FormatException(Exception):
    """System.FormatException raised by the reference's readers."""


def _net_int_parse(tok):
    """int.Parse(string), NumberStyles.Integer (Data/IdentityMapping.cs:64)."""
    if not _INT_RE.match(tok):
        raise FormatException("Input string was not in a correct format.")
    v = int(tok.strip("\t\n\v\f\r "))
    if not -2 ** 31 <= v < 2 ** 31:
        raise FormatException("Value was either too large or too small for an Int32.")
    return v


def _net_single_parse(tok):
    """float.Parse(string, InvariantCulture) of the .NET Framework / Mono: parse to double, cast to float
    (IO/StaticRatingData.cs:112)."""
    t = tok.strip("\t\n\v\f\r ")
    if t == "NaN":
        return np.float32(np.nan)
    if t in ("Infinity", "-Infinity"):
        return np.float32(np.inf if t[0] != "-" else -np.inf)
    if not _FLOAT_RE.match(tok):
        raise FormatException("Input string was not in a correct format.")
    with np.errstate(over="ignore"):
        f = np.float32(float(t))
    if np.isinf(f):
        raise FormatException("Value was either too large or too small for a Single.")
    return f


class FirstSeenMapping:
    """Data/Mapping.cs:75-85."""

    def __init__(self):
        self.original_to_internal = {}
        self.internal_to_original = []

    def to_internal(self, tok):
        if tok in self.original_to_internal:
            return self.original_to_internal[tok]
        i = len(self.original_to_internal)
        self.original_to_internal[tok] = i
        self.internal_to_original.append(tok)
        return i


class IdentityMapping:
    """Data/IdentityMapping.cs:62-67."""

    def to_internal(self, tok):
        return _net_int_parse(tok)


def read_lines(text):
    """TextReader.ReadLine: a line ends at \\n, \\r or \\r\\n; a trailing terminator does not start another line."""
    if text.startswith("\ufeff"):       # StreamReader drops the byte order mark
        text = text[1:]
    parts = re.split(r"\r\n|\n|\r", text)
    if parts and parts[-1] == "":
        parts.pop()
    return parts


def read_rating_text(text, user_mapping=None, item_mapping=None, with_ratings=True, ignore_first_line=False):
    """IO/StaticRatingData.cs:82-117 (IO/RatingData.cs:57-88 has the same line handling)."""
    um = user_mapping if user_mapping is not None else IdentityMapping()
    im = item_mapping if item_mapping is not None else IdentityMapping()
    lines = read_lines(text)
    if ignore_first_line:
        lines = lines[1:]
    users, items, values = [], [], []
    for line in lines:
        if len(line) == 0:
            continue
        tokens = re.split(r"[\t ,]", line)                      # string.Split(char[]) keeps empty tokens
        if with_ratings and len(tokens) < 3:
            raise FormatException("Expected at least 3 columns: " + line)
        if not with_ratings and len(tokens) < 2:
            raise FormatException("Expected at least 2 columns: " + line)
        users.append(um.to_internal(tokens[0]))
        items.append(im.to_internal(tokens[1]))
        values.append(_net_single_parse(tokens[2]) if with_ratings else np.float32(0))
    return np.array(users, np.int32), np.array(items, np.int32), np.array(values, np.float32)


def read_feedback_text(text, user_mapping=None, item_mapping=None, ignore_first_line=False):
    """IO/ItemData.cs:59-93."""
    um = user_mapping if user_mapping is not None else IdentityMapping()
    im = item_mapping if item_mapping is not None else IdentityMapping()
    lines = read_lines(text)
    if ignore_first_line:
        lines = lines[1:]
    users, items = [], []
    for line in lines:
        if len(line.strip("\t\n\v\f\r ")) == 0:
            continue
        tokens = re.split(r"[\t ,]", line)
        if len(tokens) < 2:
            raise FormatException("Expected at least 2 columns: " + line)
        try:
            users.append(um.to_internal(tokens[0]))
            items.append(im.to_internal(tokens[1]))
        except FormatException:
            raise FormatException("Could not read line '%s'" % line)
    return np.array(users, np.int32), np.array(items, np.int32)


# ---- ranking measures and Eval.Items.Evaluate (test infrastructure for items_eval_kernel) -----------------------------
def auc_compute(ranked_items, relevant_items, num_dropped_items):
    """Eval/Measures/AUC.cs:39-69."""
    relevant = set(relevant_items)
    num_relevant_items = len(relevant & set(ranked_items))
    num_eval_items = len(ranked_items) + num_dropped_items
    num_eval_pairs = (num_eval_items - num_relevant_items) * num_relevant_items
    if num_eval_pairs < 0:
        raise Exception("num_eval_pairs cannot be less than 0")
    if num_eval_pairs == 0:
        return 0.5
    num_correct_pairs = 0
    hit_count = 0
    for item_id in ranked_items:
        if item_id not in relevant:
            num_correct_pairs += hit_count
        else:
            hit_count += 1
    missing_relevant_items = len(relevant - set(ranked_items))
    if num_dropped_items - missing_relevant_items < 0:
        raise Exception("Should not happen.")
    num_correct_pairs += hit_count * (num_dropped_items - missing_relevant_items)
    return num_correct_pairs / num_eval_pairs


def ap_compute(ranked_items, correct_items):
    """Eval/Measures/PrecisionAndRecall.cs AP."""
    correct = set(correct_items)
    hit_count, avg_prec_sum = 0, 0.0
    for i, item_id in enumerate(ranked_items):
        if item_id in correct:
            hit_count += 1
            avg_prec_sum += hit_count / (i + 1)
    return avg_prec_sum / len(correct) if hit_count != 0 else 0.0


def hits_at(ranked_items, correct_items, n):
    """PrecisionAndRecall.HitsAt."""
    if n < 1:
        raise ValueError("n must be at least 1.")
    correct = set(correct_items)
    hit_count = 0
    for i, item_id in enumerate(ranked_items):
        if item_id not in correct:
            continue
        if i < n:
            hit_count += 1
        else:
            break
    return hit_count


def precision_at(ranked_items, correct_items, n):
    return hits_at(ranked_items, correct_items, n) / n


def recall_at(ranked_items, correct_items, n):
    return hits_at(ranked_items, correct_items, n) / len(set(correct_items))


def ndcg_compute(ranked_items, correct_items):
    """Eval/Measures/NDCG.cs (Math.Log(x, 2) = Log(x) / Log(2))."""
    import math
    correct = set(correct_items)
    dcg = 0.0
    idcg = 0.0
    for i in range(len(correct)):
        idcg += 1 / (math.log(i + 2) / math.log(2))
    for i, item_id in enumerate(ranked_items):
        if item_id not in correct:
            continue
        dcg += 1 / (math.log(i + 2) / math.log(2))
    return dcg / idcg


def reciprocal_rank(ranked_items, correct_items):
    """Eval/Measures/ReciprocalRank.cs."""
    correct = set(correct_items)
    for pos, item_id in enumerate(ranked_items):
        if item_id in correct:
            return 1.0 / (pos + 1)
    return 0.0


ITEM_MEASURES = ["AUC", "MAP", "NDCG", "MRR", "prec@5", "prec@10", "recall@5", "recall@10"]


def first_seen(ids):
    """HashSet<int> filled in order and copied out (Data/DataSet.cs:112-131): distinct ids in order of first appearance."""
    seen, out = set(), []
    for x in ids:
        x = int(x)
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


def items_evaluate(recommend, test_users_ids, test_items_ids, train_users_ids, train_items_ids, test_users=None,
                   candidate_items=None, repeated_events=False, n=-1):
    """Eval/Items.cs:126-209 with the candidate list already fixed by the caller (Items.Candidates + its shuffle).
    recommend(user, n, ignore_items, candidate_items) -> [(item, score)] as Recommender.Recommend.
    Returns (results dict, per-user rows) -- the sums are float32 additions in test_users order (single-threaded order)."""
    test_rows, train_rows = {}, {}
    for u, i in zip(test_users_ids, test_items_ids):
        test_rows.setdefault(int(u), set()).add(int(i))
    for u, i in zip(train_users_ids, train_items_ids):
        train_rows.setdefault(int(u), set()).add(int(i))
    if test_users is None:
        test_users = first_seen(test_users_ids)
    cand_set = set(int(c) for c in candidate_items)
    sums = {m: np.float32(0) for m in ITEM_MEASURES}
    rows = {}
    num_users = 0
    for user_id in test_users:
        correct_items = test_rows.get(user_id, set()) & cand_set
        if len(correct_items) == 0:
            continue
        ignore = set() if repeated_events else set(train_rows.get(user_id, set()))
        ignore &= cand_set
        num_candidates_for_this_user = len(candidate_items) - len(ignore)
        if len(correct_items) == num_candidates_for_this_user:
            continue
        prediction = recommend(user_id, n, ignore, candidate_items)
        prediction_list = [t[0] for t in prediction]
        num_dropped_items = num_candidates_for_this_user - len(prediction)
        row = [auc_compute(prediction_list, correct_items, num_dropped_items), ap_compute(prediction_list, correct_items),
               ndcg_compute(prediction_list, correct_items), reciprocal_rank(prediction_list, correct_items),
               precision_at(prediction_list, correct_items, 5), precision_at(prediction_list, correct_items, 10),
               recall_at(prediction_list, correct_items, 5), recall_at(prediction_list, correct_items, 10)]
        num_users += 1
        rows[user_id] = np.array(row, np.float32)
        for m, x in zip(ITEM_MEASURES, row):
            sums[m] = np.float32(sums[m] + np.float32(x))
    with np.errstate(invalid="ignore", divide="ignore"):
        result = {m: np.float32(sums[m] / np.float32(num_users)) for m in ITEM_MEASURES}
    result["num_users"] = num_users
    result["num_lists"] = num_users
    result["num_items"] = len(candidate_items)
    return result, rows

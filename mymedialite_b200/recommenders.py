"""Host-side mirror of the reference's recommender surface for the CUDA-backed classes.

The reference is C#; this image has no .NET/Mono, so the classes a maintainer would add to MyMediaLite.dll
(csharp/*.cs, bound through P/Invoke) are mirrored here in Python over the same C ABI, with
the same property names, defaults, call order and error behaviour, so that the tests read like the reference's own
(src/Tests/RatingPrediction/BiasedMatrixFactorizationTest.cs, RatingPredictorsTest.cs, ...).

  RatingPrediction.MatrixFactorization        -> MatrixFactorization        (MatrixFactorization.cs:35-418)
  RatingPrediction.BiasedMatrixFactorization  -> BiasedMatrixFactorization  (BiasedMatrixFactorization.cs:61-563)
  ItemRecommendation.WRMF                     -> WRMF                       (ItemRecommendation/WRMF.cs, MF.cs)

There is no CPU fallback: every Train / Iterate / Predict / Recommend runs in libmmlb200.so on the GPU.
"""
import numpy as np

from . import _capi, engine, modelio, sysrandom

import os

_ctx = {}


def context(n_gpus=1):
    """One library context per process and GPU count (mml_ctx): GPUs 0 .. n_gpus - 1, one process driving all of them."""
    n_gpus = max(int(n_gpus), 1)
    if n_gpus not in _ctx:
        _ctx[n_gpus] = engine.Context(0, n_gpus=n_gpus)
    return _ctx[n_gpus]


# The one documented engine knob (process-wide, not a recommender option: the option set stays the reference's + NumGpus).
#   order: "auto" (default)  -- Iterate() runs the parallel epoch kernel (the DSGD block schedule, whatever MaxThreads says;
#                               MaxThreads keeps its other meaning, UpdateLearnRate twice per epoch when > 1), except on
#                               data sets below SERIAL_BELOW ratings, where the exact single-threaded order costs nothing;
#          "reference"       -- MaxThreads = 1 walks RandomIndex in the reference's order on one warp (parity runs);
#          "parallel"        -- always the parallel kernel.
#   init:  "host" (default)  -- InitModel draws from MyMediaLite.Random on the host, in the reference's order;
#          "device"          -- counter-based generator on the device (large models: no host sampling).
# Environment: MMLB200_ORDER, MMLB200_INIT.
ENGINE = {"order": os.environ.get("MMLB200_ORDER", "auto"), "init": os.environ.get("MMLB200_INIT", "host")}
SERIAL_BELOW = 20000


def set_engine(order=None, init=None):
    if order is not None:
        if order not in ("auto", "reference", "parallel"):
            raise ValueError("order must be auto, reference or parallel")
        ENGINE["order"] = order
    if init is not None:
        if init not in ("host", "device"):
            raise ValueError("init must be host or device")
        ENGINE["init"] = init


class Ratings:
    """IRatings (Data/IRatings.cs, Data/Ratings.cs, Data/StaticRatings.cs): COO triples + MaxUserID / MaxItemID."""

    def __init__(self, users=(), items=(), values=()):
        self.Users = np.ascontiguousarray(users, np.int32)
        self.Items = np.ascontiguousarray(items, np.int32)
        self.Values = np.ascontiguousarray(values, np.float32)
        self.MaxUserID = int(self.Users.max()) if self.Users.size else -1
        self.MaxItemID = int(self.Items.max()) if self.Items.size else -1
        self._random_index = None

    @property
    def Count(self):
        return int(self.Users.size)

    def Add(self, user, item, value):
        self.Users = np.append(self.Users, np.int32(user))
        self.Items = np.append(self.Items, np.int32(item))
        self.Values = np.append(self.Values, np.float32(value))
        self.MaxUserID = max(self.MaxUserID, int(user))
        self.MaxItemID = max(self.MaxItemID, int(item))
        self._random_index = None

    @property
    def RandomIndex(self):
        """DataSet.RandomIndex (Data/DataSet.cs:100-109, 193-202): shuffled once, rebuilt only when Count changes.
        The swap targets come from MyMediaLite.Random on the host; the permutation is applied on the device."""
        if self._random_index is None or self._random_index.size != self.Count:
            H = sysrandom.get_instance().shuffle_targets(self.Count)
            perm = np.arange(self.Count, dtype=np.int32)
            if self.Count:
                ctx = context()
                _capi.check(ctx.lib.mml_shuffle_apply(ctx.h, perm, H, self.Count))
            self._random_index = perm
        return self._random_index


class PosOnlyFeedback:
    """IPosOnlyFeedback (Data/PosOnlyFeedback.cs): (user, item) events; the user / item matrices are sets."""

    def __init__(self, users=(), items=()):
        self.Users = np.ascontiguousarray(users, np.int32)
        self.Items = np.ascontiguousarray(items, np.int32)
        self.MaxUserID = int(self.Users.max()) if self.Users.size else -1
        self.MaxItemID = int(self.Items.max()) if self.Items.size else -1

    @property
    def Count(self):
        return int(self.Users.size)


def _net_bool(b):
    return "True" if b else "False"


class _Recommender:
    MaxUserID = -1
    MaxItemID = -1

    def CanPredict(self, user_id, item_id):
        return user_id <= self.MaxUserID and item_id <= self.MaxItemID


class MatrixFactorization(_Recommender):
    """RatingPrediction.MatrixFactorization on the GPU: r = global_bias + p_u . q_i, SGD."""
    _biased = 0
    _type_name = "MyMediaLite.RatingPrediction.CudaMatrixFactorization"

    def __init__(self):
        # MatrixFactorization.cs:87-96
        self.Regularization = 0.015
        self.LearnRate = 0.01
        self.Decay = 1.0
        self.NumIter = 30
        self.InitStdDev = 0.1
        self.InitMean = 0.0
        self.NumFactors = 10
        # the one addition to the reference's option set: GPUs this process trains on (mml_ctx_create(n_gpus))
        self.NumGpus = 1
        self._is_parallel = False
        self.MinRating = 1.0
        self.MaxRating = 5.0
        self.Ratings = None
        self._model = None
        self._dev_ratings = None

    # -- plumbing -----------------------------------------------------------------------------------------------
    def _parallel(self):
        """Whether Iterate() runs the parallel epoch kernel (see ENGINE above). Several GPUs always do."""
        if int(self.NumGpus) > 1 or ENGINE["order"] == "parallel":
            return True
        if ENGINE["order"] == "reference":
            return getattr(self, "MaxThreads", 1) > 1
        n = self.Ratings.Count if self.Ratings is not None else 0
        return getattr(self, "MaxThreads", 1) > 1 or n >= SERIAL_BELOW

    def _params(self):
        return engine.default_params(
            biased=self._biased, num_factors=int(self.NumFactors), learn_rate=float(self.LearnRate), decay=float(self.Decay),
            regularization=float(self.Regularization),
            schedule=_capi.SCHEDULE_DSGD if self._parallel() else _capi.SCHEDULE_SERIAL)

    def _init_model(self):
        """InitModel (MatrixFactorization.cs:99-116): a fresh device model (Train() allocates a new handle, it never
        mutates one a Clone() may share)."""
        if self.Ratings is None:
            raise ValueError("Ratings is not set")
        r = self.Ratings
        self.MaxUserID, self.MaxItemID = r.MaxUserID, r.MaxItemID
        ctx = context(self.NumGpus)
        self._dev_ratings = engine.DeviceRatings(ctx, r.Users, r.Items, r.Values, r.MaxUserID, r.MaxItemID)
        if r.Count:
            _, self.MinRating, self.MaxRating = self._dev_ratings.stats()
        params = self._params()
        self._model = engine.SgdModel(ctx, self._dev_ratings, params)
        self._is_parallel = params.schedule == _capi.SCHEDULE_DSGD
        if ENGINE["init"] == "device":
            self._model.init_model(sysrandom.get_instance().next(), self.InitMean, self.InitStdDev)
        else:
            rng = sysrandom.get_instance()
            U = rng.init_normal(r.MaxUserID + 1, int(self.NumFactors), self.InitMean, self.InitStdDev)
            V = rng.init_normal(r.MaxItemID + 1, int(self.NumFactors), self.InitMean, self.InitStdDev)
            self._model.set_model(U, V)

    InitModel = _init_model

    @property
    def current_learnrate(self):
        return self._model.learnrate

    # -- IRecommender / IIterativeModel ------------------------------------------------------------------------
    def Train(self):
        self._init_model()
        for _ in range(int(self.NumIter)):
            self.Iterate()

    def Iterate(self):
        if self._is_parallel:
            G = self._model.strata_info()["G"]
            seq = sysrandom.get_instance().shuffle(np.arange(G))       # BiasedMatrixFactorization.cs:210-211
            self._model.iterate(subepoch_sequence=seq)
        else:
            self._model.iterate(random_index=self.Ratings.RandomIndex)

    def Predict(self, user_id, item_id):
        return float(self._model.predict([user_id], [item_id])[0])

    def PredictMany(self, users, items):
        return self._model.predict(users, items)

    def Evaluate(self, test):
        """Eval.Ratings.Evaluate (Eval/Ratings.cs:73-139) on the device."""
        return self._model.evaluate(test.Users, test.Items, test.Values)

    def Recommend(self, user_id, n=-1, ignore_items=None, candidate_items=None):
        """Recommender.Recommend (Recommender.cs:52-103) with this predictor's Predict as score."""
        if candidate_items is None:
            candidate_items = np.arange(0, max(self.MaxItemID - 1, 0), dtype=np.int32)   # :57-58, the reference's own default
        cand = np.ascontiguousarray(candidate_items, np.int32)
        if ignore_items is not None and len(ignore_items):
            cand_ok = cand[~np.isin(cand, np.asarray(list(ignore_items), np.int32))]
        else:
            cand_ok = cand
        scores = self._model.predict(np.full(cand_ok.size, user_id, np.int32), cand_ok)
        order = np.argsort(-scores, kind="stable")
        if n >= 0:
            order = order[:n]
        return [(int(cand_ok[t]), float(scores[t])) for t in order]

    def ComputeObjective(self):
        return self._model.objective()

    # -- IncrementalRatingPredictor / IFoldInRatingPredictor ---------------------------------------------------
    UpdateUsers = True
    UpdateItems = True

    def _retrain(self, entity_id, by_item):
        """RetrainUser / RetrainItem (MatrixFactorization.cs:141-160, BiasedMatrixFactorization.cs:419-431): the row is
        re-drawn (RowInitNormal, host RNG), the bias zeroed (biased model), then LearnFactors (:198-202): NumIter passes
        over ByUser[u] / ByItem[i] in ascending rating-index order updating that side only; plain MF decays the learn
        rate after every pass (:195)."""
        row = sysrandom.get_instance().init_normal(1, int(self.NumFactors), self.InitMean, self.InitStdDev)
        self._model.set_rows([entity_id], row, [0.0] if self._biased else None, by_item=by_item)
        ids = self.Ratings.Items if by_item else self.Ratings.Users
        idx = np.nonzero(ids == entity_id)[0].astype(np.int32)
        self._model.learn_factors(idx, int(self.NumIter), update_user=not by_item, update_item=by_item)

    def RetrainUser(self, user_id):
        if self.UpdateUsers:
            self._retrain(user_id, False)

    def RetrainItem(self, item_id):
        if self.UpdateItems:
            self._retrain(item_id, True)

    def RemoveUser(self, user_id):
        """MatrixFactorization.cs:300-306 / BiasedMatrixFactorization.cs:433-438: the row (and bias) is set to zero."""
        self._model.set_rows([user_id], np.zeros((1, int(self.NumFactors)), np.float32), [0.0] if self._biased else None)

    def RemoveItem(self, item_id):
        self._model.set_rows([item_id], np.zeros((1, int(self.NumFactors)), np.float32), [0.0] if self._biased else None,
                             by_item=True)

    def FoldInMany(self, rated_items_per_user):
        """FoldIn (MatrixFactorization.cs:323-347, BiasedMatrixFactorization.cs:445-492) for several new users in one
        launch. The RNG draws happen user by user in the reference's order: the InitNormal vector, then the shuffle."""
        rng = sysrandom.get_instance()
        k = int(self.NumFactors)
        init = np.zeros((len(rated_items_per_user), k), np.float32)
        items, values = [], []
        for j, rated in enumerate(rated_items_per_user):
            init[j] = rng.init_normal(1, k, self.InitMean, self.InitStdDev)[0]
            order = rng.shuffle(np.arange(len(rated)))
            items.append([rated[t][0] for t in order])
            values.append([rated[t][1] for t in order])
        return self._model.fold_in(items, values, init, int(self.NumIter))

    def FoldIn(self, rated_items):
        return self.FoldInMany([rated_items])[0]

    def ScoreItems(self, rated_items, candidate_items=None):
        """MatrixFactorization.cs:350-363; candidates default as FoldInRatingPredictorExtensions.cs:64 (0..MaxItemID-2)."""
        if candidate_items is None:
            candidate_items = np.arange(0, max(self.MaxItemID - 1, 0), dtype=np.int32)
        cand = np.ascontiguousarray(candidate_items, np.int32)
        scores = self._model.score_items(self.FoldIn(rated_items), cand)[0]
        return [(int(i), float(s)) for i, s in zip(cand, scores)]

    def RecommendItems(self, rated_items, candidate_items=None, n=-1):
        """FoldInRatingPredictorExtensions.cs:35-51: OrderByDescending (stable) then Take(n)."""
        scored = self.ScoreItems(rated_items, candidate_items)
        order = np.argsort(-np.array([s for _, s in scored], np.float32), kind="stable")
        if n >= 0:
            order = order[:n]
        return [scored[t] for t in order]

    # -- model files ------------------------------------------------------------------------------------------
    def SaveModel(self, filename):
        m = self._model.get_model()
        with open(filename, "w") as w:
            modelio.write_header(w, self._type_name)
            w.write(modelio.fmt(m["global_bias"]) + "\n")
            modelio.write_matrix(w, m["U"])
            modelio.write_matrix(w, m["V"])

    def LoadModel(self, filename):
        with open(filename) as r:
            modelio.read_header(r, self._type_name)
            bias = np.float32(float(r.readline()))
            U, V = modelio.read_matrix(r), modelio.read_matrix(r)
        self._adopt(U, V, None, None, bias, None, None)

    def _adopt(self, U, V, bu, bi, bias, min_rating, max_rating):
        if U.shape[1] != V.shape[1]:
            raise IOError("Number of user and item factors must match: %d != %d" % (U.shape[1], V.shape[1]))
        self.MaxUserID, self.MaxItemID = U.shape[0] - 1, V.shape[0] - 1
        self.NumFactors = U.shape[1]
        if min_rating is not None:
            self.MinRating, self.MaxRating = float(min_rating), float(max_rating)
        # a model without training data: one pseudo rating per id keeps every row (rows without ratings are zeroed
        # by InitModel only) and carries the rating scale
        ctx = context()
        nu, ni = U.shape[0], V.shape[0]
        n = max(nu, ni)
        uu = (np.arange(n) % nu).astype(np.int32); ii = (np.arange(n) % ni).astype(np.int32)
        vv = np.full(n, self.MinRating, np.float32)
        self._dev_ratings = engine.DeviceRatings(ctx, uu, ii, vv, nu - 1, ni - 1)
        p = self._params()
        p.schedule = _capi.SCHEDULE_SERIAL
        self._is_parallel = False
        self._model = engine.SgdModel(ctx, self._dev_ratings, p)
        self._model.set_model(U, V, bu, bi)
        _capi.check(ctx.lib.mml_sgd_set_scale(self._model.h, float(self.MinRating), float(self.MaxRating), float(bias)))

    def ToString(self):
        return "%s num_factors=%d regularization=%s learn_rate=%s learn_rate_decay=%s num_iter=%d" % (
            type(self).__name__, self.NumFactors, modelio.fmt(self.Regularization), modelio.fmt(self.LearnRate),
            modelio.fmt(self.Decay), self.NumIter)

    __str__ = ToString


class BiasedMatrixFactorization(MatrixFactorization):
    """RatingPrediction.BiasedMatrixFactorization on the GPU (sigmoid link, biases, RMSE / MAE / logistic loss)."""
    _biased = 1
    _type_name = "MyMediaLite.RatingPrediction.CudaBiasedMatrixFactorization"
    _LOSS = {"RMSE": _capi.LOSS_RMSE, "MAE": _capi.LOSS_MAE, "LogisticLoss": _capi.LOSS_LOGISTIC}

    def __init__(self):
        super().__init__()
        # BiasedMatrixFactorization.cs:85-141
        self.BiasReg = 0.01
        self.BiasLearnRate = 1.0
        self.RegU = 0.015
        self.RegI = 0.015
        self.FrequencyRegularization = False
        self.Loss = "RMSE"
        self.MaxThreads = 1
        self.BoldDriver = False
        self.NaiveParallelization = False

    def __setattr__(self, name, value):
        if name == "Regularization":          # :97-104: the setter fans out to RegU / RegI
            object.__setattr__(self, "RegU", value)
            object.__setattr__(self, "RegI", value)
        object.__setattr__(self, name, value)

    def _params(self):
        # MaxThreads > 1 selects the reference's DSGD block schedule (:178-184); on the GPU the worker groups are CTAs, and
        # the parallel kernel is also what MaxThreads = 1 runs unless the engine order says "reference" (ENGINE above)
        dsgd = self._parallel()
        # NaiveParallelization (:136-141, :201-204): RandomIndex dealt into lists (MultiCore.PartitionIndices), one per worker,
        # all walked at once with no exclusivity -- MML_SCHEDULE_NAIVE (one GPU; several GPUs keep the block schedule)
        naive = dsgd and self.NaiveParallelization and int(self.NumGpus) <= 1
        return engine.default_params(**dict(
            biased=1, num_factors=int(self.NumFactors), learn_rate=float(self.LearnRate), decay=float(self.Decay),
            regularization=float(self.Regularization), bias_learn_rate=float(self.BiasLearnRate), bias_reg=float(self.BiasReg),
            reg_u=float(self.RegU), reg_i=float(self.RegI), frequency_regularization=int(bool(self.FrequencyRegularization)),
            loss=self._LOSS[self.Loss], bold_driver=int(bool(self.BoldDriver)), max_threads=int(self.MaxThreads),
            schedule=_capi.SCHEDULE_NAIVE if naive else (_capi.SCHEDULE_DSGD if dsgd else _capi.SCHEDULE_SERIAL)))

    def SaveModel(self, filename):
        m = self._model.get_model()
        with open(filename, "w") as w:
            modelio.write_header(w, self._type_name)
            w.write(modelio.fmt(m["global_bias"]) + "\n")
            w.write(modelio.fmt(self.MinRating) + "\n")
            w.write(modelio.fmt(self.MaxRating) + "\n")
            modelio.write_vector(w, m["bu"])
            modelio.write_matrix(w, m["U"])
            modelio.write_vector(w, m["bi"])
            modelio.write_matrix(w, m["V"])

    def LoadModel(self, filename):
        with open(filename) as r:
            modelio.read_header(r, self._type_name)
            bias = np.float32(float(r.readline()))
            mn = np.float32(float(r.readline())); mx = np.float32(float(r.readline()))
            bu = modelio.read_vector(r); U = modelio.read_matrix(r)
            bi = modelio.read_vector(r); V = modelio.read_matrix(r)
        if bu.size != U.shape[0]:
            raise IOError("Number of users must be the same for biases and factors: %d != %d" % (bu.size, U.shape[0]))
        if bi.size != V.shape[0]:
            raise IOError("Number of items must be the same for biases and factors: %d != %d" % (bi.size, V.shape[0]))
        self._adopt(U, V, bu, bi, bias, mn, mx)

    def ToString(self):
        return ("%s num_factors=%d bias_reg=%s reg_u=%s reg_i=%s frequency_regularization=%s learn_rate=%s "
                "bias_learn_rate=%s learn_rate_decay=%s num_iter=%d bold_driver=%s loss=%s max_threads=%d "
                "naive_parallelization=%s" % (
                    type(self).__name__, self.NumFactors, modelio.fmt(self.BiasReg), modelio.fmt(self.RegU), modelio.fmt(self.RegI),
                    _net_bool(self.FrequencyRegularization), modelio.fmt(self.LearnRate), modelio.fmt(self.BiasLearnRate),
                    modelio.fmt(self.Decay), self.NumIter, _net_bool(self.BoldDriver), self.Loss, self.MaxThreads,
                    _net_bool(self.NaiveParallelization)))

    __str__ = ToString


class WRMF(_Recommender):
    """ItemRecommendation.WRMF on the GPU (ItemRecommendation/WRMF.cs:56-180, MF.cs:37-196)."""
    _type_name = "MyMediaLite.ItemRecommendation.CudaWRMF"

    def __init__(self):
        self.NumFactors = 10
        self.NumIter = 15
        self.Alpha = 1.0
        self.Regularization = 0.015
        self.InitMean = 0.0
        self.InitStdDev = 0.1
        self.NumGpus = 1
        self.Feedback = None
        self._model = None
        self._fb = None
        self._cache = None          # all-users top-N of the current model: (n, candidate key, train rows, items, scores, counts)

    def _new_model(self, n_users, n_items, users, items):
        ctx = context(self.NumGpus)
        self._fb = engine.DeviceFeedback(ctx, users, items, n_users - 1, n_items - 1)
        self._model = engine.WrmfModel(ctx, self._fb, int(self.NumFactors), float(self.Alpha), float(self.Regularization))

    def InitModel(self):
        f = self.Feedback
        self.MaxUserID, self.MaxItemID = f.MaxUserID, f.MaxItemID
        self._new_model(f.MaxUserID + 1, f.MaxItemID + 1, f.Users, f.Items)
        if ENGINE["init"] == "device":
            self._model.init_model(sysrandom.get_instance().next(), self.InitMean, self.InitStdDev)
        else:
            rng = sysrandom.get_instance()   # MF.cs:56-57: user matrix first, no zeroing of empty rows
            U = rng.init_normal(f.MaxUserID + 1, int(self.NumFactors), self.InitMean, self.InitStdDev)
            V = rng.init_normal(f.MaxItemID + 1, int(self.NumFactors), self.InitMean, self.InitStdDev)
            self._model.set_model(U, V)

    def Train(self):
        self.InitModel()
        for _ in range(int(self.NumIter)):
            self.Iterate()

    def Iterate(self):
        self._model.iterate()
        self._cache = None

    # -- per-user Recommend() served from one batched all-users call ------------------------------------------------------
    # Eval.Items.Evaluate (Eval/Items.cs:147-164) and WritePredictions (ItemRecommendation/Extensions.cs:65-128) call
    # Recommend(user, n, ignore = the user's training items, candidates) once per user, from TPL threads. The first such call
    # after the model changed computes the lists of ALL users in one device call (ignore rows = the training feedback) and
    # keeps them; every later call with the same n and candidates whose ignore list is the user's training row is a lookup.
    def _train_rows(self):
        f = self.Feedback
        if f is None or f.Count == 0:
            return None
        order = np.lexsort((f.Items, f.Users))
        u, i = f.Users[order], f.Items[order]
        keep = np.ones(u.size, bool)
        keep[1:] = (u[1:] != u[:-1]) | (i[1:] != i[:-1])          # the user matrix is a set (PosOnlyFeedback.cs:35-83)
        u, i = u[keep], i[keep]
        ptr = np.zeros(self.MaxUserID + 2, np.int64)
        np.add.at(ptr, u + 1, 1)
        return np.cumsum(ptr), i.astype(np.int32)

    def _cached(self, user_id, n, ignore_items, cand):
        if n <= 0 or user_id < 0 or user_id > self.MaxUserID or self.Feedback is None:
            return None
        key = (int(n), cand.size, hash(cand.tobytes()))
        c = self._cache
        if c is None or c["key"] != key:
            rows = self._train_rows()
            if rows is None:
                return None
            c = dict(key=key, ptr=rows[0], idx=rows[1], lists=None)
        ptr, idx = c["ptr"], c["idx"]
        row = idx[ptr[user_id]:ptr[user_id + 1]]
        given = np.unique(np.asarray(list(ignore_items) if ignore_items is not None else [], np.int32))
        if given.size != row.size or not np.array_equal(given, row):
            return None                       # some other ignore list: answered directly
        if c["lists"] is None:
            users = np.arange(self.MaxUserID + 1, dtype=np.int32)
            c["lists"] = self._model.recommend(users, n, cand, (ptr, idx if idx.size else np.zeros(1, np.int32)), raw=True)
            self._cache = c
        oi, os_, oc = c["lists"]
        k = int(oc[user_id])
        return [(int(a), float(b)) for a, b in zip(oi[user_id, :k], os_[user_id, :k])]

    def Predict(self, user_id, item_id):
        # MF.cs:151-157
        if user_id > self.MaxUserID or item_id > self.MaxItemID or user_id < 0 or item_id < 0:
            return float(np.finfo(np.float32).min)
        res = self._model.recommend([user_id], 1, [item_id])[0]
        return float(res[1][0]) if len(res[1]) else float(np.finfo(np.float32).min)

    def Recommend(self, user_id, n=-1, ignore_items=None, candidate_items=None):
        if candidate_items is None:
            candidate_items = np.arange(0, max(self.MaxItemID - 1, 0), dtype=np.int32)   # Recommender.cs:57-58
        cand = np.ascontiguousarray(candidate_items, np.int32)
        hit = self._cached(user_id, n, ignore_items, cand)
        if hit is not None:
            return hit
        ign = None if ignore_items is None else [np.asarray(list(ignore_items), np.int32)]
        items, scores = self._model.recommend([user_id], n, cand, ign)[0]
        return [(int(i), float(s)) for i, s in zip(items, scores)]

    def RetrainUser(self, user_id):
        """WRMF.cs:159-163."""
        self._model.retrain([user_id], by_item=False)
        self._cache = None

    def RetrainItem(self, item_id):
        """WRMF.cs:166-170."""
        self._model.retrain([item_id], by_item=True)
        self._cache = None

    def RecommendMany(self, users, n, ignore_lists=None, candidate_items=None):
        """The all-users loop of ItemRecommendation/Extensions.WritePredictions (:65-128) in one device call."""
        return self._model.recommend(users, n, candidate_items, ignore_lists)

    def SaveModel(self, filename):
        U, V = self._model.get_model()
        with open(filename, "w") as w:
            modelio.write_header(w, self._type_name)
            modelio.write_matrix(w, U)
            modelio.write_matrix(w, V)

    def LoadModel(self, filename):
        with open(filename) as r:
            modelio.read_header(r, self._type_name)
            U, V = modelio.read_matrix(r), modelio.read_matrix(r)
        if U.shape[1] != V.shape[1]:
            raise IOError("Number of user and item factors must match: %d != %d" % (U.shape[1], V.shape[1]))
        self.MaxUserID, self.MaxItemID = U.shape[0] - 1, V.shape[0] - 1
        self.NumFactors = U.shape[1]
        self._new_model(U.shape[0], V.shape[0], np.zeros(0, np.int32), np.zeros(0, np.int32))
        self._model.set_model(U, V)
        self._cache = None

    def ToString(self):
        return "%s num_factors=%d regularization=%s alpha=%s num_iter=%d" % (
            type(self).__name__, self.NumFactors, modelio.fmt(self.Regularization), modelio.fmt(self.Alpha), self.NumIter)

    __str__ = ToString

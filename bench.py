#!/usr/bin/env python
"""Benchmark of the hot path: BiasedMatrixFactorization SGD epochs (ratings/s) on synthetic data of the shapes
BASELINE.json names, with WRMF (config 3) and top-N (config 5) measured in the same run.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ml10m|netflix|...]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

A step is one Iterate() = one full SGD epoch over the resident training ratings. `value` is whole-job ratings/s
with the ratings resident in HBM, timed with CUDA events on the library's stream (max over ranks); `e2e` is the
same metric through the find-iter loop a MyMediaLite user runs (Iterate() + Evaluate(test) with HOST test arrays,
SURVEY.md section 3.2), host<->device copies inside the timed region. N > 1 is weak scaling by default: every rank
brings a user block of the named shape (its own users, the shared item catalogue); item blocks rotate round the ring.
The same line carries, as sub-objects (skipped with --sgd-only):
  strong     the FIXED config-4 problem (100M ratings) sharded over the N ranks (BASELINE.json configs[3] as written)
  zipf0      the same SGD measurement under the pure Zipf(0.8) popularity law (pop_offset = 0) -- the workload as SURVEY 8d
             states it; the default flattens the head (synthetic.POP_OFFSET), see DESIGN.md section 6
  train_e2e  Train() with NumIter = 30 from host arrays: upload, rating-matrix build, strata build and the 30 epochs
  wrmf_c3    WRMF ALS epoch (config 3: 138k x 27k, 20M events, k = 128), rows sharded over the ranks
  topn_c5    top-10 Recommend() for 1M users x 100k items (config 5), users sharded over the ranks
`--impl reference` times the CPU restatement of the reference's own loops (oracle/; the C# original cannot run here: no
Mono/.NET in the image) on the host cores, on the full workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_users, n_items, n_ratings, levels, k, seed, description)
    "ml10m": (71_500, 10_700, 10_000_000, "half", 64, 20260102,
              "BiasedMatrixFactorization k=64 SGD, synthetic MovieLens-10M shape (71.5k x 10.7k, 10M ratings)"),
    "netflix": (480_000, 17_800, 100_000_000, "int", 128, 20260104,
                "BiasedMatrixFactorization k=128 DSGD, synthetic Netflix shape (480k x 17.8k, 100M ratings)"),
    "netflix10": (48_000, 17_800, 10_000_000, "int", 128, 20260104,
                  "BiasedMatrixFactorization k=128 DSGD, 1/10 of the synthetic Netflix shape (48k x 17.8k, 10M ratings)"),
    "nf_sub8": (480_000, 2_225, 12_500_000, "int", 128, 20260104,
                "one GPU-level sub-epoch of the Netflix shape at 8 GPUs (480k users x 1/8 of the items, 12.5M ratings), diagnostic"),
    "tiny": (3_000, 800, 300_000, "half", 64, 7, "debug-sized"),
}
C3 = dict(users=138_000, items=27_000, events=20_000_000, k=128, seed=20260103)
C5 = dict(users=1_000_000, items=100_000, k=128, n=10, ignore=20, seed=20260105)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_facts():
    """Per-launch ncu counters of the dominant kernel for the CURRENT grid (profiles/traffic_r2.json, written from the
    ncu --set full capture of this command by scripts/ncu_summary.py); {} when the capture has not been taken."""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                with open(p) as f:
                    d = json.load(f)
                d["_source"] = "profiles/" + name
                return d
            except Exception:
                pass
    return {}


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason samples during the timed region (profiling recipe's clocks line), read in-process
    through NVML every 20 ms (spawning nvidia-smi takes ~0.3 s per sample and stalls concurrent CUDA driver calls);
    falls back to nvidia-smi if NVML cannot be loaded."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6:
            self.sm.append(float(r[0])); self.mx.append(float(r[1]))
            for c, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if r[2 + c].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.5)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_data(name, rank=0, world=1, pop_offset=None, shard=False):
    """shard = False: the rank's OWN user block of the named shape (weak scaling): n_users users of its own (global id =
    local * world + rank, so that user % world == rank, the reference's block rule) over the shared item catalogue.
    shard = True: the ONE problem of the named shape, of which the rank keeps the users with user % world == rank
    (strong scaling; every rank generates the same arrays from the same seed)."""
    from mymedialite_b200 import synthetic
    nu, ni, n, levels, k, seed, desc = WORKLOADS[name]
    # n ratings in the TRAINING set (the shape the metric is quoted on) + 10 % test ratings on top; generated on the
    # rank's GPU when there is one (the numpy generator needs minutes for 10^8 ratings)
    gen, kw = synthetic.ratings, {}
    try:
        import torch
        if torch.cuda.is_available():
            gen, kw = synthetic.ratings_cuda, {"device": "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))}
    except ImportError:
        pass
    d = gen(nu, ni, int(n / 0.9) + 1024, levels, seed + (0 if shard else 1000 * rank), item_seed=seed, pop_offset=pop_offset, **kw)
    u, i, v = d["train"]
    if u.size > n:
        u, i, v = u[:n], i[:n], v[:n]
    d["train"] = (u, i, v)
    if world > 1 and shard:
        for part in ("train", "test"):
            u, i, v = d[part]
            keep = (u % world) == rank
            d[part] = (u[keep].copy(), i[keep].copy(), v[keep].copy())
        d["n_users"], d["n_items"] = nu, ni
        return d, k, desc
    if world > 1:
        d["train"] = ((d["train"][0] * world + rank).astype(np.int32), d["train"][1], d["train"][2])
        d["test"] = ((d["test"][0] * world + rank).astype(np.int32), d["test"][1], d["test"][2])
    d["n_users"], d["n_items"] = nu * world, ni
    return d, k, desc


class Dist:
    """torch.distributed brackets of the timing (barrier, max / sum over ranks); the data path's collectives are the
    library's own."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.pg = None
        self.uid = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            from mymedialite_b200 import engine
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            t = torch.from_numpy(engine.Context.unique_id() if self.rank == 0 else np.zeros(128, np.uint8)).cuda()
            dist.broadcast(t, 0)
            self.uid = t.cpu().numpy()
            self.pg = dist

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.synchronize()
        if self.pg is not None:
            self.pg.barrier()

    def _reduce(self, x, op):
        if self.pg is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        self.pg.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.pg.ReduceOp.MAX if self.pg else None)

    def sum(self, x):
        return self._reduce(x, self.pg.ReduceOp.SUM if self.pg else None)


def sgd_params(args, k):
    from mymedialite_b200 import engine
    return engine.default_params(biased=1, num_factors=k, num_groups=args.groups, num_subgroups=args.subgroups,
                                 persistent=args.persistent, hot_item_factor=args.hot, hot_copies=args.copies,
                                 intra_block=args.intra, hot_merge_average=args.hot_avg, ctas_per_group=args.cpg)


def time_sgd(args, D, ctx, d, k, steps, warmup, e2e=True):
    """Builds the model on the rank's ratings and times `steps` epochs (device events, L2 flushed before each, max over
    ranks) and, with e2e, `steps` rounds of Iterate() + Evaluate(host test arrays)."""
    from mymedialite_b200 import engine
    import torch
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    n = int(u.size)
    # the e2e leg copies the step's inputs (the test ratings) from PINNED host memory every step
    tu, ti, tv = (torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy() for x in (tu, ti, tv))
    t0 = time.time()
    ratings = engine.DeviceRatings(ctx, u, i, v, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1)
    params = sgd_params(args, k)
    model = engine.SgdModel(ctx, ratings, params)
    model.init_model(1, 0.0, 0.1)
    ctx.synchronize()
    build_s = time.time() - t0
    info = model.strata_info()
    rs = np.random.RandomState(1)      # same stream on every rank: all ranks use the same sub-epoch order

    def seq():
        return rs.permutation(info["G"]).astype(np.int32)

    for _ in range(warmup):
        model.iterate(seq())
    D.barrier(ctx)
    launches0, _ = model.stats()
    ms = []
    for _ in range(steps):
        ctx.flush_l2()
        D.barrier(ctx)
        model.iterate(seq())
        ms.append(D.max(model.stats()[1]))
    D.barrier(ctx)
    launches1, _ = model.stats()
    out = {"n": n, "n_total": int(round(D.sum(n))), "ms": ms, "build_s": build_s, "info": info, "params": params,
           "launches": int(launches1 - launches0), "epochs_run": warmup + steps}
    if e2e:
        model.iterate(seq())                      # one untimed end-to-end step: the first Evaluate() allocates its device scratch
        model.evaluate(tu, ti, tv)
        ctx.synchronize()
        out["epochs_run"] += 1
        D.barrier(ctx)
        t0 = time.time()
        rmse = None
        each = []
        for _ in range(steps):
            t1 = time.time()
            model.iterate(seq())
            rmse = model.evaluate(tu, ti, tv)["RMSE"]
            each.append(round((time.time() - t1) * 1e3, 3))
        ctx.synchronize()
        out["e2e_s"] = D.max(time.time() - t0)
        out["e2e_ms_each"] = each
        out["test_rmse"] = rmse
        out["h2d_bytes"] = int(round(D.sum(12 * tu.size + 4 * info["G"])))
        out["d2h_bytes"] = int(round(D.sum(16 + 8 * 4 * 1184)))
        out["epochs_run"] += steps
    out["train_rmse"] = model.evaluate_train()["RMSE"]
    if not e2e:
        out["test_rmse"] = model.evaluate(tu, ti, tv)["RMSE"]
    model.close(); ratings.close()
    return out


def train_e2e(args, ctx, d, k, num_iter=30):
    """Train() as a host calls it: upload of the host COO arrays, rating-matrix build, strata build, InitModel on the
    device, NumIter epochs, one synchronisation at the end. Wall clock."""
    from mymedialite_b200 import engine
    u, i, v = d["train"]
    ctx.synchronize()
    t0 = time.time()
    ratings = engine.DeviceRatings(ctx, u, i, v, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1)
    t1 = time.time()
    model = engine.SgdModel(ctx, ratings, sgd_params(args, k))
    model.init_model(1, 0.0, 0.1)
    ctx.synchronize()
    t2 = time.time()
    G = model.strata_info()["G"]
    rs = np.random.RandomState(3)
    for _ in range(num_iter):
        model.iterate(rs.permutation(G).astype(np.int32))
    ctx.synchronize()
    t3 = time.time()
    model.close(); ratings.close()
    n = int(u.size)
    return {"num_iter": num_iter, "seconds": round(t3 - t0, 3), "upload_and_matrix_build_s": round(t1 - t0, 3),
            "strata_build_and_init_s": round(t2 - t1, 3), "epochs_s": round(t3 - t2, 3),
            "value": n * num_iter / (t3 - t0), "unit": "ratings/s",
            "what": "Train(): host COO arrays -> device, rating-matrix and strata build, InitModel, %d epochs; wall clock" % num_iter}


def measure_tf32_peak():
    """Dense TF32 GEMM throughput of this GPU, measured the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3,
    best of 10): the denominator of the tensor-pipe fractions of the TF32 kernels (BASELINE.md section 2 asks for it)."""
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def bench_wrmf_c3(args, D, ctx, tf32_peak):
    """Config 3: WRMF.Iterate() epochs; rows of each half-sweep sharded over the ranks (strong scaling)."""
    from mymedialite_b200 import engine, synthetic
    c = C3
    u, i = synthetic.implicit_cuda(c["users"], c["items"], c["events"], c["seed"], device="cuda:%d" % D.local)
    f = engine.DeviceFeedback(ctx, u, i, max_user=c["users"] - 1, max_item=c["items"] - 1)
    m = engine.WrmfModel(ctx, f, c["k"])
    m.init_model(1)
    ctx.synchronize()
    epochs = max(args.wrmf_epochs, 2)
    ms = []
    for e in range(epochs + 1):
        D.barrier(ctx)
        m.iterate()
        x = D.max(m.stats()[1])
        if e > 0:          # the first epoch warms the workspace allocations
            ms.append(x)
    nnz = int(f.nnz)
    flop = 2.0 * c["k"] * c["k"] * (2.0 * nnz + c["users"] + c["items"])
    epoch_ms = float(np.median(ms))
    out = {"epoch_ms": round(epoch_ms, 3), "epoch_s": round(epoch_ms * 1e-3, 5), "epoch_ms_each": [round(x, 2) for x in ms],
           "n_gpus": D.world, "scaling": "strong", "shape": dict(users=c["users"], items=c["items"], nnz=nnz, k=c["k"]),
           "algorithmic_tflops": flop / (epoch_ms * 1e-3) / 1e12,
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flop / (epoch_ms * 1e-3) / 1e12 / D.world,
                        "peak": tf32_peak / 3.0, "peak_kind": "TF32 dense GEMM measured in this run (%.0f TF/s) / 3: the Gram sums "
                        "run as 3 TF32 MMAs per fp32 product (hi*hi + hi*lo + lo*hi)" % tf32_peak,
                        "frac": flop / (epoch_ms * 1e-3) / 1e12 / D.world / (tf32_peak / 3.0),
                        "flop_per_epoch": flop, "note": "whole epoch (Gram sums + solves + gathers) over the tensor-work flop count of SURVEY 8d"}}
    if D.rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        rows = 2048
        uptr, ucols = O.feedback_csr(u, i, c["users"] - 1)
        rs = np.random.RandomState(5)
        U = (0.1 * rs.randn(c["users"], c["k"])).astype(np.float32); V = (0.1 * rs.randn(c["items"], c["k"])).astype(np.float32)
        Us = U[:rows].copy()
        t0 = time.time()
        O.wrmf_optimize(uptr[:rows + 1], ucols, Us, V, omp_threads=cores)
        dt = time.time() - t0
        ev = int(uptr[rows])
        # one epoch = both half-sweeps over all rows: extrapolated by events (the k^2 work per event dominates at k = 128)
        est = dt * (2.0 * nnz) / max(ev, 1)
        out["cpu_baseline"] = {"value": est, "unit": "s/epoch (extrapolated)", "cores": cores, "kind": "port",
                               "seconds": round(dt, 2),
                               "sample": "oracle WRMF.Optimize row loop (WRMF.cs:110-156 restated, OpenMP over rows) on the first %d "
                                         "user rows (%d events) incl. one Gram matrix of V; epoch estimate = time x 2 nnz / events" % (rows, ev)}
    m.close(); f.close()
    return out


def bench_topn_c5(args, D, ctx, tf32_peak):
    """Config 5: top-10 for all users; users sharded over the ranks (contiguous ranges), V replicated, no collective."""
    from mymedialite_b200 import engine
    c = C5
    total = c["users"]
    rs = np.random.default_rng(c["seed"])
    V = rs.standard_normal((c["items"], c["k"]), dtype=np.float32) * np.float32(0.1)
    lo, hi = D.rank * total // D.world, (D.rank + 1) * total // D.world
    nu = hi - lo
    rs = np.random.default_rng(c["seed"] + 1 + D.rank)
    U = rs.standard_normal((nu, c["k"]), dtype=np.float32) * np.float32(0.1)
    users = np.arange(nu, dtype=np.int32)
    ign_idx = np.ascontiguousarray(rs.integers(0, c["items"], (nu, c["ignore"]), dtype=np.int32)).reshape(-1)
    ign_ptr = np.arange(nu + 1, dtype=np.int64) * c["ignore"]
    n_out = c["n"]
    # Recommend() on a device-resident WRMF model (CudaWRMF.Recommend for all users = mml_wrmf_recommend): the factors are set
    # once, outside the timed call; user ids and ignore lists go in and the lists come out of HOST arrays inside it
    fb = engine.DeviceFeedback(ctx, np.zeros(0, np.int32), np.zeros(0, np.int32), max_user=nu - 1, max_item=c["items"] - 1)
    wm = engine.WrmfModel(ctx, fb, c["k"])
    wm.set_model(U, V)
    oi = np.zeros((nu, n_out), np.int32); os_ = np.zeros((nu, n_out), np.float32); oc = np.zeros(nu, np.int32)
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    calls = []
    for r in range(max(args.topn_reps, 2) + 1):
        D.barrier(ctx)
        t0 = time.time()
        engine.check(ctx.lib.mml_wrmf_recommend(wm.h, users, nu, c["n"], None, c["items"], ign_ptr, ign_idx, oi, os_, oc))
        wall = D.max(time.time() - t0)
        st = engine.topn_last_stats()
        if r > 0:
            calls.append((wall * 1e3, D.max(st["tensor_path_ms"]), st))
    wm.close(); fb.close()
    call_ms = float(np.median([x[0] for x in calls])); path_ms = float(np.median([x[1] for x in calls]))
    flop = 2.0 * total * c["items"] * c["k"]
    out = {"call_ms": round(call_ms, 2), "device_path_ms": round(path_ms, 2), "call_ms_each": [round(x[0], 1) for x in calls],
           "users_per_s": total / (call_ms * 1e-3), "n_gpus": D.world, "scaling": "strong",
           "shape": dict(users=total, items=c["items"], k=c["k"], n=c["n"], ignore_per_user=c["ignore"]),
           "users_exact_path": int(calls[-1][2]["users_exact_path"]),
           "h2d_bytes_per_call": int(round(D.sum(users.nbytes + ign_idx.nbytes + ign_ptr.nbytes))),
           "d2h_bytes_per_call": int(round(D.sum(oi.nbytes + os_.nbytes + oc.nbytes))),
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flop / (call_ms * 1e-3) / 1e12 / D.world, "peak": tf32_peak,
                        "peak_kind": "TF32 dense GEMM measured in this run", "frac": flop / (call_ms * 1e-3) / 1e12 / D.world / tf32_peak,
                        "flop_per_call": flop, "note": "whole mml_wrmf_recommend call on the device-resident model: user ids and ignore lists up from host arrays, scoring "
                        "GEMM + fused top-k, exact finish, lists down into host arrays"}}
    if D.rank == 0 and not args.no_cpu:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sample = list(range(0, min(nu, 4 * cores)))

        def one(b):
            return O.recommend_mf(U, V, b, c["n"], None, np.sort(ign_idx[b * c["ignore"]:(b + 1) * c["ignore"]]))
        t0 = time.time()
        with ThreadPoolExecutor(cores) as ex:
            res = list(ex.map(one, sample))
        dt = time.time() - t0
        same = all(np.array_equal(res[t][0], oi[b, :oc[b]]) and
                   np.array_equal(np.asarray(res[t][1], np.float32).view(np.uint32), os_[b, :oc[b]].view(np.uint32))
                   for t, b in enumerate(sample))
        out["cpu_baseline"] = {"value": len(sample) / dt, "unit": "users/s", "cores": cores, "kind": "port", "seconds": round(dt, 2),
                               "sample": "oracle Recommend() (Recommender.cs:52-103 restated) for the first %d users of rank 0, one host "
                                         "thread per user" % len(sample)}
        out["bit_identical_to_oracle_on_sample"] = bool(same)
    return out


def run_ours(args):
    from mymedialite_b200 import engine
    D = Dist()
    world, rank = D.world, D.rank
    strong = args.scaling == "strong"
    d, k, desc = make_data(args.workload, rank, world, args.pop_offset, shard=strong)
    ctx = engine.Context(D.local, rank, world, D.uid)

    sampler = ClockSampler(D.local)
    sampler.start()
    r = time_sgd(args, D, ctx, d, k, args.steps, args.warmup, e2e=True)
    sampler.stop_flag.set(); sampler.join()
    n, n_total, ms, info = r["n"], r["n_total"], r["ms"], r["info"]
    dev_ms = float(np.sum(ms))
    pk, pk_kind = peaks()
    bytes_per_rating = 16 * k + 28
    ms_per_step = dev_ms / args.steps
    value = n_total * args.steps / (dev_ms * 1e-3)
    achieved = bytes_per_rating * (n_total / world) / (ms_per_step * 1e-3) / 1e9      # per GPU
    facts = ncu_facts().get(args.workload) if world == 1 else None
    if isinstance(facts, (int, float)):      # round-1 file: the DRAM bytes only
        facts = {"dram_bytes_per_launch": facts}
    facts = facts or {}
    sm_mhz = sampler.summary()["sm_mhz"] or 1965.0
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "traffic": facts.get("dram_bytes_per_launch"), "peak_kind": pk_kind,
            "bytes_per_rating": bytes_per_rating, "kernel": "sgd_epoch_kernel (one launch per epoch and item block)",
            "note": "algorithmic bytes of SURVEY 8d (both rows charged for every rating); the kernel keeps user rows in registers across "
                    "a run and the item matrix in the L2, so this fraction can exceed 1 -- the limits that bind are below"}
    if facts.get("warp_inst_per_launch"):
        wipr = facts["warp_inst_per_launch"] / facts.get("ratings_per_launch", n)
        issue_peak = 148 * 4 * sm_mhz * 1e6          # warp instructions per second the 4 schedulers of every SM can issue
        roof["binding"] = {
            "issue": {"warp_inst_per_rating": wipr, "achieved_ginst_s": wipr * (n_total / world) / (ms_per_step * 1e-3) / 1e9,
                      "peak_ginst_s": issue_peak / 1e9, "frac": wipr * (n_total / world) / (ms_per_step * 1e-3) / issue_peak,
                      "what": "issue slots: ncu smsp__inst_executed.sum per launch / ratings x measured ratings/s over 148 SMs x 4 schedulers x SM clock"},
            "l2_atomics": {"red_sectors_per_rating": facts.get("l2_red_sectors_per_launch", 0) / facts.get("ratings_per_launch", n),
                           "lts_throughput_pct_ncu": facts.get("lts_throughput_pct"),
                           "what": "item rows are updated by red.global.add.v4.f32 executed in the L2 slices"},
            "source": facts.get("source", ncu_facts().get("_source"))}
    # the limit that binds, measured live: the rate at which THIS GPU's L2 serves one row read (L1 bypassed) plus one row of
    # red.global.add.v4.f32 per touch on a table the size of the item matrix -- the epoch kernel's item-row traffic per rating
    # without its arithmetic (mml_ctx_probe_l2). User rows, entries and biases come on top, so 1.0 is not reachable.
    try:
        n_items_tab = int(d["n_items"])
        rows_rr = ctx.probe_l2(2, n_items_tab, 32 * ((k + 31) // 32), 3)
        rows_rd = ctx.probe_l2(0, n_items_tab, 32 * ((k + 31) // 32), 3)
        rows_red = ctx.probe_l2(1, n_items_tab, 32 * ((k + 31) // 32), 3)
        per_gpu = (n_total / world) / (ms_per_step * 1e-3)
        roof.setdefault("binding", {})["l2"] = {
            "achieved_rows_s": per_gpu, "peak_rows_s": rows_rr, "frac": per_gpu / rows_rr,
            "peak_read_rows_s": rows_rd, "peak_red_rows_s": rows_red, "row_bytes": 4 * 32 * ((k + 31) // 32),
            "what": "item-row traffic of the epoch kernel (per rating: one row read + one row of vector atomics, executed as "
                    "read-modify-writes in the L2 slices) against mml_ctx_probe_l2 mode 2 on a %d-row table, measured in this run; "
                    "at N > 1 the achieved rate includes the ring exchange" % n_items_tab}
    except Exception as ex:      # a diagnostic: never fails the bench line
        roof.setdefault("binding", {})["l2"] = {"error": str(ex)}
    out = {
        "metric": "BiasedMF SGD ratings/sec", "value": value, "unit": "ratings/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + ((" per GPU (users sharded, item catalogue shared)" if not strong else
                                        ", ONE problem sharded by user over the GPUs") if world > 1 else ""),
                   "n_ratings": n_total, "num_factors": k,
                   "schedule": "DSGD G=%d x %d CTAs, W=%d %s%s%s" % (info["G"], info["cpg"], info["W"], "async" if args.intra else "rounds",
                                                          " persistent" if r["params"].persistent != 0 else "",
                                                          (", item-block ring over %d GPUs" % world) if world > 1 else ""),
                   "l2": "flushed before every timed epoch (384 MB memset)", "strata_build_s": round(r["build_s"], 3),
                   "item_popularity": "Zipf-Mandelbrot (i + %g)^-0.8" % (30.0 if args.pop_offset is None else args.pop_offset)},
        "e2e": {"value": n_total * args.steps / r["e2e_s"], "unit": "ratings/s",
                "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": r["d2h_bytes"],
                "ms_each_rank0": r["e2e_ms_each"],
                "what": "Iterate() + Evaluate(test) per step through the C ABI, test ratings in pinned host memory; bytes summed over ranks"},
        "gpu_launches": r["launches"],
        "roofline": roof,
        "clocks": sampler.summary(),
        "rmse": {"train": r["train_rmse"], "test": r["test_rmse"], "epochs": r["epochs_run"]},
        "ms_each": [round(x, 3) for x in ms],
    }
    if not args.no_cpu and world == 1:      # rank 0 at N = 1 only: the check model must not enter the ring collectives alone
        out["cpu_baseline"], out["rmse_vs_ref"] = cpu_baseline(args, d, k, ctx)
    if not args.sgd_only:
        if world == 1 and not strong:
            out["train_e2e"] = train_e2e(args, ctx, d, k)
        del d
        if args.workload == "netflix" and args.pop_offset is None:
            # the pure Zipf(0.8) law of SURVEY 8d beside the default (flattened head)
            dz, _, _ = make_data(args.workload, rank, world, 0.0, shard=strong)
            rz = time_sgd(args, D, ctx, dz, k, max(args.steps // 2, 2), 3, e2e=False)
            zms = float(np.mean(rz["ms"]))
            out["zipf0"] = {"value": rz["n_total"] / (zms * 1e-3), "unit": "ratings/s", "ms_per_step": zms, "pop_offset": 0.0,
                            "rmse": {"train": rz["train_rmse"], "test": rz["test_rmse"], "epochs": rz["epochs_run"]},
                            "what": "same measurement with item popularity i^-0.8 (top item 2.8 % of all ratings)"}
            del dz
        if world > 1 and not strong and args.workload == "netflix":
            ds, _, _ = make_data(args.workload, rank, world, args.pop_offset, shard=True)
            rst = time_sgd(args, D, ctx, ds, k, args.steps, 3, e2e=False)
            sms = float(np.mean(rst["ms"]))
            out["strong"] = {"value": rst["n_total"] / (sms * 1e-3), "unit": "ratings/s", "ms_per_step": sms, "n_ratings": rst["n_total"],
                             "scaling": "strong", "rmse": {"train": rst["train_rmse"], "test": rst["test_rmse"], "epochs": rst["epochs_run"]},
                             "what": "the fixed 100M-rating problem, users sharded over %d GPUs (user %% N), item blocks round the ring" % world}
            del ds
        tf32 = measure_tf32_peak()
        out["tf32_tflops_measured"] = tf32
        out["wrmf_c3"] = bench_wrmf_c3(args, D, ctx, tf32)
        out["topn_c5"] = bench_topn_c5(args, D, ctx, tf32)
    if rank == 0:
        print(json.dumps(out), flush=True)
    D.barrier()


def cpu_baseline(args, d, k, ctx=None):
    """The oracle's restatement of BiasedMatrixFactorization.Iterate timed on this box's host cores:
    single-threaded (MaxThreads=1) over a bounded prefix of the workload. The same prefix, from the same initial
    factors, then runs one epoch of the GPU's default schedule: "test RMSE vs ref" of the metric (gate 0.5 %);
    the 10-epoch per-epoch gate at 10M ratings is tests/test_rmse_gate_gpu.py."""
    from oracle import oracle as O
    u, i, v = d["train"]
    m = min(u.size, args.cpu_sample)
    us, is_, vs = u[:m].copy(), i[:m].copy(), v[:m].copy()
    om = O.Model(us, is_, vs, biased=True, num_factors=k, max_user=int(us.max()), max_item=d["n_items"] - 1)
    rng = O.Random(1)
    om.init(rng)
    U0, V0 = om.user_factors.copy(), om.item_factors.copy()
    t0 = time.time()
    om.iterate(rng)
    dt = time.time() - t0
    base = {"value": m / dt, "unit": "ratings/s", "cores": 1, "kind": "port",
            "sample": "one single-threaded epoch (MaxThreads=1 order) over the first %d ratings of the workload; "
                      "C restatement of the reference loop (no Mono/.NET in this image)" % m,
            "seconds": round(dt, 2)}
    check = None
    if ctx is not None:
        from mymedialite_b200 import engine
        tu, ti, tv = d["test"]
        mt = min(tu.size, max(m // 9, 1))
        tu, ti, tv = tu[:mt], ti[:mt], tv[:mt]
        r = engine.DeviceRatings(ctx, us, is_, vs, max_user=int(us.max()), max_item=d["n_items"] - 1)
        gm = engine.SgdModel(ctx, r, engine.default_params(biased=1, num_factors=k, num_subgroups=args.subgroups))
        gm.set_model(U0, V0)
        gm.iterate(np.random.RandomState(2).permutation(gm.strata_info()["G"]).astype(np.int32))
        g_tr, g_te = gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]
        o_tr, o_te = om.evaluate(us, is_, vs)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]
        check = {"gpu": {"train": g_tr, "test": g_te}, "ref": {"train": o_tr, "test": o_te},
                 "rel_diff": {"train": abs(g_tr - o_tr) / o_tr, "test": abs(g_te - o_te) / o_te}, "gate": 0.005,
                 "sample": "one epoch from identical initial factors on the cpu_baseline prefix (%d ratings), %d test ratings; "
                           "GPU: default parallel schedule, ref: oracle in the reference's single-threaded order" % (m, mt)}
        gm.close(); r.close()
    return base, check


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port) with all host threads: DSGD blocks on
    OpenMP threads, as BiasedMatrixFactorization does with MaxThreads = cores. Runs the WHOLE workload of the other arm
    (same config); --cpu-sample bounds it only when given explicitly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    d, k, desc = make_data(args.workload, pop_offset=args.pop_offset)
    u, i, v = d["train"]
    cores = os.cpu_count() or 1
    m = u.size if args.cpu_sample <= 0 else min(u.size, args.cpu_sample)
    us, is_, vs = (u, i, v) if m == u.size else (u[:m].copy(), i[:m].copy(), v[:m].copy())
    om = O.Model(us, is_, vs, biased=True, num_factors=k, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1,
                 max_threads=cores, omp_threads=cores)
    rng = O.Random(1)
    t0 = time.time()
    om.init(rng)          # InitModel + PartitionUsersAndItems (MultiCore.cs:43-73): not timed, as the strata build is not on the other arm
    setup_s = time.time() - t0
    for _ in range(args.warmup):
        om.iterate(rng)
    t0 = time.time()
    for _ in range(args.steps):
        om.iterate(rng)
    dt = time.time() - t0
    value = m * args.steps / dt
    sample = ("%s of the workload per step (%d ratings), DSGD block schedule (MultiCore.cs:43-73) on %d OpenMP threads; "
              "C restatement of the reference loop (no Mono/.NET in this image)" % ("all" if m == u.size else "a prefix", m, cores))
    out = {
        "impl": "reference", "metric": "BiasedMF SGD ratings/sec", "value": value, "unit": "ratings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "n_ratings": int(m), "num_factors": k, "setup_s": round(setup_s, 2)},
        "cpu_baseline": {"value": value, "unit": "ratings/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ratings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = a user block of the named shape per GPU (default); strong = the one named problem sharded")
    ap.add_argument("--sgd-only", action="store_true", help="skip the zipf0 / strong / train_e2e / wrmf_c3 / topn_c5 sub-objects")
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--cpg", type=int, default=0, help="CTAs per worker group (async mode); groups default to SMs / cpg")
    ap.add_argument("--subgroups", type=int, default=0, help="warps per CTA; 0 = the library's default for the row length and GPU count")
    ap.add_argument("--persistent", type=int, default=-1)
    ap.add_argument("--hot", type=float, default=0.0)
    ap.add_argument("--copies", type=int, default=0)
    ap.add_argument("--hot-avg", type=int, default=1)
    ap.add_argument("--intra", type=int, default=1, help="0 = conflict-free rounds, 1 = async (default)")
    ap.add_argument("--cpu-sample", type=int, default=-1,
                    help="ratings of the cpu_baseline / reference-arm sample; default: 10M (cpu_baseline), the whole workload (reference arm)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--wrmf-epochs", type=int, default=3)
    ap.add_argument("--topn-reps", type=int, default=3)
    ap.add_argument("--pop-offset", type=float, default=None,
                    help="item popularity (i + offset)^-0.8; default synthetic.POP_OFFSET = 30, 0 = the pure Zipf law")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.cpu_sample <= 0:
            args.cpu_sample = 10_000_000
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: BiasedMatrixFactorization SGD epochs (ratings/s) on synthetic data of the shapes
BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ml10m|netflix|...]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

A step is one Iterate() = one full SGD epoch over the resident training ratings. `value` is whole-job ratings/s
with the ratings resident in HBM, timed with CUDA events on the library's stream (max over ranks); `e2e` is the
same metric through the find-iter loop a MyMediaLite user runs (Iterate() + Evaluate(test) with HOST test arrays,
SURVEY.md section 3.2), host<->device copies inside the timed region. N > 1 is weak scaling: every rank brings a
user block of the named shape (its own users, the shared item catalogue); item blocks rotate round the NCCL ring.
`--impl reference` times the CPU restatement of the reference's own loops (oracle/; the C# original cannot run
here: no Mono/.NET in the image) on the host cores.
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


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


def make_data(name, rank=0, world=1, pop_offset=None):
    """The rank's user shard: n_users users of its own (global id = local * world + rank, so that
    user % world == rank, the reference's block rule) over the shared item catalogue."""
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
    d = gen(nu, ni, int(n / 0.9) + 1024, levels, seed + 1000 * rank, item_seed=seed, pop_offset=pop_offset, **kw)
    u, i, v = d["train"]
    if u.size > n:
        u, i, v = u[:n], i[:n], v[:n]
    d["train"] = (u, i, v)
    if world > 1:
        d["train"] = ((d["train"][0] * world + rank).astype(np.int32), d["train"][1], d["train"][2])
        d["test"] = ((d["test"][0] * world + rank).astype(np.int32), d["test"][1], d["test"][2])
    d["n_users"], d["n_items"] = nu * world, ni
    return d, k, desc


def run_ours(args):
    from mymedialite_b200 import engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        t = torch.from_numpy(engine.Context.unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d, k, desc = make_data(args.workload, rank, world, args.pop_offset)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    n = int(u.size)
    ctx = engine.Context(local_rank, rank, world, uid)
    # the e2e leg copies the step's inputs (the test ratings) from PINNED host memory every step
    import torch
    tu, ti, tv = (torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy() for x in (tu, ti, tv))
    t0 = time.time()
    ratings = engine.DeviceRatings(ctx, u, i, v, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1)
    params = engine.default_params(biased=1, num_factors=k, num_groups=args.groups, num_subgroups=args.subgroups,
                                   persistent=args.persistent, hot_item_factor=args.hot, hot_copies=args.copies,
                                   intra_block=args.intra, hot_merge_average=args.hot_avg, ctas_per_group=args.cpg)
    model = engine.SgdModel(ctx, ratings, params)
    model.init_model(1, 0.0, 0.1)
    ctx.synchronize()
    build_s = time.time() - t0
    info = model.strata_info()
    rs = np.random.RandomState(1)      # same stream on every rank: all ranks use the same sub-epoch order

    def seq():
        return rs.permutation(info["G"]).astype(np.int32)

    for _ in range(args.warmup):
        model.iterate(seq())
    barrier()
    launches0, _ = model.stats()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # --- device-resident: K epochs, each timed by CUDA events on the library's stream (kernels + ring exchange),
    #     L2 flushed before each
    ms = []
    for _ in range(args.steps):
        ctx.flush_l2()
        barrier()
        model.iterate(seq())
        ms.append(max_over_ranks(model.stats()[1]))
    barrier()
    launches1, _ = model.stats()
    dev_ms = float(np.sum(ms))
    # --- end to end: the find-iter loop (Iterate + Evaluate on host test arrays), wall clock around synchronous calls
    barrier()
    t0 = time.time()
    rmse = None
    for _ in range(args.steps):
        model.iterate(seq())
        rmse = model.evaluate(tu, ti, tv)["RMSE"]
    ctx.synchronize()
    e2e_s = max_over_ranks(time.time() - t0)
    sampler.stop_flag.set(); sampler.join()
    train_rmse = model.evaluate_train()["RMSE"]

    n_total = n * world
    if dist is not None:
        import torch
        t = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        n_total = int(t.item())
    if rank != 0:
        if dist is not None:
            dist.barrier()
        return
    pk, pk_kind = peaks()
    bytes_per_rating = 16 * k + 28
    ms_per_step = dev_ms / args.steps
    value = n_total * args.steps / (dev_ms * 1e-3)
    achieved = bytes_per_rating * n / (ms_per_step * 1e-3) / 1e9      # per GPU
    out = {
        "metric": "BiasedMF SGD ratings/sec", "value": value, "unit": "ratings/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + (" per GPU (users sharded, item catalogue shared)" if world > 1 else ""),
                   "n_ratings": n_total, "num_factors": k,
                   "schedule": "DSGD G=%d x %d CTAs, W=%d %s%s%s" % (info["G"], info["cpg"], info["W"], "async" if args.intra else "rounds",
                                                          " persistent" if params.persistent != 0 else "",
                                                          (", item-block ring over %d GPUs" % world) if world > 1 else ""),
                   "l2": "flushed before every timed epoch (384 MB memset)", "strata_build_s": round(build_s, 3)},
        "e2e": {"value": n_total * args.steps / e2e_s, "unit": "ratings/s",
                "h2d_bytes_per_step": int(12 * tu.size + 4 * info["G"]), "d2h_bytes_per_step": 16 + 8 * 4 * 1184,
                "what": "Iterate() + Evaluate(test) per step through the C ABI, test ratings in pinned host memory"},
        "gpu_launches": int(launches1 - launches0),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_kind": pk_kind,
                     "bytes_per_rating": bytes_per_rating, "kernel": "sgd_epoch_kernel (one launch per epoch and item block)"},
        "clocks": sampler.summary(),
        "rmse": {"train": train_rmse, "test": rmse, "epochs": args.warmup + 2 * args.steps},
        "ms_each": [round(x, 3) for x in ms],
    }
    tr = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tr):
        try:
            with open(tr) as f:
                out["roofline"]["traffic"] = json.load(f).get(args.workload)
        except Exception:
            pass
    if not args.no_cpu and world == 1:      # rank 0 at N = 1 only: the check model must not enter the ring collectives alone
        out["cpu_baseline"], out["rmse_vs_ref"] = cpu_baseline(args, d, k, ctx)
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()


def cpu_baseline(args, d, k, ctx=None):
    """The oracle's restatement of BiasedMatrixFactorization.Iterate timed on this box's host cores:
    single-threaded (MaxThreads=1) over a bounded prefix of the workload. The same prefix, from the same initial
    factors, then runs one epoch of the GPU's default schedule: "test RMSE vs ref" of the metric (gate 0.5 %)."""
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
    OpenMP threads, as BiasedMatrixFactorization does with MaxThreads = cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    d, k, desc = make_data(args.workload, pop_offset=args.pop_offset)
    u, i, v = d["train"]
    cores = os.cpu_count() or 1
    m = min(u.size, args.cpu_sample)
    us, is_, vs = u[:m].copy(), i[:m].copy(), v[:m].copy()
    om = O.Model(us, is_, vs, biased=True, num_factors=k, max_user=int(us.max()), max_item=d["n_items"] - 1,
                 max_threads=cores, omp_threads=cores)
    rng = O.Random(1)
    om.init(rng)
    for _ in range(args.warmup):
        om.iterate(rng)
    t0 = time.time()
    for _ in range(args.steps):
        om.iterate(rng)
    dt = time.time() - t0
    value = m * args.steps / dt
    sample = ("%d-rating prefix of the workload per step, DSGD block schedule (MultiCore.cs:43-73) on %d OpenMP threads; "
              "C restatement of the reference loop (no Mono/.NET in this image)" % (m, cores))
    out = {
        "impl": "reference", "metric": "BiasedMF SGD ratings/sec", "value": value, "unit": "ratings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "n_ratings": int(m), "num_factors": k},
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
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--cpg", type=int, default=0, help="CTAs per worker group (async mode); groups default to SMs / cpg")
    ap.add_argument("--subgroups", type=int, default=16)
    ap.add_argument("--persistent", type=int, default=-1)
    ap.add_argument("--hot", type=float, default=0.0)
    ap.add_argument("--copies", type=int, default=0)
    ap.add_argument("--hot-avg", type=int, default=1)
    ap.add_argument("--intra", type=int, default=1, help="0 = conflict-free rounds, 1 = async (default)")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pop-offset", type=float, default=None,
                    help="item popularity (i + offset)^-0.8; default synthetic.POP_OFFSET = 30, 0 = the pure Zipf law")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.cpu_sample == 10_000_000:
            args.cpu_sample = 4_000_000
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Deterministic synthetic data of the shapes BASELINE.json names.

Planted model: rank-16 factors ~ N(0, 0.35^2), biases ~ N(0, 0.3^2), mu = 3.6, noise N(0, 0.5^2); user activity
is log-normal, item popularity Zipf-Mandelbrot p_i ~ (i + POP_OFFSET)^-0.8; (user, item) pairs are distinct; ids are
dense (every user and item appears at least once when n is large enough). The same bytes feed the GPU engine and the
CPU oracle.

POP_OFFSET = 30 flattens the head of the pure Zipf(0.8) law SURVEY.md section 8d names: pure Zipf gives the most
popular of 17.8k items 2.8 % of all ratings, twelve times the share of the real data sets' top titles (Netflix
0.23 %, MovieLens-10M 0.35 %); with the offset it is 0.25 %. It matters because SGD on one item row is a sequential
chain: a 2.8 % item alone bounds the epoch time of ANY schedule that keeps the reference's block exclusivity
(pop_offset=0 reproduces the pure law; DESIGN.md section 6 has both numbers)."""
import os

import numpy as np

SHAPES = {
    # name: (n_users, n_items, n_ratings, levels, k, seed)
    "ml10m": (71_500, 10_700, 10_000_000, "half", 64, 20260102),
    "ml20m_implicit": (138_000, 27_000, 20_000_000, None, 128, 20260103),
    "netflix": (480_000, 17_800, 100_000_000, "int", 128, 20260104),
}


POP_OFFSET = 30.0


def _pairs(rng, n_users, n_items, n, item_rng=None, pop_offset=None):
    act = rng.lognormal(0.0, 1.0, n_users)
    pop = 1.0 / (np.arange(1, n_items + 1) + (POP_OFFSET if pop_offset is None else pop_offset)) ** 0.8
    pop = pop[(item_rng or rng).permutation(n_items)]
    pop /= pop.sum()
    cdf = np.cumsum(pop)
    cdf[-1] = 1.0
    users_out, items_out = [], []
    have = 0
    want = n
    # guarantee every user and item once, then fill by the activity / popularity laws and de-duplicate
    base_u = np.concatenate([np.arange(n_users, dtype=np.int64), rng.integers(0, n_users, n_items)])
    base_i = np.concatenate([rng.integers(0, n_items, n_users), np.arange(n_items, dtype=np.int64)])
    keys = np.unique(base_u * n_items + base_i)
    while True:
        need = want - keys.size
        if need <= 0:
            break
        draw = int(need * 1.15) + 1024
        cnt = rng.multinomial(draw, act / act.sum())
        u = np.repeat(np.arange(n_users, dtype=np.int64), cnt)
        i = np.searchsorted(cdf, rng.random(u.size), side="right").astype(np.int64)
        np.minimum(i, n_items - 1, out=i)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
    if keys.size > want:
        # drop random surplus pairs, never one of the "every id once" pairs' ids entirely: surplus is tiny (<15 %)
        drop = rng.choice(keys.size, keys.size - want, replace=False)
        mask = np.ones(keys.size, bool)
        mask[drop] = False
        keys = keys[mask]
    perm = rng.permutation(keys.size)
    keys = keys[perm]
    return (keys // n_items).astype(np.int32), (keys % n_items).astype(np.int32)


def ratings(n_users, n_items, n, levels="half", seed=1, test_fraction=0.1, item_seed=None, pop_offset=None):
    """Returns dict(train=(u, i, v), test=(u, i, v), n_users, n_items). item_seed: the item side of the planted
    model and the popularity law come from their own generator, so that several user shards (one per GPU, each
    with its own `seed`) share one item catalogue."""
    rng = np.random.Generator(np.random.PCG64(seed))
    irng = np.random.Generator(np.random.PCG64(item_seed)) if item_seed is not None else None
    u, i = _pairs(rng, n_users, n_items, n, irng, pop_offset)
    rank = 16
    Pu = rng.normal(0, 0.35, (n_users, rank)).astype(np.float32)
    Qi = (irng or rng).normal(0, 0.35, (n_items, rank)).astype(np.float32)
    bu = rng.normal(0, 0.3, n_users).astype(np.float32)
    bi = (irng or rng).normal(0, 0.3, n_items).astype(np.float32)
    v = np.empty(u.size, np.float32)
    step = 1 << 22
    for s in range(0, u.size, step):
        uu, ii = u[s:s + step], i[s:s + step]
        v[s:s + step] = 3.6 + bu[uu] + bi[ii] + np.einsum("nk,nk->n", Pu[uu], Qi[ii]) + rng.normal(0, 0.5, uu.size)
    if levels == "half":
        v = np.clip(np.round(v * 2) / 2, 0.5, 5.0).astype(np.float32)
    elif levels == "int":
        v = np.clip(np.round(v), 1.0, 5.0).astype(np.float32)
    # 90/10 split by a hash of (u, i, seed)
    h = (u.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) ^ i.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
         ^ np.uint64(seed)) >> np.uint64(33)
    is_test = (h % np.uint64(1000)) < np.uint64(int(test_fraction * 1000))
    tr, te = ~is_test, is_test
    return dict(train=(u[tr].copy(), i[tr].copy(), v[tr].copy()), test=(u[te].copy(), i[te].copy(), v[te].copy()),
                n_users=n_users, n_items=n_items)


def ratings_cuda(n_users, n_items, n, levels="half", seed=1, test_fraction=0.1, item_seed=None, device="cuda", pop_offset=None):
    """Same planted model and laws as ratings(), generated with torch on the GPU (the 10^8-rating shapes take minutes in
    numpy, seconds here). Deterministic for a given seed on a given torch build; NOT the same stream as ratings().
    Benchmark plumbing only: the arrays it returns feed the engine and the CPU oracle alike."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev); g.manual_seed(int(seed))
    gi = torch.Generator(device=dev); gi.manual_seed(int(item_seed if item_seed is not None else seed) + 7919)
    act = torch.exp(torch.randn(n_users, generator=g, device=dev, dtype=torch.float64))
    if os.environ.get("MMLB200_SYN_ACT_CLIP"):      # diagnostic: cap the user activity law (how much of an epoch is the heaviest user's chain?)
        act = torch.clamp(act, max=float(os.environ["MMLB200_SYN_ACT_CLIP"]))
    pop = 1.0 / (torch.arange(1, n_items + 1, device=dev, dtype=torch.float64) + (POP_OFFSET if pop_offset is None else pop_offset)) ** 0.8
    pop = pop[torch.randperm(n_items, generator=gi, device=dev)]
    cdf = torch.cumsum(pop / pop.sum(), 0); cdf[-1] = 1.0
    ucdf = torch.cumsum(act / act.sum(), 0); ucdf[-1] = 1.0
    # every user and item once, then draws by the activity / popularity laws, de-duplicated
    ar_u = torch.arange(n_users, device=dev, dtype=torch.int64); ar_i = torch.arange(n_items, device=dev, dtype=torch.int64)
    keys = torch.unique(torch.cat([ar_u * n_items + torch.randint(0, n_items, (n_users,), generator=g, device=dev),
                                   torch.randint(0, n_users, (n_items,), generator=g, device=dev) * n_items + ar_i]))
    while keys.numel() < n:
        draw = int((n - keys.numel()) * 1.15) + 1024
        uu = torch.searchsorted(ucdf, torch.rand(draw, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=n_users - 1)
        ii = torch.searchsorted(cdf, torch.rand(draw, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, uu * n_items + ii]))
        del uu, ii
    if keys.numel() > n:
        keep = torch.randperm(keys.numel(), generator=g, device=dev)[:n]
        keys = keys[keep]
    else:
        keys = keys[torch.randperm(keys.numel(), generator=g, device=dev)]
    u = torch.div(keys, n_items, rounding_mode="floor"); i = keys - u * n_items
    del keys
    rank = 16
    Pu = 0.35 * torch.randn(n_users, rank, generator=g, device=dev); Qi = 0.35 * torch.randn(n_items, rank, generator=gi, device=dev)
    bu = 0.3 * torch.randn(n_users, generator=g, device=dev); bi = 0.3 * torch.randn(n_items, generator=gi, device=dev)
    v = torch.empty(u.numel(), device=dev, dtype=torch.float32)
    step = 1 << 24
    for s0 in range(0, u.numel(), step):
        uu, ii = u[s0:s0 + step], i[s0:s0 + step]
        v[s0:s0 + step] = 3.6 + bu[uu] + bi[ii] + (Pu[uu] * Qi[ii]).sum(1) + 0.5 * torch.randn(uu.numel(), generator=g, device=dev)
    if levels == "half":
        v = torch.clamp(torch.round(v * 2) / 2, 0.5, 5.0)
    elif levels == "int":
        v = torch.clamp(torch.round(v), 1.0, 5.0)
    is_test = torch.rand(u.numel(), generator=g, device=dev) < test_fraction
    out = {}
    for name, m in (("train", ~is_test), ("test", is_test)):
        out[name] = (u[m].to(torch.int32).cpu().numpy(), i[m].to(torch.int32).cpu().numpy(), v[m].cpu().numpy())
    out["n_users"], out["n_items"] = n_users, n_items
    del u, i, v, is_test
    torch.cuda.empty_cache()
    return out


def implicit(n_users, n_items, n, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    return _pairs(rng, n_users, n_items, n)


def implicit_cuda(n_users, n_items, n, seed=1, device="cuda", pop_offset=None):
    """n distinct (user, item) events by the same activity / popularity laws, generated with torch on the GPU (benchmark
    plumbing for the 20M-event shape; deterministic for a seed on a given torch build, NOT the stream of implicit())."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev); g.manual_seed(int(seed))
    act = torch.exp(torch.randn(n_users, generator=g, device=dev, dtype=torch.float64))
    if os.environ.get("MMLB200_SYN_ACT_CLIP"):      # diagnostic: cap the user activity law (how much of an epoch is the heaviest user's chain?)
        act = torch.clamp(act, max=float(os.environ["MMLB200_SYN_ACT_CLIP"]))
    pop = 1.0 / (torch.arange(1, n_items + 1, device=dev, dtype=torch.float64) + (POP_OFFSET if pop_offset is None else pop_offset)) ** 0.8
    pop = pop[torch.randperm(n_items, generator=g, device=dev)]
    cdf = torch.cumsum(pop / pop.sum(), 0); cdf[-1] = 1.0
    ucdf = torch.cumsum(act / act.sum(), 0); ucdf[-1] = 1.0
    ar_u = torch.arange(n_users, device=dev, dtype=torch.int64); ar_i = torch.arange(n_items, device=dev, dtype=torch.int64)
    keys = torch.unique(torch.cat([ar_u * n_items + torch.randint(0, n_items, (n_users,), generator=g, device=dev),
                                   torch.randint(0, n_users, (n_items,), generator=g, device=dev) * n_items + ar_i]))
    while keys.numel() < n:
        draw = int((n - keys.numel()) * 1.15) + 1024
        uu = torch.searchsorted(ucdf, torch.rand(draw, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=n_users - 1)
        ii = torch.searchsorted(cdf, torch.rand(draw, generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, uu * n_items + ii]))
        del uu, ii
    keys = keys[torch.randperm(keys.numel(), generator=g, device=dev)[:n]]
    u = torch.div(keys, n_items, rounding_mode="floor"); i = keys - u * n_items
    out = (u.to(torch.int32).cpu().numpy(), i.to(torch.int32).cpu().numpy())
    del keys, u, i
    torch.cuda.empty_cache()
    return out


def named(name, scale=1.0):
    nu, ni, n, levels, k, seed = SHAPES[name]
    if scale != 1.0:
        nu, ni, n = max(int(nu * scale), 8), max(int(ni * scale), 8), max(int(n * scale * scale), 64)
    if levels is None:
        return implicit(nu, ni, n, seed), k
    return ratings(nu, ni, n, levels, seed), k

"""Pins the oracle's restatement of System.Random / Utils.Shuffle / Normal sampling.

System.Random is BCL code (not in the reference tree); the known answers below are the publicly
known first outputs of `new Random(seed).Next()` on .NET Framework.
"""
import numpy as np

from oracle import oracle as O


def test_system_random_known_answers():
    kat = {0: [1559595546, 1755192844, 1649316166],
           1: [534011718, 237820880, 1002897798],
           42: [1434747710, 302596119, 269548474]}
    for seed, exp in kat.items():
        r = O.Random(seed)
        assert [r.next() for _ in range(3)] == exp


def test_next_max_known_answers():
    r = O.Random(1)
    assert [r.next_max(10) for _ in range(10)] == [2, 1, 4, 7, 6, 4, 3, 9, 1, 6]


def test_negative_seed_uses_abs():
    a, b = O.Random(-7), O.Random(7)
    assert [a.next() for _ in range(5)] == [b.next() for _ in range(5)]


def test_next_double_in_unit_interval():
    r = O.Random(123)
    v = np.array([r.next_double() for _ in range(10000)])
    assert v.min() >= 0.0 and v.max() < 1.0
    assert abs(v.mean() - 0.5) < 0.02


def test_shuffle_is_fisher_yates_descending():
    # Utils.cs:52-64: for i = n-1..0: r = Next(i+1); swap(a[i], a[r]) -- i = 0 still draws
    n = 1000
    r1, r2, r3 = O.Random(5), O.Random(5), O.Random(5)
    a = r1.shuffle(np.arange(n))
    H = r2.shuffle_targets(n)
    b = np.arange(n, dtype=np.int32)
    O.lib().mo_shuffle_apply(b, H, n)
    assert np.array_equal(a, b)
    assert np.array_equal(np.sort(a), np.arange(n))
    ref = list(range(n))
    for i in range(n - 1, -1, -1):
        j = r3.next_max(i + 1)
        ref[i], ref[j] = ref[j], ref[i]
    assert np.array_equal(a, np.array(ref))
    # the RNG streams are in the same state afterwards (n draws each)
    assert r1.next() == r2.next() == r3.next()


def test_normal_sampling_moments_and_draw_count():
    r = O.Random(3)
    x = r.init_normal(200000, 0.0, 0.1)
    assert x.dtype == np.float32
    assert abs(float(x.mean())) < 1e-3
    assert abs(float(x.std()) - 0.1) < 1e-3
    # polar method: first value reproduced by hand from the uniform stream
    r1, r2 = O.Random(9), O.Random(9)
    while True:
        a, b = r1.next_double(), r1.next_double()
        v1, v2 = 2 * a - 1, 2 * b - 1
        s = v1 * v1 + v2 * v2
        if s >= 1.0 or s == 0.0:
            continue
        exp = 0.5 + 2.0 * (v1 * np.sqrt(-2.0 * np.log(s) / s))
        break
    got = O.lib().mo_normal_sample(r2.ref, 0.5, 2.0)
    assert abs(got - exp) < 1e-12

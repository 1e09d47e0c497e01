"""MyMediaLite.Random (Random.cs:23-64) for the host side of the CUDA recommenders: the BCL's System.Random
(Knuth's subtractive generator) plus MathNet.Numerics' Normal.Sample (polar Box-Muller), Utils.Shuffle's swap
targets, and MatrixExtensions.InitNormal. The host owns all randomness: the RNG stream is
sequential by definition, so it stays on the CPU, exactly as in the reference; the device only applies it
(mml_shuffle_apply) or, for large synthetic runs, uses its own counter-based generator (init_model)."""
import math

import numpy as np

MBIG = 2147483647
MSEED = 161803398


class SystemRandom:
    def __init__(self, seed):
        seed = int(seed)
        sub = MBIG if seed == -2147483648 else abs(seed)
        mj = MSEED - sub
        sa = [0] * 56
        sa[55] = mj
        mk = 1
        for i in range(1, 55):
            ii = (21 * i) % 55
            sa[ii] = mk
            mk = mj - mk
            if mk < 0:
                mk += MBIG
            mj = sa[ii]
        for _ in range(4):
            for i in range(1, 56):
                v = sa[i] - sa[1 + (i + 30) % 55]
                # 32-bit wrap-around, then the generator's own correction
                v = (v + 2 ** 31) % 2 ** 32 - 2 ** 31
                if v < 0:
                    v += MBIG
                sa[i] = v
        self.sa, self.inext, self.inextp = sa, 0, 21

    def _sample(self):
        self.inext = 1 if self.inext + 1 >= 56 else self.inext + 1
        self.inextp = 1 if self.inextp + 1 >= 56 else self.inextp + 1
        r = self.sa[self.inext] - self.sa[self.inextp]
        if r == MBIG:
            r -= 1
        if r < 0:
            r += MBIG
        self.sa[self.inext] = r
        return r

    def next(self, max_value=None):
        if max_value is None:
            return self._sample()
        return int(self._sample() * (1.0 / MBIG) * max_value)

    def next_double(self):
        return self._sample() * (1.0 / MBIG)

    # ---- consumers on the hot path --------------------------------------------------------------------------
    def shuffle_targets(self, n):
        """H[i] = Next(i + 1) drawn for i = n-1 .. 0 (Utils.cs:52-64)."""
        H = np.empty(n, np.int32)
        for i in range(n - 1, -1, -1):
            H[i] = self.next(i + 1)
        return H

    def shuffle(self, a):
        a = np.array(a, dtype=np.int32)
        for i in range(a.size - 1, -1, -1):
            r = self.next(i + 1)
            a[i], a[r] = a[r], a[i]
        return a

    def normal(self, mean, stddev):
        """MathNet.Numerics 3.x Normal.Sample: polar Box-Muller, two NextDouble per trial, first variate returned."""
        while True:
            v1 = 2.0 * self.next_double() - 1.0
            v2 = 2.0 * self.next_double() - 1.0
            r = v1 * v1 + v2 * v2
            if r < 1.0 and r != 0.0:
                break
        fac = math.sqrt(-2.0 * math.log(r) / r)
        return mean + stddev * v1 * fac

    def init_normal(self, rows, cols, mean, stddev):
        """MatrixExtensions.InitNormal (DataType/MatrixExtensions.cs:62-69): row-major, cast to float."""
        out = np.empty(rows * cols, np.float32)
        for t in range(out.size):
            out[t] = self.normal(mean, stddev)
        return out.reshape(rows, cols)


_instance = None
_seed = None


def seed(value):
    """Random.Seed setter (Random.cs:38-50): re-creates the instance."""
    global _instance, _seed
    _seed = int(value)
    _instance = SystemRandom(_seed)


def get_instance():
    """Random.GetInstance (Random.cs:54-63): time-seeded when no seed was set."""
    global _instance
    if _instance is None:
        import time
        _instance = SystemRandom(int(time.time()) & 0x7FFFFFFF)
    return _instance

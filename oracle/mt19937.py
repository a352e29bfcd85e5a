"""numpy *legacy* RandomState (MT19937) restated in pure Python / numpy.  TEST INFRASTRUCTURE.

The reference seeds only numpy's global legacy generator
(robot-learning.py:19 ``np.random.seed(configuration.RANDOM_SEED)``) and draws from it at
environment.py:29, 33, 38, 43, 47, 54, 136, 156, 159 and robot.py:111, 640, 802-815.
numpy itself is the third-party dependency that holds the algorithm (numpy 2.3.5 here,
``numpy/random/src/mt19937/mt19937.c``, ``legacy-distributions.c``, ``_mt19937.pyx``
``_legacy_seeding``); the published algorithm is restated below and pinned in
``tests/test_oracle_mt19937.py`` against ``np.random.RandomState`` directly.

Draw protocol (all verified bit-exact against numpy):
  seed(s)                : init_genrand - mt[0]=s; mt[i]=1812433253*(mt[i-1]^(mt[i-1]>>30))+i
  random_uint32          : standard MT19937 tempering, regenerate every 624 outputs
  random_double          : a=u32>>5, b=u32>>6, (a*67108864+b)/2**53      (two u32 per double)
  uniform(lo,hi)         : lo + (hi-lo)*double
  interval(max)          : mask = next_pow2(max+1)-1; loop v = u32 & mask until v <= max
  choice([0,1,2,3])      : interval(3)      -> one u32 & 3
  choice(n,B,replace=False) : legacy shuffle of arange(n): for i=n-1..1: j=interval(i); swap; take [:B]
  gauss (np.random.normal): polar Box-Muller with a cached spare value
"""
import math

import numpy as np

N = 624
M = 397
_U32 = 0xFFFFFFFF


class LegacyMT19937:
    """One numpy-legacy MT19937 stream (scalar, pure Python)."""

    def __init__(self, seed=None):
        self.mt = [0] * N
        self.pos = N
        self.has_gauss = 0
        self.gauss_cache = 0.0
        if seed is not None:
            self.seed(seed)

    # numpy/random/src/mt19937/mt19937.c: mt19937_seed
    def seed(self, seed):
        seed &= _U32
        mt = self.mt
        for pos in range(N):
            mt[pos] = seed
            seed = (1812433253 * (seed ^ (seed >> 30)) + pos + 1) & _U32
        self.pos = N
        self.has_gauss = 0
        self.gauss_cache = 0.0

    # mt19937.c: mt19937_gen
    def _regenerate(self):
        mt = self.mt
        for kk in range(N - M):
            y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        for kk in range(N - M, N - 1):
            y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        y = (mt[N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        self.pos = 0

    def random_uint32(self):
        if self.pos == N:
            self._regenerate()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & _U32

    def random_double(self):
        a = self.random_uint32() >> 5
        b = self.random_uint32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    def uniform(self, lo, hi):
        return lo + (hi - lo) * self.random_double()

    # legacy-distributions / distributions.c: random_interval (32-bit branch)
    def interval(self, maxv):
        if maxv == 0:
            return 0
        mask = maxv
        mask |= mask >> 1
        mask |= mask >> 2
        mask |= mask >> 4
        mask |= mask >> 8
        mask |= mask >> 16
        while True:
            v = self.random_uint32() & mask
            if v <= maxv:
                return v

    # legacy-distributions.c: legacy_gauss
    def gauss(self):
        if self.has_gauss:
            tmp = self.gauss_cache
            self.gauss_cache = 0.0
            self.has_gauss = 0
            return tmp
        while True:
            x1 = 2.0 * self.random_double() - 1.0
            x2 = 2.0 * self.random_double() - 1.0
            r2 = x1 * x1 + x2 * x2
            if r2 < 1.0 and r2 != 0.0:
                break
        f = math.sqrt(-2.0 * math.log(r2) / r2)
        self.gauss_cache = f * x1
        self.has_gauss = 1
        return f * x2

    def normal(self, loc, scale):
        return loc + scale * self.gauss()

    # RandomState.shuffle (1-d ndarray branch, _shuffle_raw) + permutation()[:size]
    def choice_no_replace(self, n, size):
        x = list(range(n))
        for i in range(n - 1, 0, -1):
            j = self.interval(i)
            x[i], x[j] = x[j], x[i]
        return np.array(x[:size], dtype=np.int64)

    # state import/export against numpy for pinning tests
    def get_state(self):
        return ("MT19937", np.array(self.mt, dtype=np.uint32), self.pos, self.has_gauss, self.gauss_cache)

    def set_state(self, st):
        self.mt = [int(v) for v in st[1]]
        self.pos = int(st[2])
        self.has_gauss = int(st[3])
        self.gauss_cache = float(st[4])


class LegacyMT19937Bank:
    """n independent legacy streams, vectorised with numpy (state layout [624, n] like the device bank)."""

    def __init__(self, seeds):
        seeds = np.asarray(seeds, dtype=np.uint64) & _U32
        n = seeds.shape[0]
        self.n = n
        mt = np.zeros((N, n), dtype=np.uint64)
        s = seeds.copy()
        for pos in range(N):
            mt[pos] = s
            s = (1812433253 * (s ^ (s >> np.uint64(30))) + np.uint64(pos + 1)) & _U32
        self.mt = mt.astype(np.uint32)
        self.pos = np.full(n, N, dtype=np.int32)

    def _regenerate(self, cols):
        mt = self.mt[:, cols].astype(np.uint32)

        def tw(cur, nxt, far):
            y = (cur & np.uint32(0x80000000)) | (nxt & np.uint32(0x7FFFFFFF))
            return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))

        for kk in range(N - M):
            mt[kk] = tw(mt[kk], mt[kk + 1], mt[kk + M])
        for kk in range(N - M, N - 1):
            mt[kk] = tw(mt[kk], mt[kk + 1], mt[kk + (M - N)])
        mt[N - 1] = tw(mt[N - 1], mt[0], mt[M - 1])
        self.mt[:, cols] = mt
        self.pos[cols] = 0

    def random_uint32(self, active=None):
        """One u32 per stream where ``active`` (bool [n]); inactive streams are untouched and return 0."""
        if active is None:
            active = np.ones(self.n, dtype=bool)
        need = np.nonzero(active & (self.pos == N))[0]
        if need.size:
            self._regenerate(need)
        idx = np.nonzero(active)[0]
        out = np.zeros(self.n, dtype=np.uint32)
        y = self.mt[self.pos[idx], idx].astype(np.uint32)
        self.pos[idx] += 1
        y ^= y >> np.uint32(11)
        y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
        y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
        y ^= y >> np.uint32(18)
        out[idx] = y
        return out

    def random_double(self, active=None):
        a = (self.random_uint32(active) >> np.uint32(5)).astype(np.float64)
        b = (self.random_uint32(active) >> np.uint32(6)).astype(np.float64)
        return (a * 67108864.0 + b) / 9007199254740992.0

"""Replica of R's default RNG (Mersenne-Twister, "Rejection" sampling, R >= 3.6)
for `set.seed(s); runif(k); sample(n, k)`.

TEST INFRASTRUCTURE ONLY (oracle/).  Needed to reproduce the reference's own
sub-samples (`set.seed(123); sample(500, 100)` in R-pkg/vignettes/Introduction.Rmd:113-115
and R-pkg/tests/testthat/test_SVC_mle.R:6-7) without an R interpreter.
Algorithm restated from base R's RNG.c (MT_sgenrand/MT_genrand, R_unif_index).
"""
from __future__ import annotations

import math


class RRandom:
    N, M = 624, 397

    def __init__(self, seed: int):
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        iseed = []
        for _ in range(625):
            s = (69069 * s + 1) & 0xFFFFFFFF
            iseed.append(s)
        self.mt = iseed[1:625]
        self.mti = self.N  # dummy[0] = 624 after FixupSeeds

    def _genrand(self) -> float:
        N, M, mt = self.N, self.M, self.mt
        if self.mti >= N:
            for kk in range(N):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % N] & 0x7FFFFFFF)
                v = mt[(kk + M) % N] ^ (y >> 1)
                if y & 1:
                    v ^= 0x9908B0DF
                mt[kk] = v
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return (y & 0xFFFFFFFF) * 2.3283064365386963e-10

    def unif_rand(self) -> float:
        v = self._genrand()
        if v <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v

    def runif(self, k: int):
        return [self.unif_rand() for _ in range(k)]

    def _rbits(self, bits: int) -> float:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn: float) -> int:
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return int(dv)

    def sample(self, n: int, k: int):
        """sample(n, k) without replacement -> 1-based indices."""
        x = list(range(n))
        out = []
        m = n
        for _ in range(k):
            j = self.unif_index(m)
            out.append(x[j] + 1)
            m -= 1
            x[j] = x[m]
        return out

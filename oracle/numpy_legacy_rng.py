"""Test infrastructure: restatement of numpy's legacy RandomState stream (MT19937) in plain Python.

The reference's synthetic KV generator (nerf_attention/extract.py:182-259) draws everything from
``np.random.RandomState(layer * num_kv_heads + head)`` -- uniform, randint and randn.  numpy is a third-party
dependency of the reference (``numpy`` unpinned in pyproject.toml; 2.3 in this image) whose source is not under
/root/reference, so the algorithm is restated here from its published definition
(numpy/random/src/mt19937/mt19937.c, numpy/random/src/legacy/legacy-distributions.c,
numpy/random/src/distributions/distributions.c) and pinned against numpy itself in
tests/test_oracle_golden.py.  csrc/synth.cuh implements exactly this stream on the GPU.

Only tests may import this module.
"""

from __future__ import annotations

import math


class LegacyRandomState:
    def __init__(self, seed: int) -> None:
        # init_genrand (mt19937_seed): the integer-seed path of RandomState(seed)
        mt = [0] * 624
        mt[0] = seed & 0xffffffff
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xffffffff
        self.mt, self.pos = mt, 624
        self.has_gauss, self.gauss = False, 0.0
        self.words = 0                                   # 32-bit words consumed so far

    def _generate(self) -> None:
        mt = self.mt
        for k in range(624):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7fffffff)
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
        self.pos = 0

    def next_uint32(self) -> int:
        if self.pos == 624:
            self._generate()
        y = self.mt[self.pos]
        self.pos += 1
        self.words += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y

    def next_double(self) -> float:
        a, b = self.next_uint32() >> 5, self.next_uint32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    def uniform(self, low: float, high: float) -> float:
        return low + (high - low) * self.next_double()

    def randint(self, low: int, high: int) -> int:
        """[low, high): masked rejection on 32-bit draws; a one-value range consumes nothing."""
        rng = high - low - 1
        if rng == 0:
            return low
        mask = rng
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.next_uint32() & mask
            if v <= rng:
                return low + v

    def gauss1(self) -> float:
        """legacy_gauss: polar method; returns f*x2 and caches f*x1 for the next call."""
        if self.has_gauss:
            self.has_gauss = False
            return self.gauss
        while True:
            x1 = 2.0 * self.next_double() - 1.0
            x2 = 2.0 * self.next_double() - 1.0
            r2 = x1 * x1 + x2 * x2
            if r2 < 1.0 and r2 != 0.0:
                break
        f = math.sqrt(-2.0 * math.log(r2) / r2)
        self.gauss, self.has_gauss = f * x1, True
        return f * x2

    def randn(self, n: int) -> list[float]:
        return [self.gauss1() for _ in range(n)]

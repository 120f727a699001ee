"""std::mt19937 + libstdc++ std::generate_canonical<double, 53> restated (TEST INFRASTRUCTURE ONLY, oracle/__init__.py).

whisper.cpp gives every decoder a ``std::mt19937(0)`` and draws tokens at temperature > 0 with
``std::discrete_distribution`` (whisper_sample_token), whose call operator consumes one canonical double =
two 32-bit outputs: (x1 + x2 * 2^32) / 2^64 [libstdc++ bits/random.tcc].
"""


class Mt19937:
    def __init__(self, seed: int = 0):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.idx = mt, 624

    def next_u32(self) -> int:
        mt = self.mt
        if self.idx >= 624:
            for i in range(624):
                y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def uniform(self) -> float:
        x1, x2 = float(self.next_u32()), float(self.next_u32())
        u = (x1 + x2 * 4294967296.0) / 18446744073709551616.0
        return u if u < 1.0 else 1.0 - 2.0 ** -53

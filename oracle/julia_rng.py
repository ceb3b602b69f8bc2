"""Julia's default random stream, restated for the parity tests.

TEST INFRASTRUCTURE (like everything under oracle/): imported only by tests/ and
tests/golden/make_golden.py.  Never imported by the product package.

The reference draws every variate from Julia's *global* task-local RNG with scalar `rand()`
calls (src/problem.jl:153, src/chvdiffeq.jl:71, src/utils.jl:47, src/rejectiondiffeq.jl:35-36), and
its test-suite holds integers that depend on that stream after `Random.seed!(1234)`
(test/runtests.jl:57-69).  Reproducing the stream here turns those integers into golden vectors
for the oracle and for the GPU replay mode, without a Julia runtime:

  * generator: xoshiro256++ (Blackman & Vigna), Julia >= 1.7 `Xoshiro` / `TaskLocalRNG`
        tmp = s0 + s3 ; result = rotl(tmp, 23) + s0 ; t = s1 << 17
        s2 ^= s0 ; s3 ^= s1 ; s1 ^= s2 ; s0 ^= s3 ; s2 ^= t ; s3 = rotl(s3, 45)
  * `Random.seed!(n::Integer)` (Julia 1.7 - 1.11, n >= 0): the state is the SHA-256 digest of the
    little-endian UInt32 limbs of n, read as four little-endian UInt64 words
  * `rand()` (Float64): (next() >>> 11) * 2^-53

The restatement is self-validating: fed to the oracle's Rejection / CHV restatements it reproduces
the reference's stored results [0, 26, 83] (examples/sir-rejection.jl, test/runtests.jl:64-69) and
[0, 28, 81] (examples/sir.jl, third solve on the continuing stream, test/runtests.jl:57-62); a wrong
seeding rule, generator or float conversion cannot hit those by chance.
"""
import hashlib
import struct

import numpy as np

_M = (1 << 64) - 1


def _rotl(x, k):
    return ((x << k) & _M) | (x >> (64 - k))


class Xoshiro:
    """Julia's `Xoshiro` seeded like `Random.seed!(seed)`."""

    def __init__(self, seed):
        if seed < 0:
            raise ValueError("only non-negative integer seeds are restated")
        limbs = []
        n = int(seed)
        while True:
            limbs.append(n & 0xFFFFFFFF)
            n >>= 32
            if n == 0:
                break
        digest = hashlib.sha256(b"".join(struct.pack("<I", w) for w in limbs)).digest()
        self.s = list(struct.unpack("<4Q", digest))

    def next_u64(self):
        s0, s1, s2, s3 = self.s
        tmp = (s0 + s3) & _M
        res = (_rotl(tmp, 23) + s0) & _M
        t = (s1 << 17) & _M
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= t
        s3 = _rotl(s3, 45)
        self.s = [s0, s1, s2, s3]
        return res

    def rand(self):
        """Scalar `rand()`: Float64 in [0, 1)."""
        return (self.next_u64() >> 11) * 2.0 ** -53

    def rand_stream(self, n):
        """The next n scalar `rand()` values."""
        return np.array([self.rand() for _ in range(n)], dtype=np.float64)


def stream(seed, n):
    """First n values of `Random.seed!(seed); [rand() for _ in 1:n]`."""
    return Xoshiro(seed).rand_stream(n)

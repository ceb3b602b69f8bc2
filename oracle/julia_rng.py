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
seeding rule, generator or float conversion cannot hit those by chance.  With the array form `rand(N)`
(rand_array) it also reproduces `result.xd[1,end] == 100` of examples/neuron_rejection_exact.jl
(test/runtests.jl:71-75: RejectionExact, second of two solves on a continuing stream).
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

    def rand_array(self, n):
        """`rand(n)` (Vector{Float64}), which is NOT n scalar draws: Julia 1.7 - 1.10 fill arrays of >= 64 bytes from
        eight SIMD xoshiro256++ lanes forked off the main generator — lane state word w = MULT[w] * rand(UInt64), the
        32 seeds drawn word by word (s0 of the 8 lanes first) — 8 values per round, lane by lane; the tail that does
        not fill a round comes from the main generator (stdlib/Random/src/XoshiroSimd.jl: xoshiro_bulk)."""
        mult = (0x02011ce34bce797f, 0x5a94851fb48a6e05, 0x3688cf5d48899fa7, 0x867b4bb4c42e5661)
        out = np.empty(n, dtype=np.float64)
        i = 0
        if n * 8 >= 64:
            s0, s1, s2, s3 = ([(m * self.next_u64()) & _M for _ in range(8)] for m in mult)
            while i + 8 <= n:
                for l in range(8):
                    res = (_rotl((s0[l] + s3[l]) & _M, 23) + s0[l]) & _M
                    t = (s1[l] << 17) & _M
                    s2[l] ^= s0[l]
                    s3[l] ^= s1[l]
                    s1[l] ^= s2[l]
                    s0[l] ^= s3[l]
                    s2[l] ^= t
                    s3[l] = _rotl(s3[l], 45)
                    out[i + l] = (res >> 11) * 2.0 ** -53
                i += 8
        while i < n:
            out[i] = self.rand()
            i += 1
        return out


def stream(seed, n):
    """First n values of `Random.seed!(seed); [rand() for _ in 1:n]`."""
    return Xoshiro(seed).rand_stream(n)

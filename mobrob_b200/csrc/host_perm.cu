// Host-side minibatch permutation: the part of [SB3 2.0.0] RolloutBuffer.get that stays on the host
// (``indices = np.random.permutation(n_envs * n_steps)``, reached from
// src/mobrob/rl_control/ppo.py:73-74), written so that it keeps up with a B200: numpy's
// Fisher-Yates costs ~20 ns per index at n = 1.2e6 (one dependent cache miss per swap), ten
// epochs of that are longer than the whole device iteration.
//
// Scatter shuffle (Rao-Sandelius): deal the indices into 2^k buckets with uniform random keys
// (sequential writes), then Fisher-Yates inside each bucket (cache resident).  Dealing
// independently and shuffling every bucket uniformly yields a uniformly distributed permutation.
// The stream is xoshiro256** seeded by splitmix64 from (seed, stream): a pure function of its
// arguments, independent of thread count.  Plain host code; no device work.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

struct Xoshiro {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Xoshiro(uint64_t seed, uint64_t stream) {
        uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
        for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
    }
    static inline uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
    inline uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    // unbiased integer in [0, bound), bound < 2^32 (Lemire's multiply-shift with rejection)
    inline uint32_t below(uint32_t bound) {
        uint64_t m = (uint64_t)(uint32_t)next() * bound;
        uint32_t lo = (uint32_t)m;
        if (lo < bound) {
            const uint32_t thresh = (0u - bound) % bound;
            while (lo < thresh) {
                m = (uint64_t)(uint32_t)next() * bound;
                lo = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

}  // namespace

extern "C" int mr_host_permutation(uint64_t seed, uint64_t stream, int64_t n, int64_t* out) {
    MR_REQUIRE(out != nullptr, "NULL argument");
    MR_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "n out of range");
    if (n == 0) return MR_OK;
    Xoshiro g(seed, stream);
    // Buckets of <= 64K indices (256 KB as uint32: L2 resident).  Smaller buckets shuffle faster (L1) but
    // need more write streams in the deal, and beyond ~32 concurrent streams the deal falls off a cliff on
    // the hosts measured (1.4 ns per index at 16 streams, 6-8 ns at 64+): at n = 1.2 M, 32 buckets take
    // 8 ms per permutation, 256 buckets 19 ms.
    int bits = 0;
    while ((n >> bits) > 65536 && bits < 12) ++bits;
    const int K = 1 << bits;
    // scratch lives with the calling thread (the feeder's workers call this ten times per iteration:
    // fresh vectors cost an allocation and first-touch page faults on 7 MB every time)
    static thread_local std::vector<uint32_t> tmp;
    static thread_local std::vector<uint16_t> key;
    if (tmp.size() < (size_t)n) tmp.resize((size_t)n);
    if (K > 1) {
        if (key.size() < (size_t)n) key.resize((size_t)n);
        int64_t count[4097];
        int64_t head[4096];
        for (int b = 0; b <= K; ++b) count[b] = 0;
        const int per = 64 / (bits > 0 ? bits : 1);   // keys per 64 random bits
        const uint64_t mask = (uint64_t)(K - 1);
        for (int64_t i = 0; i < n;) {
            uint64_t r = g.next();
            for (int q = 0; q < per && i < n; ++q, ++i, r >>= bits) {
                const uint16_t b = (uint16_t)(r & mask);
                key[(size_t)i] = b;
                ++count[b + 1];
            }
        }
        for (int b = 0; b < K; ++b) count[b + 1] += count[b];
        for (int b = 0; b < K; ++b) head[b] = count[b];
        uint32_t* t = tmp.data();
        const uint16_t* kk = key.data();
        for (int64_t i = 0; i < n; ++i) t[head[kk[i]]++] = (uint32_t)i;
        for (int b = 0; b < K; ++b) {
            uint32_t* a = t + count[b];
            const int64_t len = count[b + 1] - count[b];
            for (int64_t i = len - 1; i > 0; --i) {
                const uint32_t j = g.below((uint32_t)i + 1);
                const uint32_t v = a[i]; a[i] = a[j]; a[j] = v;
            }
            int64_t* o = out + count[b];   // widen while the bucket is still in cache
            for (int64_t i = 0; i < len; ++i) o[i] = (int64_t)a[i];
        }
    } else {
        uint32_t* t = tmp.data();
        for (int64_t i = 0; i < n; ++i) t[i] = (uint32_t)i;
        for (int64_t i = n - 1; i > 0; --i) {
            const uint32_t j = g.below((uint32_t)i + 1);
            const uint32_t v = t[i]; t[i] = t[j]; t[j] = v;
        }
        for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)t[i];
    }
    return MR_OK;
}

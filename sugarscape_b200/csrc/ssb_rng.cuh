// ssb_rng.cuh -- the two random-number modes of the engine.
//
// Mode E: CPython's MT19937 main stream replayed word for word.  The Python-level algorithms
//   random._randbelow_with_getrandbits (random.py:242) and random.shuffle (random.py:350) are
//   restated on top of genrand_uint32 (Modules/_randommodule.c).  State lives in shared memory
//   of the single ordered CTA.
// Mode S: counter-based substreams.  Word j of agent `id` in timestep t is
//   mix64(step_key(seed, t) ^ (id << 32 | site << 24 | j)) >> 32 for the j-th word drawn at call
//   site `site` of the transaction; the processing order of a step is the
//   ascending order of a keyed 4-round Feistel bijection of the agent id.
#pragma once
#include <stdint.h>

namespace ssb {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t step_key(uint64_t seed, int32_t t) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(uint32_t)(t + 1));
}

// keyed bijection on [0, 2^bits), bits even: unique priorities without a sort
__host__ __device__ __forceinline__ uint32_t priority_of(uint64_t key, uint32_t id, int bits) {
    const int half = bits >> 1;
    const uint32_t mask = (1u << half) - 1u;
    uint32_t L = (id >> half) & mask, R = id & mask;
#pragma unroll
    for (uint32_t round = 0; round < 4; round++) {
        uint32_t f = (uint32_t)(mix64(key ^ 0xA5A5A5A5ull ^ ((uint64_t)(round + 1) << 56) ^ R) >> 40) & mask;
        uint32_t nl = R;
        R = L ^ f;
        L = nl;
    }
    return (L << half) | R;
}

__device__ __forceinline__ int bit_length(uint32_t n) { return 32 - __clz(n); }

// ---- mode E -------------------------------------------------------------------------------
struct RngExact {
    uint32_t *mt;   // 624 words + index at mt[624]

    __device__ void twist() {
        uint32_t y;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt[624] = 0;
    }
    __device__ __forceinline__ void at(uint32_t) {}      // one global stream: sites are not separated
    template <int G> __device__ __forceinline__ void prefetch(uint32_t *, int) {}
    __device__ __forceinline__ void end_prefetch() {}
    __device__ uint32_t next() {
        if (mt[624] >= 624u) twist();
        uint32_t y = mt[mt[624]++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

// ---- mode S -------------------------------------------------------------------------------
// Call sites of the agent transaction that draw random numbers (reference agent.py:1401, 560-564,
// 616, 505-513).  Each site of each agent has its own substream, so the two commit phases of the
// scalable kernel need no RNG state between them.
enum { SITE_MOVE = 0, SITE_TAGGING = 1, SITE_TRADING = 2, SITE_REPRODUCTION = 3, /* 4 = new tribe tags */
       SITE_LENDING = 5, SITE_DISEASE = 6 };

struct RngStream {
    uint64_t key;
    uint32_t id, site, ctr;
    const uint32_t *pre;     // words 0..63 of the current site, computed by the whole warp
    __device__ __forceinline__ uint32_t word(uint32_t j) const {
        return (uint32_t)(mix64(key ^ (((uint64_t)id << 32) | ((uint64_t)site << 24) | j)) >> 32);
    }
    __device__ __forceinline__ void at(uint32_t site_) { site = site_; ctr = 0; pre = nullptr; }
    // all G lanes of the group: fill buf[0..63] with the first 64 words of the current site
    template <int G> __device__ __forceinline__ void prefetch(uint32_t *buf, int lane) {
#pragma unroll
        for (int j = lane; j < 64; j += G) buf[j] = word((uint32_t)j);
        pre = buf;
    }
    __device__ __forceinline__ void end_prefetch() { pre = nullptr; }
    __device__ __forceinline__ uint32_t next() {
        const uint32_t j = ctr++;
        return (pre && j < 64u) ? pre[j] : word(j);
    }
};

// random._randbelow: k = n.bit_length(); draw k-bit words until one is < n
template <class Rng>
__device__ __forceinline__ uint32_t rand_below(Rng &rng, uint32_t n) {
    const int shift = 32 - bit_length(n);
    uint32_t v = rng.next() >> shift;
    while (v >= n) v = rng.next() >> shift;
    return v;
}

// random.random(): (a >> 5, b >> 6) -> 53 bits (Modules/_randommodule.c)
template <class Rng>
__device__ __forceinline__ double rand_double(Rng &rng) {
    const uint32_t a = rng.next() >> 5, b = rng.next() >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

// random.shuffle: for i = n-1 .. 1: j = _randbelow(i + 1); swap(x[i], x[j])
template <class Rng, class T>
__device__ __forceinline__ void shuffle(Rng &rng, T *x, int n) {
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = rand_below(rng, (uint32_t)i + 1u);
        T tmp = x[i]; x[i] = x[j]; x[j] = tmp;
    }
}

}  // namespace ssb

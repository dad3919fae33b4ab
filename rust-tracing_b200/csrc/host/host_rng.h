// Seeded host RNG for scene layout, BVH axis draws and Perlin tables.
// The reference uses rand 0.8.5's unseeded thread_rng (ChaCha12) at main.rs:70-91,523,613,
// bvh.rs:32 and perlin.rs:18-20,76; only the distributions are part of its behaviour
// (random::<f64>() in [0,1), gen_range half-open / inclusive), so a seeded generator with the
// same distributions replaces it. Generator: xoshiro256++ seeded through splitmix64.
#pragma once
#include <cstdint>

namespace rt_host {

struct HostRng {
    uint64_t s[4];
    explicit HostRng(uint64_t seed) {
        uint64_t z = seed;
        for (int i = 0; i < 4; ++i) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            s[i] = x ^ (x >> 31);
        }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    // rand::random::<f64>(): uniform in [0,1) with 53 random bits.
    double random() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // gen_range(min..max) for floats: half-open.
    double range(double lo, double hi) { return lo + (hi - lo) * random(); }
    // gen_range(lo..=hi) for integers.
    int range_inclusive(int lo, int hi) {
        const uint64_t n = (uint64_t)(hi - lo + 1);
        const uint64_t r = (uint64_t)(((unsigned __int128)next_u64() * n) >> 64);
        return lo + (int)r;
    }
};

}  // namespace rt_host

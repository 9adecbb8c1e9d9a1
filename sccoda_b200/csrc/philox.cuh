// Device Philox4x32-10 counter RNG and the stream contract of the samplers.
//
// The reference draws momenta / accept uniforms from TensorFlow's stateful generator
// (tf.random.set_seed, tests/unit_tests.py:165-166); that stream cannot be reproduced, so the
// contract below is ours.  counter = (index, step, chain, stream), key = 64-bit seed.
// Every draw is addressed, not sequenced: any thread can regenerate any draw, chains are
// independent of the launch geometry and of how (dataset, chain) pairs are sharded over GPUs.
#pragma once
#include <cstdint>

namespace sccoda {

enum : uint32_t {
    STREAM_MOMENTUM = 0,   // index = flat parameter index
    STREAM_ACCEPT = 1,     // index = 0
    STREAM_NUTS_DIR = 2,   // index = tree depth
    STREAM_NUTS_LEAF = 3,  // index = leaf serial number within the transition
    STREAM_NUTS_TOP = 4    // index = tree depth
};

struct RngKey {
    uint32_t k0, k1, chain;
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 24-bit uniform in (0,1): ((w >> 8) + 0.5) * 2^-24 — exactly representable in fp32 and fp64,
// so the fp32 and fp64 builds (and the CPU oracle) consume identical uniforms.
template <typename T>
__device__ __forceinline__ T u01(uint32_t w) {
    return (static_cast<T>(w >> 8) + T(0.5)) * T(1.0 / 16777216.0);
}

__device__ __forceinline__ float box_muller(float u0, float u1) {
    return sqrtf(-2.0f * __logf(u0)) * cospif(2.0f * u1);
}
__device__ __forceinline__ double box_muller(double u0, double u1) {
    return sqrt(-2.0 * log(u0)) * cospi(2.0 * u1);
}

template <typename T>
__device__ __forceinline__ T rng_normal(const RngKey& k, uint32_t index, uint32_t step) {
    const uint4 w = philox4x32_10(index, step, k.chain, STREAM_MOMENTUM, k.k0, k.k1);
    return box_muller(u01<T>(w.x), u01<T>(w.y));
}
template <typename T>
__device__ __forceinline__ T rng_uniform(const RngKey& k, uint32_t index, uint32_t step, uint32_t stream) {
    return u01<T>(philox4x32_10(index, step, k.chain, stream, k.k0, k.k1).x);
}
__device__ __forceinline__ uint32_t rng_bit(const RngKey& k, uint32_t index, uint32_t step, uint32_t stream) {
    return philox4x32_10(index, step, k.chain, stream, k.k0, k.k1).x & 1u;
}

}  // namespace sccoda

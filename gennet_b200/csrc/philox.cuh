// Counter-based RNG (Philox4x32-10, Salmon et al. 2011) so that every random number is a pure
// function of (seed, global sample index, element index): the sample stream does not depend on
// the number of GPUs or on how a batch is split across ranks.  Replaces the host
// NumPy MT19937 draws of gw_template_maker.py:187-188 and bbhMahoGANy.py:1161,1247,1277,1295.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gn {

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// uniform in (0,1), 24-bit resolution, never 0 or 1
__host__ __device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}

// two standard normals for element `elem` of sample `sample` (stream 0: synthesis noise)
__device__ __forceinline__ float2 philox_normal2(unsigned long long seed, unsigned long long sample, uint32_t elem,
                                                 uint32_t stream = 0) {
    uint4 c = make_uint4(elem, (uint32_t)sample, (uint32_t)(sample >> 32), stream);
    uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return box_muller(r.x, r.y);
}

// raw 4x32 bits for flat element block `blk` (4 outputs per counter) at `offset`
__device__ __forceinline__ uint4 philox_flat(unsigned long long seed, unsigned long long blk, uint32_t stream) {
    uint4 c = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, stream);
    return philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

}  // namespace gn

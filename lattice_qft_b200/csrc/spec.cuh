// spec.cuh — the decision spec shared by the device kernels and the host side of liblqft.
//
// Everything that decides which random bits reach which site lives here, once, so that the
// device path, the host helpers exported through lqft_host_* and (independently restated)
// the test oracle agree bit for bit.  See include/lqft.h for the prose version.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LQFT_HD __host__ __device__ __forceinline__
#else
#define LQFT_HD inline
#endif

namespace lqft {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr int kPhiloxRounds = 10;

struct u32x4 {
    uint32_t x, y, z, w;
};

LQFT_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// Philox4x32-10, counter (c0,c1,c2,c3), key (k0,k1).
LQFT_HD u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                         uint32_t k1) {
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(kPhiloxM0, c0, hi0, lo0);
        mulhilo32(kPhiloxM1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return u32x4{c0, c1, c2, c3};
}

LQFT_HD uint32_t word_of(const u32x4& r, uint32_t w) {
    return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}

// 64x64 -> high 64 bits.
LQFT_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// Uniform site in [0,n): counter (rep, stream, ctr_lo, ctr_hi).
LQFT_HD uint64_t pick_site(uint32_t k0, uint32_t k1, uint32_t stream, uint64_t counter,
                           uint32_t rep, uint64_t n) {
    u32x4 r = philox4x32(rep, stream, (uint32_t)counter, (uint32_t)(counter >> 32), k0, k1);
    uint64_t v = (uint64_t)r.x | ((uint64_t)r.y << 32);
    return mulhi64(v, n);
}

// Hot-start height of global site idx: top byte of word idx&3 of counter (idx>>2, HOTSTART, hi, 0).
LQFT_HD int32_t hot_value(uint32_t k0, uint32_t k1, uint64_t idx) {
    uint64_t g = idx >> 2;
    u32x4 r = philox4x32((uint32_t)g, 3u, (uint32_t)(g >> 32), 0u, k0, k1);
    uint32_t w = word_of(r, (uint32_t)(idx & 3u));
    return (int32_t)(int8_t)(w >> 24);
}

// Wilson offset (fields.rs:556-561, 821-837): the +y bond leaving (x,0,t) is shifted by +1
// for x < width, t < height.  Seen from the y==0 site (direction 1) the difference gains +1,
// seen from the y==1 site (direction 4) it gains -1.
LQFT_HD int32_t wilson_offset(uint32_t Y, uint32_t x, uint32_t y, uint32_t t, uint32_t width,
                              uint32_t height) {
    if (x >= width || t >= height) return 0;
    int32_t off = 0;
    if (y == 0u) off += 1;
    if (y == 1u % Y) off -= 1;
    return off;
}

}  // namespace lqft

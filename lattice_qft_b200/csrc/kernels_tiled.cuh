// kernels_tiled.cuh — the fast half-sweep: "marching" register-tiled kernel.
//
// v1 (kernels_v1.cuh) is issue bound on B200: 356 warp instructions per 128 site updates, most of
// them index arithmetic, 64-bit addressing and per-thread prologue (profiles/r1_v1_sweep_ncu.txt:
// issue slots 75 % busy, L2 24 %, DRAM 27 %).  This kernel removes that overhead:
//   * the row length is a template parameter (X = 8 G), so a row of one colour is G int4 quads and a
//     group of G lanes covers it: the x neighbour outside the lane's own quad comes from one SHFL
//     (periodic wrap included) instead of a global load;
//   * each thread marches MARCH consecutive rows (y) of one t-plane and keeps the other colour's
//     rows y-1, y, y+1 in registers, so every step issues 4 LDG.128 (own row, other colour row y+1,
//     planes t-1 and t+1) instead of 6 + 1;
//   * the warps of a block take consecutive t-planes, so the t+-1 loads of one warp are the
//     same-plane loads of its neighbours: they hit in L1;
//   * one Philox4x32-10 call per quad, with the key schedule hoisted out of the marching loop;
//   * prologue (threshold table to shared memory) and epilogue (acceptance count) are paid once per
//     MARCH steps.
// Results are bit-identical to v1 and to the oracle: same sites, same counters, same words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_v1.cuh"

namespace lqft {

constexpr int kMarchWarps = 4;  // warps (= t-planes) per block

// Round keys of Philox4x32-10 (key + r * W).  For single-chain handles they travel as a kernel
// parameter, so every LOP3 of the cipher reads its key straight from the constant bank.
struct RoundKeys {
    uint32_t k[kPhiloxRounds][2];
};

inline RoundKeys make_round_keys(uint32_t k0, uint32_t k1) {
    RoundKeys rk;
    for (int r = 0; r < kPhiloxRounds; ++r) {
        rk.k[r][0] = k0 + (uint32_t)r * kPhiloxW0;
        rk.k[r][1] = k1 + (uint32_t)r * kPhiloxW1;
    }
    return rk;
}

__device__ __forceinline__ RoundKeys make_round_keys_dev(uint32_t k0, uint32_t k1) {
    RoundKeys rk;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        rk.k[r][0] = k0 + (uint32_t)r * kPhiloxW0;
        rk.k[r][1] = k1 + (uint32_t)r * kPhiloxW1;
    }
    return rk;
}

__device__ __forceinline__ u32x4 philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const RoundKeys& rk) {
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        const uint32_t lo0 = kPhiloxM0 * c0, hi0 = __umulhi(kPhiloxM0, c0);
        const uint32_t lo1 = kPhiloxM1 * c2, hi1 = __umulhi(kPhiloxM1, c2);
        const uint32_t n0 = hi1 ^ c1 ^ rk.k[r][0];
        const uint32_t n2 = hi0 ^ c3 ^ rk.k[r][1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u32x4{c0, c1, c2, c3};
}

// Two-sided threshold table in shared memory, indexed by the clamped d = 6v - sum_n + off:
//   s_tab[0][d + kc] (coin = down, m = -d) = thr(-d)
//   s_tab[1][d + kc] (coin = up,   m = +d) = thr(+d) | 0x80000000
// so that "u <= thr[clamp(m+3)]" becomes the single unsigned compare  r <= s_tab[coin][d + kc]
// (for coin = up bit 31 of r is set on both sides).  Same decisions as metropolis_delta.
__device__ __forceinline__ uint32_t build_two_sided(uint32_t* s_tab, const uint32_t* __restrict__ thr,
                                                    uint32_t n_thr) {
    const int kc = max((int)n_thr - 1, 3);
    const int width = 2 * kc + 1;
    for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) {
        const int up = i >= width;
        const int d = (up ? i - width : i) - kc;
        const int m = up ? d : -d;
        const int k = min(max(m + 3, 0), (int)n_thr - 1);
        s_tab[i] = thr[k] | (up ? 0x80000000u : 0u);
    }
    __syncthreads();
    return (uint32_t)kc;
}

__device__ __forceinline__ int metropolis_delta2(int v, int nsum, int off, uint32_t r,
                                                 const uint32_t* __restrict__ s_dn,
                                                 const uint32_t* __restrict__ s_up, int kc) {
    const int d = min(max(6 * v - nsum + off, -kc), kc);
    const bool up = (int)r < 0;
    const uint32_t t = (up ? s_up : s_dn)[d];
    return r <= t ? (up ? 1 : -1) : 0;
}

template <int G, int MARCH, bool WILSON, bool SINGLE>
__global__ void __launch_bounds__(kMarchWarps * 32)
k_sweep_march(const Geom g, const int colour, uint64_t sweep, const uint64_t* __restrict__ ctr,
              const uint32_t tl_first, const uint32_t n_planes, int32_t* __restrict__ own_base,
              const int32_t* __restrict__ oth_base, const ChainDev* __restrict__ chains,
              const uint32_t* __restrict__ thr_all, unsigned long long* __restrict__ acc,
              const __grid_constant__ RoundKeys rk_param) {
    constexpr uint32_t XH = 4u * G;      // elements per row per colour
    constexpr uint32_t R = 32u / G;      // rows handled side by side by one warp
    extern __shared__ uint32_t s_tab[];
    __shared__ unsigned int s_cnt;
    const uint32_t chain = blockIdx.z;
    const ChainDev cd = chains[chain];
    if (ctr) sweep += *ctr;
    if (threadIdx.x == 0) s_cnt = 0;
    const int kc = (int)build_two_sided(s_tab, thr_all + (size_t)chain * kThrStride, cd.n_thr);
    const uint32_t* s_dn = s_tab + kc;
    const uint32_t* s_up = s_tab + 3 * kc + 1;

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t sub = lane / G, kq = lane % G;
    const uint32_t pl = blockIdx.y * kMarchWarps + warp;   // plane within the launch range
    const uint32_t y0 = (blockIdx.x * R + sub) * MARCH;
    int accepted = 0;
    const bool active = pl < n_planes && y0 < g.Y;
    // lanes of one row group share y0, so whole groups are in or out together: the shuffles below
    // only name lanes of this mask
    const unsigned int lanes = __ballot_sync(0xffffffffu, active);
    if (active) {
        const uint32_t tl = tl_first + pl;
        const uint32_t tg = g.t0 + tl;
        const size_t cbase = (size_t)chain * g.chain_stride;
        const uint32_t rowq = y0 * G + kq;               // int4 index of the first row's quad in a plane
        const uint32_t planeq = g.plane / 4;
        int4* p_own = reinterpret_cast<int4*>(own_base + cbase) + (plane_self(g, tl) * planeq + rowq);
        const int4* othq = reinterpret_cast<const int4*>(oth_base + cbase);
        const int4* p_oc = othq + (plane_self(g, tl) * planeq + rowq);
        const int4* p_ta = othq + (plane_above(g, tl) * planeq + rowq);
        const int4* p_tb = othq + (plane_below(g, tl) * planeq + rowq);
        // rows y0-1 and y0+MARCH wrap around the plane
        const int4* p_first = (y0 == 0) ? p_oc + (g.Y - 1) * G : p_oc - G;
        const int4* p_last = (y0 + MARCH == g.Y) ? p_oc - y0 * G : p_oc + MARCH * G;
        const bool wil_plane = WILSON && tg < cd.wil_h;
        RoundKeys rk_local;
        if (!SINGLE) rk_local = make_round_keys_dev(cd.k0, cd.k1);
        const RoundKeys& rk = SINGLE ? rk_param : rk_local;
        int4 om = __ldg(p_first);
        int4 oc = __ldg(p_oc);
        uint32_t grp = (tg * g.Y + y0) * G + kq;     // Philox counter word 0 of this quad
        const uint32_t src_up = (lane & ~(G - 1u)) | ((lane + 1u) & (G - 1u));
        const uint32_t src_dn = (lane & ~(G - 1u)) | ((lane - 1u) & (G - 1u));
        const uint32_t p0 = (y0 + tg + (uint32_t)colour) & 1u;
#pragma unroll
        for (uint32_t s = 0; s < MARCH; ++s) {
            const uint32_t p = p0 ^ (s & 1u);
            const int4 v = p_own[s * G];
            const int4 on = __ldg(s + 1 < MARCH ? p_oc + (s + 1) * G : p_last);
            const int4 ta = __ldg(p_ta + s * G);
            const int4 tb = __ldg(p_tb + s * G);
            // x neighbour outside the quad: p == 1 -> element k0+4 (next lane's .x), else k0-1 (prev lane's .w)
            const int e = __shfl_sync(lanes, p ? oc.x : oc.w, p ? src_up : src_dn);
            int4 ns;
            ns.x = oc.x + (p ? oc.y : e);
            ns.y = oc.y + (p ? oc.z : oc.x);
            ns.z = oc.z + (p ? oc.w : oc.y);
            ns.w = oc.w + (p ? e : oc.z);
            ns.x += om.x + on.x; ns.y += om.y + on.y; ns.z += om.z + on.z; ns.w += om.w + on.w;
            ns.x += ta.x + tb.x; ns.y += ta.y + tb.y; ns.z += ta.z + tb.z; ns.w += ta.w + tb.w;
            const u32x4 r = philox4x32_rk(grp + s * G, (uint32_t)colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
            int4 off = make_int4(0, 0, 0, 0);
            if (WILSON) {
                const uint32_t y = y0 + s;
                if (wil_plane && (y == 0u || y == 1u % g.Y)) {
                    const int sgn = (y == 0u ? 1 : 0) - (y == 1u % g.Y ? 1 : 0);
                    const uint32_t x0 = 8u * kq + p;
                    off.x = x0 < cd.wil_w ? sgn : 0;
                    off.y = x0 + 2u < cd.wil_w ? sgn : 0;
                    off.z = x0 + 4u < cd.wil_w ? sgn : 0;
                    off.w = x0 + 6u < cd.wil_w ? sgn : 0;
                }
            }
            int4 dl;
            dl.x = metropolis_delta2(v.x, ns.x, off.x, r.x, s_dn, s_up, kc);
            dl.y = metropolis_delta2(v.y, ns.y, off.y, r.y, s_dn, s_up, kc);
            dl.z = metropolis_delta2(v.z, ns.z, off.z, r.z, s_dn, s_up, kc);
            dl.w = metropolis_delta2(v.w, ns.w, off.w, r.w, s_dn, s_up, kc);
            accepted += dl.x * dl.x + dl.y * dl.y + dl.z * dl.z + dl.w * dl.w;
            p_own[s * G] = make_int4(v.x + dl.x, v.y + dl.y, v.z + dl.z, v.w + dl.w);
            om = oc;
            oc = on;
        }
    }
    block_count_flush(&s_cnt, (unsigned int)accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

struct TiledPlan {
    bool enabled = false;
    int g = 0;       // quads per row
    int march = 0;   // rows per thread
};

// The marching kernel applies when X = 8 G with G a power of two <= 32, Y is a multiple of the
// march length and all element offsets / Philox group indices fit 32 bits.
inline cudaError_t tiled_prepare(TiledPlan& p, const Geom& g, uint32_t, int, bool force_off) {
    p.enabled = false;
    if (force_off) return cudaSuccess;
    const uint32_t G = g.X / 8;
    if (g.X % 8 != 0 || G == 0 || G > 32 || (G & (G - 1)) != 0) return cudaSuccess;
    if ((uint64_t)g.X * g.Y * g.T >= (1ull << 35)) return cudaSuccess;
    if (g.chain_stride >= (1ull << 31)) return cudaSuccess;
    int march = 0;
    for (int m : {8, 4, 2})
        if (g.Y % m == 0) { march = m; break; }
    if (!march) return cudaSuccess;
    p.g = (int)G;
    p.march = march;
    p.enabled = true;
    return cudaSuccess;
}
inline void tiled_release(TiledPlan&) {}

template <int G, int MARCH>
inline cudaError_t march_launch(const Geom& g, int colour, uint64_t sweep, const uint64_t* ctr,
                                uint32_t tl_first, uint32_t n_planes, uint32_t n_chains, int32_t* own,
                                const int32_t* oth, const ChainDev* chains, const uint32_t* thr,
                                unsigned long long* acc, uint32_t max_thr, bool wilson, const RoundKeys* single,
                                cudaStream_t s) {
    constexpr uint32_t R = 32u / G;
    dim3 grid((g.Y + R * MARCH - 1) / (R * MARCH), (n_planes + kMarchWarps - 1) / kMarchWarps, n_chains);
    const uint32_t kc = max_thr > 4 ? max_thr - 1 : 3;
    const size_t smem = sizeof(uint32_t) * 2 * (2 * kc + 1);
    const RoundKeys rk = single ? *single : RoundKeys{};
#define LQFT_GO(W, S) \
    k_sweep_march<G, MARCH, W, S><<<grid, kMarchWarps * 32, smem, s>>>(g, colour, sweep, ctr, tl_first, n_planes, own, \
                                                                       oth, chains, thr, acc, rk)
    if (single) {
        if (wilson) LQFT_GO(true, true); else LQFT_GO(false, true);
    } else {
        if (wilson) LQFT_GO(true, false); else LQFT_GO(false, false);
    }
#undef LQFT_GO
    return cudaGetLastError();
}

template <int G>
inline cudaError_t march_dispatch_m(int march, const Geom& g, int colour, uint64_t sweep, const uint64_t* ctr,
                                    uint32_t tl_first, uint32_t n_planes, uint32_t n_chains, int32_t* own,
                                    const int32_t* oth, const ChainDev* chains, const uint32_t* thr,
                                    unsigned long long* acc, uint32_t max_thr, bool wilson, const RoundKeys* single, cudaStream_t s) {
    switch (march) {
        case 8: return march_launch<G, 8>(g, colour, sweep, ctr, tl_first, n_planes, n_chains, own, oth, chains, thr, acc, max_thr, wilson, single, s);
        case 4: return march_launch<G, 4>(g, colour, sweep, ctr, tl_first, n_planes, n_chains, own, oth, chains, thr, acc, max_thr, wilson, single, s);
        default: return march_launch<G, 2>(g, colour, sweep, ctr, tl_first, n_planes, n_chains, own, oth, chains, thr, acc, max_thr, wilson, single, s);
    }
}

// One half-sweep of `colour` over local planes [tl_first, tl_first + n_planes).
inline cudaError_t tiled_half(const TiledPlan& p, const Geom& g, int colour, uint64_t sweep, const uint64_t* ctr,
                              uint32_t tl_first, uint32_t n_planes, uint32_t n_chains, int32_t* own,
                              const int32_t* oth, const ChainDev* chains, const uint32_t* thr,
                              unsigned long long* acc, uint32_t max_thr, bool wilson, const RoundKeys* single,
                              cudaStream_t s) {
#define LQFT_MARCH(GV) \
    case GV: return march_dispatch_m<GV>(p.march, g, colour, sweep, ctr, tl_first, n_planes, n_chains, own, oth, chains, thr, acc, max_thr, wilson, single, s)
    switch (p.g) {
        LQFT_MARCH(1);
        LQFT_MARCH(2);
        LQFT_MARCH(4);
        LQFT_MARCH(8);
        LQFT_MARCH(16);
        LQFT_MARCH(32);
        default: return cudaErrorInvalidValue;
    }
#undef LQFT_MARCH
}

}  // namespace lqft

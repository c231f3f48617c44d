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

// Tunables (overridable at build time for experiments; defaults are the measured best on B200).
#ifndef LQFT_MARCH_WARPS
#define LQFT_MARCH_WARPS 4
#endif
#ifndef LQFT_MARCH_LONG
#define LQFT_MARCH_LONG 8
#endif
#ifndef LQFT_MARCH_MINBLOCKS
#define LQFT_MARCH_MINBLOCKS 7
#endif
constexpr int kMarchWarps = LQFT_MARCH_WARPS;  // warps (= t-planes) per block
constexpr int kMarchLong = LQFT_MARCH_LONG;    // rows per thread when Y allows

// Programmatic dependent launch: a sweep kernel lets its successor start (and run its prologue) at
// once, and touches field data / the device sweep counter only after its predecessor has completed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Round keys of Philox4x32-10 (key + r * W), computed once per thread and kept in registers.
struct RoundKeys {
    uint32_t k[kPhiloxRounds][2];
};

__device__ __forceinline__ RoundKeys make_round_keys(uint32_t k0, uint32_t k1) {
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
        // one 32 x 32 -> 64 multiply per product (IMAD.WIDE); written as separate low / high halves the compiler
        // splits them into IMAD + IMAD.HI when registers are tight
        const unsigned long long p0 = (unsigned long long)kPhiloxM0 * c0, p1 = (unsigned long long)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k[r][0];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k[r][1];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    return u32x4{c0, c1, c2, c3};
}

// Two-sided decision table in shared memory, indexed by (coin, clamped d = 6v - sum_n + off):
//   down half, entry d + kc : { thr(m = -d),              -1 }
//   up   half, entry d + kc : { thr(m = +d) | 0x80000000, +1 }
// so "u <= thr[clamp(m+3)]" is the single unsigned compare  r <= entry.x  (for coin = up bit 31 of
// r is set on both sides) and the accepted step is entry.y.  Same decisions as metropolis_delta.
__device__ __forceinline__ int build_two_sided(uint2* s_tab, const uint32_t* __restrict__ thr, uint32_t n_thr) {
    const int kc = max((int)n_thr - 1, 3);
    const int width = 2 * kc + 1;
    for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) {
        const int up = i >= width;
        const int d = (up ? i - width : i) - kc;
        const int m = up ? d : -d;
        const int k = min(max(m + 3, 0), (int)n_thr - 1);
        s_tab[i] = make_uint2(thr[k] | (up ? 0x80000000u : 0u), up ? 1u : 0xffffffffu);
    }
    __syncthreads();
    return kc;
}

template <bool COUNT>
__device__ __forceinline__ int metropolis_update2(int v, int nsum, int off, uint32_t r,
                                                  const uint2* __restrict__ s_dn, const uint2* __restrict__ s_up,
                                                  int kc, int& accepted) {
    const int d = min(max(6 * v - nsum + off, -kc), kc);
    const uint2 t = ((int)r < 0 ? s_up : s_dn)[d];
    if (r <= t.x) {
        v += (int)t.y;
        if (COUNT) accepted += 1;
    }
    return v;
}

// MARCH rows of one quad column; P0 = x parity of the first row's sites (alternates row by row).
template <int G, int MARCH, bool WILSON, bool COUNT, uint32_t P0>
__device__ __forceinline__ void march_rows(int4* __restrict__ p_own, const int4* __restrict__ p_oc,
                                           const int4* __restrict__ p_ta, const int4* __restrict__ p_tb,
                                           const int4* __restrict__ p_first, const int4* __restrict__ p_last,
                                           const unsigned int lanes, const uint32_t src_up, const uint32_t src_dn,
                                           const uint32_t grp, const uint32_t colour, const uint64_t sweep,
                                           const RoundKeys& rk, const uint2* __restrict__ s_dn,
                                           const uint2* __restrict__ s_up, const int kc, const uint32_t y0,
                                           const uint32_t Y, const uint32_t kq, const bool wil_plane,
                                           const uint32_t wil_w, int& accepted) {
    int4 om = __ldg(p_first);
    int4 oc = __ldg(p_oc);
#pragma unroll
    for (uint32_t s = 0; s < MARCH; ++s) {
        constexpr uint32_t dummy = 0;
        (void)dummy;
        const bool p = ((P0 ^ s) & 1u) != 0u;  // compile time after unrolling
        const int4 v = p_own[s * G];
        const int4 on = __ldg(s + 1 < MARCH ? p_oc + (s + 1) * G : p_last);
        const int4 ta = __ldg(p_ta + s * G);
        const int4 tb = __ldg(p_tb + s * G);
        // x neighbour outside the quad: p -> element k0+4 (next lane's .x), else k0-1 (previous lane's .w)
        const int e = __shfl_sync(lanes, p ? oc.x : oc.w, p ? src_up : src_dn);
        int4 ns;
        ns.x = oc.x + (p ? oc.y : e);
        ns.y = oc.y + (p ? oc.z : oc.x);
        ns.z = oc.z + (p ? oc.w : oc.y);
        ns.w = oc.w + (p ? e : oc.z);
        ns.x += om.x + on.x; ns.y += om.y + on.y; ns.z += om.z + on.z; ns.w += om.w + on.w;
        ns.x += ta.x + tb.x; ns.y += ta.y + tb.y; ns.z += ta.z + tb.z; ns.w += ta.w + tb.w;
        const u32x4 r = philox4x32_rk(grp + s * G, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
        int4 off = make_int4(0, 0, 0, 0);
        if (WILSON) {
            const uint32_t y = y0 + s;
            if (wil_plane && (y == 0u || y == 1u % Y)) {
                const int sgn = (y == 0u ? 1 : 0) - (y == 1u % Y ? 1 : 0);
                const uint32_t x0 = 8u * kq + (p ? 1u : 0u);
                off.x = x0 < wil_w ? sgn : 0;
                off.y = x0 + 2u < wil_w ? sgn : 0;
                off.z = x0 + 4u < wil_w ? sgn : 0;
                off.w = x0 + 6u < wil_w ? sgn : 0;
            }
        }
        int4 w;
        w.x = metropolis_update2<COUNT>(v.x, ns.x, off.x, r.x, s_dn, s_up, kc, accepted);
        w.y = metropolis_update2<COUNT>(v.y, ns.y, off.y, r.y, s_dn, s_up, kc, accepted);
        w.z = metropolis_update2<COUNT>(v.z, ns.z, off.z, r.z, s_dn, s_up, kc, accepted);
        w.w = metropolis_update2<COUNT>(v.w, ns.w, off.w, r.w, s_dn, s_up, kc, accepted);
        p_own[s * G] = w;
        om = oc;
        oc = on;
    }
}

template <int G, int MARCH, bool WILSON, bool COUNT>
__global__ void __launch_bounds__(kMarchWarps * 32, LQFT_MARCH_MINBLOCKS)
k_sweep_march(const Geom g, const int colour, uint64_t sweep, const uint64_t* __restrict__ ctr,
              const uint32_t tl_first, const uint32_t n_planes, int32_t* __restrict__ own_base,
              const int32_t* __restrict__ oth_base, const ChainDev* __restrict__ chains,
              const uint32_t* __restrict__ thr_all, unsigned long long* __restrict__ acc) {
    constexpr uint32_t R = 32u / G;      // rows handled side by side by one warp
    extern __shared__ uint2 s_tab2[];
    __shared__ unsigned int s_cnt;
    pdl_launch_dependents();
    const uint32_t chain = blockIdx.z;
    const ChainDev cd = chains[chain];
    if (COUNT && threadIdx.x == 0) s_cnt = 0;
    const int kc = build_two_sided(s_tab2, thr_all + (size_t)chain * kThrStride, cd.n_thr);
    const uint2* s_dn = s_tab2 + kc;
    const uint2* s_up = s_tab2 + 3 * kc + 1;
    pdl_wait();
    if (ctr) sweep += *ctr;

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t sub = lane / G, kq = lane % G;
    const uint32_t pl = blockIdx.y * kMarchWarps + warp;   // plane within the launch range
    const uint32_t y0 = (blockIdx.x * R + sub) * MARCH;
    int accepted = 0;
    const bool active = pl < n_planes && y0 < g.Y;
    // lanes of one row group share y0, so whole groups are in or out together: the shuffles only
    // name lanes of this mask
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
        const RoundKeys rk = make_round_keys(cd.k0, cd.k1);
        const uint32_t grp = (tg * g.Y + y0) * G + kq;     // Philox counter word 0 of the first quad
        const uint32_t src_up = (lane & ~(G - 1u)) | ((lane + 1u) & (G - 1u));
        const uint32_t src_dn = (lane & ~(G - 1u)) | ((lane - 1u) & (G - 1u));
        // y0 is a multiple of the (even) march length: the parity is uniform over the warp
        if ((tg + (uint32_t)colour) & 1u)
            march_rows<G, MARCH, WILSON, COUNT, 1u>(p_own, p_oc, p_ta, p_tb, p_first, p_last, lanes, src_up, src_dn, grp,
                                                    (uint32_t)colour, sweep, rk, s_dn, s_up, kc, y0, g.Y, kq, wil_plane,
                                                    cd.wil_w, accepted);
        else
            march_rows<G, MARCH, WILSON, COUNT, 0u>(p_own, p_oc, p_ta, p_tb, p_first, p_last, lanes, src_up, src_dn, grp,
                                                    (uint32_t)colour, sweep, rk, s_dn, s_up, kc, y0, g.Y, kq, wil_plane,
                                                    cd.wil_w, accepted);
    }
    if (COUNT)
        block_count_flush(&s_cnt, (unsigned int)accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

struct TiledPlan {
    bool enabled = false;
    int g = 0;       // quads per row
    int march = 0;   // rows per thread
};

// The marching kernel applies when X = 8 G with G a power of two <= 32 and all element offsets and
// Philox group indices fit 32 bits.  (Y is even, so a march of 2 always divides it.)
inline cudaError_t tiled_prepare(TiledPlan& p, const Geom& g, uint32_t, int, bool force_off) {
    p.enabled = false;
    if (force_off) return cudaSuccess;
    const uint32_t G = g.X / 8;
    if (g.X % 8 != 0 || G == 0 || G > 32 || (G & (G - 1)) != 0) return cudaSuccess;
    if ((uint64_t)g.X * g.Y * g.T >= (1ull << 35)) return cudaSuccess;
    if (g.chain_stride >= (1ull << 31)) return cudaSuccess;
    p.g = (int)G;
    p.march = (g.Y % kMarchLong == 0) ? kMarchLong : 2;
    p.enabled = true;
    return cudaSuccess;
}
inline void tiled_release(TiledPlan&) {}

// peer-memory halo descriptors (used by the streaming kernel, kernels_stream.cuh)
struct HaloSync {  // lives in each rank's device memory, mapped into both neighbours
    unsigned long long from_lo;   // last phase published by the lower neighbour (rank - 1)
    unsigned long long from_hi;   // last phase published by the upper neighbour (rank + 1)
    unsigned int ticket;          // boundary blocks of the running launch that have finished
    unsigned int error;           // set when a flag wait timed out
};

struct HaloP2P {
    int32_t* lo_ghost = nullptr;  // lower neighbour's UPPER ghost plane of the colour being updated (chain 0)
    int32_t* hi_ghost = nullptr;  // upper neighbour's LOWER ghost plane
    HaloSync* lo_sync = nullptr;  // lower neighbour's HaloSync (we write its from_hi)
    HaloSync* hi_sync = nullptr;  // upper neighbour's HaloSync (we write its from_lo)
    HaloSync* mine = nullptr;
    unsigned long long phase = 0;     // published when this launch's boundary blocks are done
    unsigned long long wait_for = 0;  // ghost readers wait until the neighbour has published >= this
    uint32_t n_boundary_blocks = 0;
    uint32_t enabled = 0;
};

struct MarchArgs {
    Geom g;
    int colour;
    uint64_t sweep;
    const uint64_t* ctr;
    uint32_t tl_first, n_planes, n_chains;
    int32_t* own;
    const int32_t* oth;
    const ChainDev* chains;
    const uint32_t* thr;
    unsigned long long* acc;
    uint32_t max_thr;
    bool wilson, count;
    cudaStream_t stream;
    bool pdl;  // launch with programmatic stream serialization
    HaloP2P halo;  // peer-memory halo push (streaming kernel, world_size > 1)
};

template <int G, int MARCH>
inline cudaError_t march_launch(const MarchArgs& a) {
    constexpr uint32_t R = 32u / G;
    dim3 grid((a.g.Y + R * MARCH - 1) / (R * MARCH), (a.n_planes + kMarchWarps - 1) / kMarchWarps, a.n_chains);
    const uint32_t kc = a.max_thr > 4 ? a.max_thr - 1 : 3;
    const size_t smem = sizeof(uint2) * 2 * (2 * kc + 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kMarchWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
#define LQFT_GO(W, C)                                                                                              \
    return cudaLaunchKernelEx(&cfg, k_sweep_march<G, MARCH, W, C>, a.g, a.colour, a.sweep, a.ctr, a.tl_first, a.n_planes, \
                              a.own, a.oth, a.chains, a.thr, a.acc)
    if (a.wilson) {
        if (a.count) LQFT_GO(true, true); else LQFT_GO(true, false);
    } else {
        if (a.count) LQFT_GO(false, true); else LQFT_GO(false, false);
    }
#undef LQFT_GO
    return cudaGetLastError();
}

// One half-sweep of `colour` over local planes [tl_first, tl_first + n_planes).
inline cudaError_t tiled_half(const TiledPlan& p, const MarchArgs& a) {
#define LQFT_MARCH(GV) \
    case GV: return p.march == kMarchLong ? march_launch<GV, kMarchLong>(a) : march_launch<GV, 2>(a)
    switch (p.g) {
#ifndef LQFT_G32_ONLY
        LQFT_MARCH(1);
        LQFT_MARCH(2);
        LQFT_MARCH(4);
        LQFT_MARCH(8);
        LQFT_MARCH(16);
#endif
        LQFT_MARCH(32);
        default: return cudaErrorInvalidValue;
    }
#undef LQFT_MARCH
}

}  // namespace lqft

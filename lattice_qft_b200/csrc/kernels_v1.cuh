// kernels_v1.cuh — two-pass red-black kernels on the colour-split layout (global loads,
// L1/L2 reuse only).  They are the general path (any even extents, any slab, i32 heights) and the
// cross-check for the 16-bit kernels of kernels_narrow.cuh / kernels_cluster.cuh.
//
// Device layout (per chain, per colour c in {0,1}):
//     field[c][chain][plane p][row y][k]      k < Xh = X/2, element = site x = 2k + ((y+t+c)&1)
// i.e. position hi = idx >> 1 of the lexicographic index idx = x + X*(y + Y*t)
// (lattice.rs:93-104).  With world_size > 1 each rank stores its Tl owned planes between two
// ghost planes (ghost = 1); with one rank t wraps by index (ghost = 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spec.cuh"

namespace lqft {

struct Geom {
    uint32_t X, Y, T, Xh;   // global extents, Xh = X/2
    uint32_t Tl, t0;        // owned planes, first owned global plane
    uint32_t ghost;         // ghost planes per side (0: periodic wrap by index; slabs: 1, or the deep halo of the narrow copy)
    uint32_t plane;         // Xh*Y elements per plane per colour
    uint32_t gbuf;          // 16-bit copy on slabs: which of its two sets of ghost planes is current (kernels_narrow.cuh)
    uint32_t entry_bytes;   // 8 = sizeof(uint2): stride of the 16-bit kernels' decision table, handed over as data (n16_lookup)
    uint64_t chain_stride;  // elements per chain per colour: plane * (Tl + 2*ghost); the 16-bit copy on slabs: Tl + 4*ghost
};

struct ChainDev {
    uint32_t k0, k1;   // Philox key = seed lo/hi
    uint32_t n_thr;    // entries of the threshold table
    uint32_t wil_w, wil_h;
    uint32_t pad0, pad1, pad2;
};

constexpr uint32_t kThrStride = 2048;  // == LQFT_MAX_THRESHOLDS
constexpr int kAccSlots = 64;          // spread of the acceptance counters per chain

// Storage plane of local plane tl.  Negative numbers (wrapped) are the lower ghost planes, numbers >= Tl the upper
// ones.  Layout per chain: [lower ghosts][owned][upper ghosts], and for the 16-bit copy on slabs a second set of
// ghost planes behind it, [lower 1][upper 1]; gbuf says which set is current (the other one is being filled by the
// neighbours, kernels_narrow.cuh).
__device__ __forceinline__ uint32_t plane_self(const Geom& g, uint32_t tl) {
    if (g.gbuf && tl >= g.Tl) return ((int32_t)tl < 0 ? g.Tl : 0u) + 3u * g.ghost + tl;
    return tl + g.ghost;
}
__device__ __forceinline__ uint32_t plane_below(const Geom& g, uint32_t tl) {
    return g.ghost ? plane_self(g, tl - 1u) : (tl == 0 ? g.Tl - 1 : tl - 1);  // storage plane index of t-1
}
__device__ __forceinline__ uint32_t plane_above(const Geom& g, uint32_t tl) {
    return g.ghost ? plane_self(g, tl + 1u) : (tl + 1 == g.Tl ? 0 : tl + 1);  // storage plane index of t+1
}

// The Metropolis decision of metropolis_single (algorithm.rs:137-157) in integers:
// proposal v -> v + s, s = +-1 from the coin; new - old = 6 + 2 s (6v - sum_n + off);
// accept iff draw <= exp((old-new)*beta)  <=>  u <= thr[clamp(m+3)].
__device__ __forceinline__ int metropolis_delta(int v, int nsum, int off, uint32_t r,
                                                const uint32_t* __restrict__ s_thr, int kmax) {
    int d = 6 * v - nsum + off;
    bool up = (r >> 31) != 0u;
    int m = up ? d : -d;
    int k = min(max(m + 3, 0), kmax);
    uint32_t u = r & 0x7fffffffu;
    bool acc = u <= s_thr[k];
    return acc ? (up ? 1 : -1) : 0;
}

__device__ __forceinline__ void load_thresholds(uint32_t* s_thr, const uint32_t* __restrict__ thr,
                                                uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
}

__device__ __forceinline__ void block_count_flush(unsigned int* s_cnt, unsigned int mine,
                                                  unsigned long long* acc_slot) {
    unsigned int w = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31u) == 0u && w) atomicAdd(s_cnt, w);
    __syncthreads();
    if (threadIdx.x == 0 && *s_cnt) atomicAdd(acc_slot, (unsigned long long)*s_cnt);
}

// ---- generic half-sweep: one thread per site of colour `colour`, any even extents --------
// grid = (ceil(plane/blockDim), n_planes, n_chains); plane range [tl_first, tl_first + gridDim.y)
template <bool WILSON>
__global__ void __launch_bounds__(256)
k_sweep_generic(Geom g, int colour, uint64_t sweep, const uint64_t* __restrict__ ctr,
                uint32_t tl_first, int32_t* __restrict__ own_base,
                const int32_t* __restrict__ oth_base, const ChainDev* __restrict__ chains,
                const uint32_t* __restrict__ thr_all, unsigned long long* __restrict__ acc) {
    extern __shared__ uint32_t s_thr[];
    __shared__ unsigned int s_cnt;
    const uint32_t chain = blockIdx.z;
    const ChainDev cd = chains[chain];
    if (ctr) sweep += *ctr;  // graph replays keep the running sweep counter on the device
    if (threadIdx.x == 0) s_cnt = 0;
    load_thresholds(s_thr, thr_all + (size_t)chain * kThrStride, cd.n_thr);

    const uint32_t tl = tl_first + blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int accepted = 0;
    if (i < g.plane) {
        const uint32_t y = i / g.Xh, k = i - y * g.Xh;
        const uint32_t tg = g.t0 + tl;
        const uint32_t p = (y + tg + (uint32_t)colour) & 1u;
        int32_t* own = own_base + (size_t)chain * g.chain_stride;
        const int32_t* oth = oth_base + (size_t)chain * g.chain_stride;
        const size_t ps = (size_t)plane_self(g, tl) * g.plane;
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
        const uint32_t kx = p ? (k + 1 == g.Xh ? 0 : k + 1) : (k == 0 ? g.Xh - 1 : k - 1);
        const int v = own[row + k];
        int nsum = oth[row + k] + oth[row + kx];
        nsum += oth[ps + (size_t)yp * g.Xh + k] + oth[ps + (size_t)ym * g.Xh + k];
        nsum += oth[(size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k];
        nsum += oth[(size_t)plane_below(g, tl) * g.plane + (size_t)y * g.Xh + k];
        const uint64_t hi = ((uint64_t)tg * g.Y + y) * g.Xh + k;  // global idx >> 1
        const uint64_t grp = hi >> 2;
        u32x4 r = philox4x32((uint32_t)grp, (uint32_t)colour | ((uint32_t)(grp >> 32) << 8),
                             (uint32_t)sweep, (uint32_t)(sweep >> 32), cd.k0, cd.k1);
        int off = 0;
        if (WILSON) off = wilson_offset(g.Y, 2 * k + p, y, tg, cd.wil_w, cd.wil_h);
        int dlt = metropolis_delta(v, nsum, off, word_of(r, (uint32_t)(hi & 3u)), s_thr,
                                   (int)cd.n_thr - 1);
        if (dlt) {
            own[row + k] = v + dlt;
            accepted = 1;
        }
    }
    block_count_flush(&s_cnt, accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

// ---- vectorised half-sweep: one thread per 4 consecutive same-colour sites (X % 8 == 0) ---
// One Philox call feeds the 4 sites (they are exactly one counter group).
template <bool WILSON>
__global__ void __launch_bounds__(256)
k_sweep_vec4(Geom g, int colour, uint64_t sweep, const uint64_t* __restrict__ ctr, uint32_t tl_first,
             int32_t* __restrict__ own_base,
             const int32_t* __restrict__ oth_base, const ChainDev* __restrict__ chains,
             const uint32_t* __restrict__ thr_all, unsigned long long* __restrict__ acc) {
    extern __shared__ uint32_t s_thr[];
    __shared__ unsigned int s_cnt;
    const uint32_t chain = blockIdx.z;
    const ChainDev cd = chains[chain];
    if (ctr) sweep += *ctr;  // graph replays keep the running sweep counter on the device
    if (threadIdx.x == 0) s_cnt = 0;
    load_thresholds(s_thr, thr_all + (size_t)chain * kThrStride, cd.n_thr);

    const uint32_t tl = tl_first + blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t Qx = g.Xh >> 2;
    unsigned int accepted = 0;
    if (i < (g.plane >> 2)) {
        const uint32_t y = i / Qx, kq = i - y * Qx, k0 = kq << 2;
        const uint32_t tg = g.t0 + tl;
        const uint32_t p = (y + tg + (uint32_t)colour) & 1u;
        int32_t* own = own_base + (size_t)chain * g.chain_stride;
        const int32_t* oth = oth_base + (size_t)chain * g.chain_stride;
        const size_t ps = (size_t)plane_self(g, tl) * g.plane;
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
        const uint32_t kx = p ? (k0 + 4 == g.Xh ? 0 : k0 + 4) : (k0 == 0 ? g.Xh - 1 : k0 - 1);
        const int4 v = *reinterpret_cast<const int4*>(own + row + k0);
        const int4 o = *reinterpret_cast<const int4*>(oth + row + k0);
        const int e = oth[row + kx];
        const int4 a = *reinterpret_cast<const int4*>(oth + ps + (size_t)yp * g.Xh + k0);
        const int4 b = *reinterpret_cast<const int4*>(oth + ps + (size_t)ym * g.Xh + k0);
        const int4 c = *reinterpret_cast<const int4*>(
            oth + (size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k0);
        const int4 d = *reinterpret_cast<const int4*>(
            oth + (size_t)plane_below(g, tl) * g.plane + (size_t)y * g.Xh + k0);
        // x neighbours: p == 0: site j sits at x = 2(k0+j), nbrs o[j-1], o[j];  p == 1: o[j], o[j+1]
        int4 ns;
        ns.x = o.x + (p ? o.y : e);
        ns.y = o.y + (p ? o.z : o.x);
        ns.z = o.z + (p ? o.w : o.y);
        ns.w = o.w + (p ? e : o.z);
        ns.x += a.x + b.x + c.x + d.x;
        ns.y += a.y + b.y + c.y + d.y;
        ns.z += a.z + b.z + c.z + d.z;
        ns.w += a.w + b.w + c.w + d.w;
        const uint64_t grp = ((uint64_t)tg * g.Y + y) * Qx + kq;
        u32x4 r = philox4x32((uint32_t)grp, (uint32_t)colour | ((uint32_t)(grp >> 32) << 8),
                             (uint32_t)sweep, (uint32_t)(sweep >> 32), cd.k0, cd.k1);
        int4 off = make_int4(0, 0, 0, 0);
        if (WILSON) {
            const uint32_t x0 = 2 * k0 + p;
            off.x = wilson_offset(g.Y, x0, y, tg, cd.wil_w, cd.wil_h);
            off.y = wilson_offset(g.Y, x0 + 2, y, tg, cd.wil_w, cd.wil_h);
            off.z = wilson_offset(g.Y, x0 + 4, y, tg, cd.wil_w, cd.wil_h);
            off.w = wilson_offset(g.Y, x0 + 6, y, tg, cd.wil_w, cd.wil_h);
        }
        const int kmax = (int)cd.n_thr - 1;
        int4 dl;
        dl.x = metropolis_delta(v.x, ns.x, off.x, r.x, s_thr, kmax);
        dl.y = metropolis_delta(v.y, ns.y, off.y, r.y, s_thr, kmax);
        dl.z = metropolis_delta(v.z, ns.z, off.z, r.z, s_thr, kmax);
        dl.w = metropolis_delta(v.w, ns.w, off.w, r.w, s_thr, kmax);
        accepted = (dl.x != 0) + (dl.y != 0) + (dl.z != 0) + (dl.w != 0);
        if (accepted) {
            *reinterpret_cast<int4*>(own + row + k0) =
                make_int4(v.x + dl.x, v.y + dl.y, v.z + dl.z, v.w + dl.w);
        }
    }
    block_count_flush(&s_cnt, accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

// ---- layout conversion: lexicographic host order <-> colour split -------------------------
// One thread per x-pair (x = 2k, 2k+1) of an owned plane.  grid = (ceil(plane/bd), Tl, 1).
// H = int32_t: the i32 arrays.  H = uint16_t: the 16-bit arrays (kernels_narrow.cuh), stored = value - centre + bias
// with *centre the chain's centre; k_split then also reports the largest |value - centre| it met (*worst, atomicMax;
// capped at 2^24), from which the caller decides whether the 16-bit copy can hold the chain at all.
template <typename H>
__global__ void k_split(Geom g, const int32_t* __restrict__ lex, H* __restrict__ c0, H* __restrict__ c1,
                        const int32_t* __restrict__ centre, int32_t bias, uint32_t* __restrict__ worst) {
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t far = 0;
    if (i < g.plane) {
        const uint32_t y = i / g.Xh;
        const uint32_t q = (y + g.t0 + tl) & 1u;  // colour of the even-x site of this pair
        const size_t src = ((size_t)tl * g.plane + i);
        const int2 v = reinterpret_cast<const int2*>(lex)[src];
        const size_t dst = (size_t)plane_self(g, tl) * g.plane + i;
        if constexpr (sizeof(H) == 2) {
            const long long dx = (long long)v.x - *centre, dy = (long long)v.y - *centre;
            far = (uint32_t)min(max(llabs(dx), llabs(dy)), (long long)(1 << 24));
            (q ? c1 : c0)[dst] = (H)(dx + bias);
            (q ? c0 : c1)[dst] = (H)(dy + bias);
        } else {
            (q ? c1 : c0)[dst] = v.x;
            (q ? c0 : c1)[dst] = v.y;
        }
    }
    if constexpr (sizeof(H) == 2) {
        far = __reduce_max_sync(0xffffffffu, far);
        if ((threadIdx.x & 31u) == 0u && far) atomicMax(worst, far);
    }
}

template <typename H>
__global__ void k_merge(Geom g, int32_t* __restrict__ lex, const H* __restrict__ c0, const H* __restrict__ c1,
                        const int32_t* __restrict__ offset, const int32_t* __restrict__ centre, int32_t bias) {
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const uint32_t y = i / g.Xh;
    const uint32_t q = (y + g.t0 + tl) & 1u;
    const size_t src = (size_t)plane_self(g, tl) * g.plane + i;
    // true height = stored - offset (pending normalisation shift of this chain, see k_pick_shift)
    const int32_t off = *offset - (centre ? *centre - bias : 0);
    int2 v;
    v.x = (int32_t)(q ? c1 : c0)[src] - off;
    v.y = (int32_t)(q ? c0 : c1)[src] - off;
    reinterpret_cast<int2*>(lex)[(size_t)tl * g.plane + i] = v;
}

// Hot start (computation.rs:225-226): heights uniform in [-128,127], keyed by global index; `bias` is added to what
// is stored (the 16-bit arrays with centre 0).  grid = (ceil(plane/bd), Tl, n_chains)
template <typename H>
__global__ void k_init_hot(Geom g, H* __restrict__ c0_base, H* __restrict__ c1_base, const ChainDev* __restrict__ chains,
                           int32_t bias) {
    const uint32_t chain = blockIdx.z, tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const ChainDev cd = chains[chain];
    const uint32_t y = i / g.Xh;
    const uint32_t tg = g.t0 + tl;
    const uint32_t q = (y + tg) & 1u;
    const uint64_t idx = 2 * ((uint64_t)tg * g.plane + i);  // global index of the even-x site
    const int v0 = hot_value(cd.k0, cd.k1, idx), v1 = hot_value(cd.k0, cd.k1, idx + 1);
    H* c0 = c0_base + (size_t)chain * g.chain_stride;
    H* c1 = c1_base + (size_t)chain * g.chain_stride;
    const size_t dst = (size_t)plane_self(g, tl) * g.plane + i;
    (q ? c1 : c0)[dst] = (H)(v0 + bias);
    (q ? c0 : c1)[dst] = (H)(v1 + bias);
}

// ---- normalize (fields.rs:616-633) ----------------------------------------------------------
// The action, every observable and every Metropolis decision depend on height differences only, so
// shift_values never has to touch the field: each chain carries an offset, true height = stored - offset,
// and normalize_random (subtract the TRUE height of the picked site) simply sets offset := stored[site].
// The offset is applied where absolute heights leave the device (k_merge, lqft_slice_moments).
// Step 1: one thread per chain reads the stored height of the picked site (0 if another rank owns it).
// H = uint16_t reads the 16-bit storage: stored = biased - bias + centre[chain] (`rebase` = centre, bias).
template <typename H>
__global__ void k_pick_shift(Geom g, const H* __restrict__ c0_base,
                             const H* __restrict__ c1_base,
                             const ChainDev* __restrict__ chains, uint32_t n_chains,
                             uint64_t counter, const uint64_t* __restrict__ ctr,
                             int64_t explicit_site, int32_t only_chain,
                             int32_t* __restrict__ shift, const int32_t* __restrict__ centre = nullptr, int32_t bias = 0) {
    const uint32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    if (only_chain >= 0 && (int32_t)chain != only_chain) {
        shift[chain] = 0;
        return;
    }
    const ChainDev cd = chains[chain];
    if (ctr) counter += *ctr;
    const uint64_t n = (uint64_t)g.X * g.Y * g.T;
    const uint64_t site = explicit_site >= 0 ? (uint64_t)explicit_site
                                             : pick_site(cd.k0, cd.k1, 2u, counter, 0u, n);
    const uint64_t hi = site >> 1;
    const uint32_t tg = (uint32_t)(hi / g.plane);
    const uint32_t i = (uint32_t)(hi - (uint64_t)tg * g.plane);
    int32_t s = 0;
    if (tg >= g.t0 && tg < g.t0 + g.Tl) {
        const uint32_t y = i / g.Xh;
        const uint32_t x = 2 * (i - y * g.Xh) + (uint32_t)(site & 1u);
        const uint32_t col = (x + y + tg) & 1u;
        const H* f = (col ? c1_base : c0_base) + (size_t)chain * g.chain_stride;
        s = (int32_t)f[(size_t)plane_self(g, tg - g.t0) * g.plane + i];
        if (centre) s += centre[chain] - bias;
    }
    shift[chain] = s;
}

// Step 2: offset[chain] := picked[chain] for the chains being normalised (after the all-reduce that hands the
// owner's value to every rank).
__global__ void k_set_offset(const int32_t* __restrict__ picked, int32_t* __restrict__ offset, uint32_t n_chains,
                             int32_t only_chain) {
    const uint32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    if (only_chain >= 0 && (int32_t)chain != only_chain) return;
    offset[chain] = picked[chain];
}
__global__ void k_add_offset(int32_t* __restrict__ offset, uint32_t chain, int32_t shift) { offset[chain] += shift; }

// ---- observables ---------------------------------------------------------------------------
// Block-level exact int64 sums, one atomic per block per quantity.
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NQ>
__device__ __forceinline__ void block_sum_flush(long long (&q)[NQ], unsigned long long* const (&dst)[NQ]) {
    __shared__ long long s_part[NQ][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        long long w = warp_sum_ll(q[j]);
        if (lane == 0) s_part[j][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        long long tot = 0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) tot += s_part[threadIdx.x][w];
        if (tot != 0 && dst[threadIdx.x]) atomicAdd(dst[threadIdx.x], (unsigned long long)tot);
    }
}

// Energy (fields.rs:728-737), action (fields.rs:719-726) and per-slice moments in one pass.
// One thread per hi position; it owns the colour-0 and the colour-1 site stored there.
// grid = (ceil(plane/bd), Tl, n_chains);  obs[chain][0]=E, [1]=S;  mom[chain][tg][0]=S1,[1]=S2.
__global__ void __launch_bounds__(256)
k_observe(Geom g, const int32_t* __restrict__ c0_base, const int32_t* __restrict__ c1_base,
          unsigned long long* __restrict__ obs, unsigned long long* __restrict__ mom) {
    const uint32_t chain = blockIdx.z, tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    long long q[4] = {0, 0, 0, 0};
    const uint32_t tg = g.t0 + tl;
    if (i < g.plane) {
        const int32_t* c0 = c0_base + (size_t)chain * g.chain_stride;
        const int32_t* c1 = c1_base + (size_t)chain * g.chain_stride;
        const uint32_t y = i / g.Xh, k = i - y * g.Xh;
        const uint32_t p0 = (y + tg) & 1u;  // x parity of the colour-0 site in this row
        const size_t ps = (size_t)plane_self(g, tl) * g.plane;
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1;
        const uint32_t kp = (k + 1 == g.Xh) ? 0 : k + 1;
        const size_t up = (size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k;
        const size_t yr = ps + (size_t)yp * g.Xh + k;
        const long long a = c0[row + k], b = c1[row + k];
        // +x neighbour: the site with odd x (parity 1) looks at position k+1 of the other colour
        const long long ax = p0 ? c1[row + kp] : b;
        const long long bx = p0 ? a : c0[row + kp];
        const long long ay = c1[yr], by = c0[yr];
        const long long at = c1[up], bt = c0[up];
        const long long dx = (a - ax) * (a - ax) + (b - bx) * (b - bx);
        const long long dy = (a - ay) * (a - ay) + (b - by) * (b - by);
        const long long dt = (a - at) * (a - at) + (b - bt) * (b - bt);
        q[0] = dx + dy - dt;
        q[1] = dx + dy + dt;
        q[2] = a + b;
        q[3] = a * a + b * b;
    }
    unsigned long long* const dst[4] = {
        obs + (size_t)chain * 2, obs + (size_t)chain * 2 + 1,
        mom ? mom + ((size_t)chain * g.T + tg) * 2 : nullptr,
        mom ? mom + ((size_t)chain * g.T + tg) * 2 + 1 : nullptr};
    block_sum_flush<4>(q, dst);
}

// Vectorised form of k_observe for X % 8 == 0: one thread per quad position owns four colour-0 and four
// colour-1 sites (6 LDG.128 + 1 LDG for 8 sites).  Same exact int64 sums.
// grid = (ceil(plane/4/bd), Tl, n_chains)
// H = int32_t, or uint16_t for the 16-bit storage (kernels_narrow.cuh): the bias cancels in every
// difference, so the energy sums are the same; slice moments need true heights and are i32 only.
__device__ __forceinline__ int4 load_h4(const int32_t* p) { return *reinterpret_cast<const int4*>(p); }
__device__ __forceinline__ int4 load_h4(const uint16_t* p) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    return make_int4((int)(v.x & 0xffffu), (int)(v.x >> 16), (int)(v.y & 0xffffu), (int)(v.y >> 16));
}


// Append the accumulated energy of every chain to its log and clear the accumulators.
// log[chain * cap + *n_meas] = E ; one thread per chain; thread 0 bumps *n_meas afterwards.
__global__ void k_energy_log(unsigned long long* __restrict__ obs, long long* __restrict__ log,
                             uint32_t cap, uint32_t n_chains, uint32_t* __restrict__ n_meas) {
    const uint32_t m = *n_meas;
    for (uint32_t c = threadIdx.x; c < n_chains; c += blockDim.x) {
        if (m < cap) log[(size_t)c * cap + m] = (long long)obs[(size_t)c * 2];
        obs[(size_t)c * 2] = 0;
        obs[(size_t)c * 2 + 1] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) *n_meas = m + 1;
}

// Advance the device-resident sweep counter at the end of a captured cycle.
__global__ void k_tick(uint64_t* ctr, uint64_t by, unsigned long long* phase = nullptr, unsigned long long phases = 0) {
    *ctr += by;
    if (phase) *phase += phases;
}

// Heights of a list of global sites (0 where another rank owns the site).
__global__ void k_gather_sites(Geom g, const int32_t* __restrict__ c0, const int32_t* __restrict__ c1,
                               const uint64_t* __restrict__ sites, uint32_t n,
                               long long* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t site = sites[j];
    const uint64_t hi = site >> 1;
    const uint32_t tg = (uint32_t)(hi / g.plane);
    const uint32_t i = (uint32_t)(hi - (uint64_t)tg * g.plane);
    long long v = 0;
    if (tg >= g.t0 && tg < g.t0 + g.Tl) {
        const uint32_t y = i / g.Xh;
        const uint32_t x = 2 * (i - y * g.Xh) + (uint32_t)(site & 1u);
        const uint32_t col = (x + y + tg) & 1u;
        v = (col ? c1 : c0)[(size_t)plane_self(g, tg - g.t0) * g.plane + i];
    }
    out[j] = v;
}

// DifferencePlot (outputdata.rs:241-276, kahan.rs:100-107): per-bond running sum of
// bond_difference in the 3 positive directions, Kahan-compensated like WelfordsAlgorithm64.sum.
// sum/comp are [3][Tl*plane*2] in lexicographic slab order.  grid = (ceil(plane/bd), Tl, 1)
__global__ void k_difference_update(Geom g, const int32_t* __restrict__ c0,
                                    const int32_t* __restrict__ c1, double* __restrict__ sum,
                                    double* __restrict__ comp) {
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const uint32_t tg = g.t0 + tl;
    const uint32_t y = i / g.Xh, k = i - y * g.Xh;
    const uint32_t p0 = (y + tg) & 1u;
    const size_t ps = (size_t)plane_self(g, tl) * g.plane;
    const size_t row = ps + (size_t)y * g.Xh;
    const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1;
    const uint32_t kp = (k + 1 == g.Xh) ? 0 : k + 1;
    const size_t up = (size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k;
    const size_t yr = ps + (size_t)yp * g.Xh + k;
    const int a = c0[row + k], b = c1[row + k];
    const int ax = p0 ? c1[row + kp] : b, bx = p0 ? a : c0[row + kp];
    const int ay = c1[yr], by = c0[yr], at = c1[up], bt = c0[up];
    // lexicographic positions of the two sites: even x first
    const size_t n_slab = (size_t)g.Tl * g.plane * 2;
    const size_t lex_even = ((size_t)tl * g.plane + i) * 2, lex_odd = lex_even + 1;
    const size_t la = p0 ? lex_odd : lex_even, lb = p0 ? lex_even : lex_odd;
    const double d[3][2] = {{(double)(a - ax), (double)(b - bx)},
                            {(double)(a - ay), (double)(b - by)},
                            {(double)(a - at), (double)(b - bt)}};
#pragma unroll
    for (int dir = 0; dir < 3; ++dir) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const size_t o = (size_t)dir * n_slab + (s ? lb : la);
            // KahanSummation::add (kahan.rs:21-31)
            const double yv = d[dir][s] - comp[o];
            const double tv = sum[o] + yv;
            comp[o] = (tv - sum[o]) - yv;
            sum[o] = tv;
        }
    }
}

}  // namespace lqft

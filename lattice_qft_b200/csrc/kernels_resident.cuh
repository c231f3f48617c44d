// kernels_resident.cuh — small lattices: a whole run inside one kernel launch.
//
// The reference's own runs use 16^3 ... 36^3 lattices (src/bin/sim.rs:27-31, 410 600 sweeps per chain).  A
// field of 16 KiB ... 182 KiB fits in one SM's shared memory, and at that size the multi-launch path is
// bound by launch latency (~4 us per half-sweep for 2 048 site updates).  Here one CTA owns one chain: it
// loads both colours into shared memory once and executes the loop body of Simulation::run
// (computation.rs:229-270) — sweep, energy when due, normalize when due — for all requested steps, with two
// block barriers per sweep and no other synchronisation.  Chains of a batch run on different SMs.
// Same sites, same Philox counters, same decisions as every other path: bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_common.cuh"

namespace lqft {

constexpr int kResidentThreads = 1024;
constexpr int kResidentMaxGroups = 6;  // counter groups (4 sites) per thread and colour: up to 36^3

struct ResidentArgs {
    Geom g;
    uint64_t sweep_base, first_step, n_steps;
    uint32_t energy_every, normalize_every;
    uint32_t elog_cap;      // measurements per chain the log can hold
    uint32_t count;         // accumulate acceptance counts
};

__device__ __forceinline__ long long block_sum_ll(long long v, long long* s_red) {
    v = warp_sum_ll(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // s_red may still be read from a previous reduction
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    long long tot = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) tot += s_red[w];
    return tot;
}

// grid = n_chains, block = kResidentThreads, dynamic smem = 2 * (N/2) * 4 + table.
template <bool WILSON, int VEC>
__global__ void __launch_bounds__(kResidentThreads, 1)
k_run_resident(const ResidentArgs a, int32_t* __restrict__ c0_base, int32_t* __restrict__ c1_base,
               const ChainDev* __restrict__ chains, const uint32_t* __restrict__ thr_all,
               int32_t* __restrict__ offset, long long* __restrict__ elog, unsigned long long* __restrict__ acc) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ long long s_red[32];
    __shared__ int s_pick;
    const Geom& g = a.g;
    const uint32_t chain = blockIdx.x;
    const ChainDev cd = chains[chain];
    const uint32_t half = g.plane * g.T;  // sites per colour
    int32_t* s_c[2] = {reinterpret_cast<int32_t*>(s_raw), reinterpret_cast<int32_t*>(s_raw) + half};
    uint2* s_tab2 = reinterpret_cast<uint2*>(s_raw + (size_t)2 * half * sizeof(int32_t));
    int32_t* gc[2] = {c0_base + (size_t)chain * g.chain_stride, c1_base + (size_t)chain * g.chain_stride};
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
        s_c[0][i] = gc[0][i];
        s_c[1][i] = gc[1][i];
    }
    const int kc = build_two_sided(s_tab2, thr_all + (size_t)chain * kThrStride, cd.n_thr);  // ends with a barrier
    const uint2* s_dn = s_tab2 + kc;
    const uint2* s_up = s_tab2 + 3 * kc + 1;
    const RoundKeys rk = make_round_keys(cd.k0, cd.k1);

    // the thread's counter groups (4 consecutive positions of a colour array) and the coordinates of their
    // first position: fixed for the whole run
    const uint32_t n_groups = half >> 2;  // half % 4 == 0 is checked by the host
    uint32_t gk[kResidentMaxGroups], gy[kResidentMaxGroups], gt[kResidentMaxGroups];
#pragma unroll
    for (int j = 0; j < kResidentMaxGroups; ++j) {
        const uint32_t grp = threadIdx.x + (uint32_t)j * blockDim.x;
        const uint32_t hi = grp << 2;
        const uint32_t row = hi / g.Xh;
        gk[j] = hi - row * g.Xh;
        gt[j] = row / g.Y;
        gy[j] = row - gt[j] * g.Y;
    }
    int accepted = 0;
    int32_t off_true = offset[chain];  // true height = stored - offset (lazy normalisation)
    uint32_t n_meas = 0;
    const uint64_t n_sites = (uint64_t)half * 2;

    for (uint64_t step = a.first_step; step < a.first_step + a.n_steps; ++step) {
        const uint64_t sweep = a.sweep_base + step;
#pragma unroll 1
        for (uint32_t colour = 0; colour < 2; ++colour) {
            int32_t* own = s_c[colour];
            const int32_t* oth = s_c[colour ^ 1];
#pragma unroll
            for (int j = 0; j < kResidentMaxGroups; ++j) {
                const uint32_t grp = threadIdx.x + (uint32_t)j * blockDim.x;
                if (grp < n_groups) {
                    const u32x4 r = philox4x32_rk(grp, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
                    if (VEC == 4) {
                        // Xh % 4 == 0: the group is one aligned quad of a row.  128-bit shared-memory accesses (a
                        // scalar walk over 4 consecutive positions would be a 4-way bank conflict on every load).
                        const uint32_t k0 = gk[j], y = gy[j], t = gt[j];
                        const uint32_t p = (y + t + colour) & 1u;
                        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
                        const uint32_t tp = (t + 1 == g.T) ? 0 : t + 1, tm = (t == 0) ? g.T - 1 : t - 1;
                        const uint32_t rowo = (t * g.Y + y) * g.Xh;
                        const uint32_t kx = p ? (k0 + 4 == g.Xh ? 0 : k0 + 4) : (k0 == 0 ? g.Xh - 1 : k0 - 1);
                        const int4 v = *reinterpret_cast<const int4*>(own + rowo + k0);
                        const int4 o = *reinterpret_cast<const int4*>(oth + rowo + k0);
                        const int ex = oth[rowo + kx];
                        const int4 na = *reinterpret_cast<const int4*>(oth + (t * g.Y + yp) * g.Xh + k0);
                        const int4 nb = *reinterpret_cast<const int4*>(oth + (t * g.Y + ym) * g.Xh + k0);
                        const int4 nc = *reinterpret_cast<const int4*>(oth + (tp * g.Y + y) * g.Xh + k0);
                        const int4 nd = *reinterpret_cast<const int4*>(oth + (tm * g.Y + y) * g.Xh + k0);
                        int4 ns;
                        ns.x = o.x + (p ? o.y : ex) + na.x + nb.x + nc.x + nd.x;
                        ns.y = o.y + (p ? o.z : o.x) + na.y + nb.y + nc.y + nd.y;
                        ns.z = o.z + (p ? o.w : o.y) + na.z + nb.z + nc.z + nd.z;
                        ns.w = o.w + (p ? ex : o.z) + na.w + nb.w + nc.w + nd.w;
                        int4 wo = make_int4(0, 0, 0, 0);
                        if (WILSON) {
                            const uint32_t x0 = 2 * k0 + p;
                            wo.x = wilson_offset(g.Y, x0, y, t, cd.wil_w, cd.wil_h);
                            wo.y = wilson_offset(g.Y, x0 + 2, y, t, cd.wil_w, cd.wil_h);
                            wo.z = wilson_offset(g.Y, x0 + 4, y, t, cd.wil_w, cd.wil_h);
                            wo.w = wilson_offset(g.Y, x0 + 6, y, t, cd.wil_w, cd.wil_h);
                        }
                        int4 w;
                        w.x = metropolis_update2<true>(v.x, ns.x, wo.x, r.x, s_dn, s_up, kc, accepted);
                        w.y = metropolis_update2<true>(v.y, ns.y, wo.y, r.y, s_dn, s_up, kc, accepted);
                        w.z = metropolis_update2<true>(v.z, ns.z, wo.z, r.z, s_dn, s_up, kc, accepted);
                        w.w = metropolis_update2<true>(v.w, ns.w, wo.w, r.w, s_dn, s_up, kc, accepted);
                        *reinterpret_cast<int4*>(own + rowo + k0) = w;
                    } else if (VEC == 2) {
                        // Xh % 2 == 0 (36^3: Xh = 18): the group is two aligned pairs, each inside one row;
                        // 64-bit shared-memory accesses
                        uint32_t k = gk[j], y = gy[j], t = gt[j];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint32_t p = (y + t + colour) & 1u;
                            const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
                            const uint32_t tp = (t + 1 == g.T) ? 0 : t + 1, tm = (t == 0) ? g.T - 1 : t - 1;
                            const uint32_t rowo = (t * g.Y + y) * g.Xh;
                            const uint32_t kx = p ? (k + 2 == g.Xh ? 0 : k + 2) : (k == 0 ? g.Xh - 1 : k - 1);
                            const int2 v = *reinterpret_cast<const int2*>(own + rowo + k);
                            const int2 o = *reinterpret_cast<const int2*>(oth + rowo + k);
                            const int ex = oth[rowo + kx];
                            const int2 na = *reinterpret_cast<const int2*>(oth + (t * g.Y + yp) * g.Xh + k);
                            const int2 nb = *reinterpret_cast<const int2*>(oth + (t * g.Y + ym) * g.Xh + k);
                            const int2 nc = *reinterpret_cast<const int2*>(oth + (tp * g.Y + y) * g.Xh + k);
                            const int2 nd = *reinterpret_cast<const int2*>(oth + (tm * g.Y + y) * g.Xh + k);
                            const int ns0 = o.x + (p ? o.y : ex) + na.x + nb.x + nc.x + nd.x;
                            const int ns1 = o.y + (p ? ex : o.x) + na.y + nb.y + nc.y + nd.y;
                            int w0 = 0, w1 = 0;
                            if (WILSON) {
                                w0 = wilson_offset(g.Y, 2 * k + p, y, t, cd.wil_w, cd.wil_h);
                                w1 = wilson_offset(g.Y, 2 * k + 2 + p, y, t, cd.wil_w, cd.wil_h);
                            }
                            int2 w;
                            w.x = metropolis_update2<true>(v.x, ns0, w0, word_of(r, (uint32_t)(2 * e)), s_dn, s_up, kc, accepted);
                            w.y = metropolis_update2<true>(v.y, ns1, w1, word_of(r, (uint32_t)(2 * e + 1)), s_dn, s_up, kc, accepted);
                            *reinterpret_cast<int2*>(own + rowo + k) = w;
                            k += 2;
                            if (k == g.Xh) {
                                k = 0;
                                if (++y == g.Y) { y = 0; ++t; }
                            }
                        }
                    } else {
                        uint32_t k = gk[j], y = gy[j], t = gt[j];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t p = (y + t + colour) & 1u;
                            const uint32_t rowo = (t * g.Y + y) * g.Xh;
                            const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
                            const uint32_t tp = (t + 1 == g.T) ? 0 : t + 1, tm = (t == 0) ? g.T - 1 : t - 1;
                            const uint32_t kx = p ? (k + 1 == g.Xh ? 0 : k + 1) : (k == 0 ? g.Xh - 1 : k - 1);
                            const int v = own[rowo + k];
                            int nsum = oth[rowo + k] + oth[rowo + kx];
                            nsum += oth[(t * g.Y + yp) * g.Xh + k] + oth[(t * g.Y + ym) * g.Xh + k];
                            nsum += oth[(tp * g.Y + y) * g.Xh + k] + oth[(tm * g.Y + y) * g.Xh + k];
                            int woff = 0;
                            if (WILSON) woff = wilson_offset(g.Y, 2 * k + p, y, t, cd.wil_w, cd.wil_h);
                            own[rowo + k] = metropolis_update2<true>(v, nsum, woff, word_of(r, (uint32_t)e), s_dn, s_up, kc, accepted);
                            if (++k == g.Xh) {  // next position of the colour array: next row, maybe next plane
                                k = 0;
                                if (++y == g.Y) { y = 0; ++t; }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (a.energy_every && step % a.energy_every == 0) {
            // Action::integer_energy (fields.rs:728-737): sum over sites of dx^2 + dy^2 - dt^2, exact
            long long e = 0;
            for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
                const uint32_t row = i / g.Xh, k = i - row * g.Xh;
                const uint32_t t = row / g.Y, y = row - t * g.Y;
                const uint32_t p0 = (y + t) & 1u;
                const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, tp = (t + 1 == g.T) ? 0 : t + 1;
                const uint32_t kp = (k + 1 == g.Xh) ? 0 : k + 1;
                const long long va = s_c[0][i], vb = s_c[1][i];
                const long long ax = p0 ? s_c[1][row * g.Xh + kp] : vb, bx = p0 ? va : s_c[0][row * g.Xh + kp];
                const uint32_t iy = (t * g.Y + yp) * g.Xh + k, it = (tp * g.Y + y) * g.Xh + k;
                const long long ay = s_c[1][iy], by = s_c[0][iy], at = s_c[1][it], bt = s_c[0][it];
                e += (va - ax) * (va - ax) + (vb - bx) * (vb - bx) + (va - ay) * (va - ay) + (vb - by) * (vb - by) -
                     (va - at) * (va - at) - (vb - bt) * (vb - bt);
            }
            const long long tot = block_sum_ll(e, s_red);
            if (threadIdx.x == 0 && n_meas < a.elog_cap) elog[(size_t)chain * a.elog_cap + n_meas] = tot;
            ++n_meas;
        }
        if (a.normalize_every && step % a.normalize_every == 0) {
            // normalize_random: offset := stored height of the picked site (kernels_v1.cuh, k_pick_shift)
            if (threadIdx.x == 0) {
                const uint64_t site = pick_site(cd.k0, cd.k1, 2u, sweep, 0u, n_sites);
                const uint32_t hi = (uint32_t)(site >> 1);
                const uint32_t row = hi / g.Xh, k = hi - row * g.Xh;
                const uint32_t t = row / g.Y, y = row - t * g.Y;
                const uint32_t x = 2 * k + (uint32_t)(site & 1u);
                s_pick = s_c[(x + y + t) & 1u][hi];
            }
            __syncthreads();
            off_true = s_pick;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
        gc[0][i] = s_c[0][i];
        gc[1][i] = s_c[1][i];
    }
    if (threadIdx.x == 0) offset[chain] = off_true;
    if (a.count) {
        const long long tot = block_sum_ll((long long)accepted, s_red);
        if (threadIdx.x == 0) atomicAdd(acc + (size_t)chain * kAccSlots, (unsigned long long)tot);
    }
}

inline size_t resident_smem_bytes(const Geom& g, uint32_t max_thr) {
    const uint32_t kc = max_thr > 4 ? max_thr - 1 : 3;
    return (size_t)g.plane * g.T * 2 * sizeof(int32_t) + sizeof(uint2) * 2 * (2 * kc + 1);
}

// One CTA per chain can hold the lattice when both colours plus the decision table fit in shared memory and
// a thread has at most kResidentMaxGroups counter groups per colour.
inline bool resident_fits(const Geom& g, uint32_t max_thr, size_t smem_limit) {
    const uint64_t half = (uint64_t)g.plane * g.T;
    if (g.ghost || half % 4 != 0) return false;
    if (half / 4 > (uint64_t)kResidentThreads * kResidentMaxGroups) return false;
    return resident_smem_bytes(g, max_thr) <= smem_limit;
}

}  // namespace lqft

// kernels_f64.cuh — real-valued height fields (Field<f64,...>, admitted by HeightVariable, fields.rs:573-599).
//
// No driver of the reference uses them, so this is the general scalar path only (one thread per site, any
// even extents, one GPU).  Arithmetic follows the reference literally in f64: both local actions are summed
// in the neighbour order [+x,+y,+t,-x,-y,-t] with separate multiplies and adds (no FMA contraction), and the
// accept test is draw <= exp((old - new) * temp) (algorithm.rs:146-153) with CUDA's exp.  Decisions can only
// differ from a CPU evaluation when draw lies within an ulp of the probability.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_v1.cuh"

namespace lqft {

__device__ __forceinline__ double sq_rn(double d) { return __dmul_rn(d, d); }

template <bool WILSON>
__global__ void __launch_bounds__(256)
k_sweep_generic_f64(Geom g, int colour, uint64_t sweep, const uint64_t* __restrict__ ctr, double* __restrict__ own_base,
                    const double* __restrict__ oth_base, const ChainDev* __restrict__ chains,
                    const double* __restrict__ betas, unsigned long long* __restrict__ acc) {
    __shared__ unsigned int s_cnt;
    const uint32_t chain = blockIdx.z;
    const ChainDev cd = chains[chain];
    const double beta = betas[chain];
    if (ctr) sweep += *ctr;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int accepted = 0;
    if (i < g.plane) {
        const uint32_t y = i / g.Xh, k = i - y * g.Xh;
        const uint32_t tg = g.t0 + tl;
        const uint32_t p = (y + tg + (uint32_t)colour) & 1u;
        double* own = own_base + (size_t)chain * g.chain_stride;
        const double* oth = oth_base + (size_t)chain * g.chain_stride;
        const size_t ps = (size_t)plane_self(g, tl) * g.plane;
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
        // site x = 2k + p: +x neighbour at position k + p (mod Xh), -x neighbour at k + p - 1 (mod Xh)
        const uint32_t kxp = p ? (k + 1 == g.Xh ? 0 : k + 1) : k;
        const uint32_t kxm = p ? k : (k == 0 ? g.Xh - 1 : k - 1);
        const double n[6] = {oth[row + kxp], oth[ps + (size_t)yp * g.Xh + k],
                             oth[(size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k],
                             oth[row + kxm], oth[ps + (size_t)ym * g.Xh + k],
                             oth[(size_t)plane_below(g, tl) * g.plane + (size_t)y * g.Xh + k]};
        double mod[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (WILSON) {
            const uint32_t x = 2 * k + p;
            if (x < cd.wil_w && tg < cd.wil_h) {  // fields.rs:821-837
                if (y == 0u) mod[1] = 1.0;
                if (y == 1u % g.Y) mod[4] = -1.0;
            }
        }
        const uint64_t hi = ((uint64_t)tg * g.Y + y) * g.Xh + k;
        const uint64_t grp = hi >> 2;
        const u32x4 r = philox4x32((uint32_t)grp, (uint32_t)colour | ((uint32_t)(grp >> 32) << 8), (uint32_t)sweep,
                                   (uint32_t)(sweep >> 32), cd.k0, cd.k1);
        const uint32_t word = word_of(r, (uint32_t)(hi & 3u));
        const double old_value = own[row + k];
        const double new_value = (word >> 31) ? __dadd_rn(old_value, 1.0) : __dadd_rn(old_value, -1.0);
        double old_action = 0.0, new_action = 0.0;
#pragma unroll
        for (int d = 0; d < 6; ++d) {  // fields.rs:743-750 / 809-815
            old_action = __dadd_rn(old_action, sq_rn(__dadd_rn(__dadd_rn(old_value, -n[d]), mod[d])));
            new_action = __dadd_rn(new_action, sq_rn(__dadd_rn(__dadd_rn(new_value, -n[d]), mod[d])));
        }
        const double draw = (double)(word & 0x7fffffffu) * (1.0 / 2147483648.0);
        const double prob = exp(__dmul_rn(__dadd_rn(old_action, -new_action), beta));
        if (draw <= prob) {
            own[row + k] = new_value;
            accepted = 1;
        }
    }
    block_count_flush(&s_cnt, accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

__global__ void k_split_f64(Geom g, const double* __restrict__ lex, double* __restrict__ c0, double* __restrict__ c1) {
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const uint32_t y = i / g.Xh;
    const uint32_t q = (y + g.t0 + tl) & 1u;
    const double2 v = reinterpret_cast<const double2*>(lex)[(size_t)tl * g.plane + i];
    const size_t dst = (size_t)plane_self(g, tl) * g.plane + i;
    (q ? c1 : c0)[dst] = v.x;
    (q ? c0 : c1)[dst] = v.y;
}

__global__ void k_merge_f64(Geom g, double* __restrict__ lex, const double* __restrict__ c0,
                            const double* __restrict__ c1, const double* __restrict__ offset) {
    const uint32_t tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const uint32_t y = i / g.Xh;
    const uint32_t q = (y + g.t0 + tl) & 1u;
    const size_t src = (size_t)plane_self(g, tl) * g.plane + i;
    const double off = *offset;
    double2 v;
    v.x = (q ? c1 : c0)[src] - off;
    v.y = (q ? c0 : c1)[src] - off;
    reinterpret_cast<double2*>(lex)[(size_t)tl * g.plane + i] = v;
}

__global__ void k_init_hot_f64(Geom g, double* __restrict__ c0_base, double* __restrict__ c1_base,
                               const ChainDev* __restrict__ chains) {
    const uint32_t chain = blockIdx.z, tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    const ChainDev cd = chains[chain];
    const uint32_t y = i / g.Xh;
    const uint32_t tg = g.t0 + tl;
    const uint32_t q = (y + tg) & 1u;
    const uint64_t idx = 2 * ((uint64_t)tg * g.plane + i);
    const size_t dst = (size_t)chain * g.chain_stride + (size_t)plane_self(g, tl) * g.plane + i;
    (q ? c1_base : c0_base)[dst] = (double)hot_value(cd.k0, cd.k1, idx);
    (q ? c0_base : c1_base)[dst] = (double)hot_value(cd.k0, cd.k1, idx + 1);
}

// Stored height of a global site of one chain.
__device__ __forceinline__ double site_value_f64(const Geom& g, const double* c0, const double* c1, uint64_t site) {
    const uint64_t hi = site >> 1;
    const uint32_t tg = (uint32_t)(hi / g.plane);
    const uint32_t i = (uint32_t)(hi - (uint64_t)tg * g.plane);
    const uint32_t y = i / g.Xh;
    const uint32_t x = 2 * (i - y * g.Xh) + (uint32_t)(site & 1u);
    const uint32_t col = (x + y + tg) & 1u;
    return (col ? c1 : c0)[(size_t)plane_self(g, tg - g.t0) * g.plane + i];
}

// normalize_random for f64 chains: offset := stored height of the picked site (see kernels_v1.cuh, k_pick_shift)
__global__ void k_pick_offset_f64(Geom g, const double* __restrict__ c0_base, const double* __restrict__ c1_base,
                                  const ChainDev* __restrict__ chains, uint32_t n_chains, uint64_t counter,
                                  const uint64_t* __restrict__ ctr, int64_t explicit_site, int32_t only_chain,
                                  double* __restrict__ offset) {
    const uint32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    if (only_chain >= 0 && (int32_t)chain != only_chain) return;
    const ChainDev cd = chains[chain];
    if (ctr) counter += *ctr;
    const uint64_t n = (uint64_t)g.X * g.Y * g.T;
    const uint64_t site = explicit_site >= 0 ? (uint64_t)explicit_site : pick_site(cd.k0, cd.k1, 2u, counter, 0u, n);
    offset[chain] = site_value_f64(g, c0_base + (size_t)chain * g.chain_stride, c1_base + (size_t)chain * g.chain_stride, site);
}
__global__ void k_add_offset_f64(double* __restrict__ offset, uint32_t chain, double shift) { offset[chain] += shift; }

__global__ void k_gather_sites_f64(Geom g, const double* __restrict__ c0, const double* __restrict__ c1,
                                   const uint64_t* __restrict__ sites, uint32_t n, double* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = site_value_f64(g, c0, c1, sites[j]);
}

// ---- observables: compensated fp64 reductions (the role of kahan.rs) ---------------------------------------
// Two-sum (Knuth): s = a + b exactly as s + err.
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& err) {
    s = __dadd_rn(a, b);
    const double bb = __dadd_rn(s, -a);
    err = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
}
struct Comp {  // running compensated sum
    double hi, lo;
};
__device__ __forceinline__ Comp comp_add(Comp a, double x) {
    double s, e;
    two_sum(a.hi, x, s, e);
    return Comp{s, __dadd_rn(a.lo, e)};
}
__device__ __forceinline__ Comp comp_merge(Comp a, Comp b) {
    double s, e;
    two_sum(a.hi, b.hi, s, e);
    return Comp{s, __dadd_rn(__dadd_rn(a.lo, b.lo), e)};
}
__device__ __forceinline__ Comp warp_comp_sum(Comp v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Comp other{__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o)};
        v = comp_merge(v, other);
    }
    return v;
}

// Pass 1: per-block compensated partial sums of energy, action and (optionally) slice moments.
// grid = (bx, Tl, n_chains); part[((chain * Tl + tl) * bx + blockIdx.x) * 4 + q] as (hi, lo) pairs.
__global__ void __launch_bounds__(256)
k_observe_f64(Geom g, const double* __restrict__ c0_base, const double* __restrict__ c1_base, double2* __restrict__ part) {
    __shared__ Comp s_part[4][8];
    const uint32_t chain = blockIdx.z, tl = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Comp q[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    if (i < g.plane) {
        const double* c0 = c0_base + (size_t)chain * g.chain_stride;
        const double* c1 = c1_base + (size_t)chain * g.chain_stride;
        const uint32_t y = i / g.Xh, k = i - y * g.Xh;
        const uint32_t p0 = (y + g.t0 + tl) & 1u;
        const size_t ps = (size_t)plane_self(g, tl) * g.plane;
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1;
        const uint32_t kp = (k + 1 == g.Xh) ? 0 : k + 1;
        const size_t up = (size_t)plane_above(g, tl) * g.plane + (size_t)y * g.Xh + k;
        const size_t yr = ps + (size_t)yp * g.Xh + k;
        const double a = c0[row + k], b = c1[row + k];
        const double ax = p0 ? c1[row + kp] : b, bx = p0 ? a : c0[row + kp];
        const double dxa = sq_rn(a - ax), dxb = sq_rn(b - bx);
        const double dya = sq_rn(a - c1[yr]), dyb = sq_rn(b - c0[yr]);
        const double dta = sq_rn(a - c1[up]), dtb = sq_rn(b - c0[up]);
        const double terms_e[6] = {dxa, dxb, dya, dyb, -dta, -dtb};
        const double terms_s[6] = {dxa, dxb, dya, dyb, dta, dtb};
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            q[0] = comp_add(q[0], terms_e[j]);
            q[1] = comp_add(q[1], terms_s[j]);
        }
        q[2] = comp_add(comp_add(q[2], a), b);
        q[3] = comp_add(comp_add(q[3], sq_rn(a)), sq_rn(b));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const Comp w = warp_comp_sum(q[j]);
        if (lane == 0) s_part[j][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        Comp tot{0, 0};
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) tot = comp_merge(tot, s_part[threadIdx.x][w]);
        part[(((size_t)chain * g.Tl + tl) * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = make_double2(tot.hi, tot.lo);
    }
}

// Pass 2 (deterministic order): one block per chain folds the partials.  out[chain*2 + {0,1}] = energy, action;
// mom[(chain*T + t)*2 + {0,1}] = S1, S2 (when mom != nullptr).
__global__ void k_observe_fold_f64(Geom g, const double2* __restrict__ part, uint32_t bx, double* __restrict__ out,
                                   double* __restrict__ mom) {
    const uint32_t chain = blockIdx.x;
    if (threadIdx.x < 2) {
        Comp tot{0, 0};
        for (uint32_t tl = 0; tl < g.Tl; ++tl)
            for (uint32_t b = 0; b < bx; ++b) {
                const double2 p = part[(((size_t)chain * g.Tl + tl) * bx + b) * 4 + threadIdx.x];
                tot = comp_merge(tot, Comp{p.x, p.y});
            }
        out[chain * 2 + threadIdx.x] = __dadd_rn(tot.hi, tot.lo);
    }
    if (mom) {
        for (uint32_t tl = threadIdx.x; tl < g.Tl; tl += blockDim.x)
            for (int j = 0; j < 2; ++j) {
                Comp tot{0, 0};
                for (uint32_t b = 0; b < bx; ++b) {
                    const double2 p = part[(((size_t)chain * g.Tl + tl) * bx + b) * 4 + 2 + j];
                    tot = comp_merge(tot, Comp{p.x, p.y});
                }
                mom[((size_t)chain * g.T + g.t0 + tl) * 2 + j] = __dadd_rn(tot.hi, tot.lo);
            }
    }
}

// Append the energy of every chain to its log (lqft_run).
__global__ void k_energy_log_f64(const double* __restrict__ obs, double* __restrict__ log, uint32_t cap, uint32_t n_chains,
                                 uint32_t* __restrict__ n_meas) {
    const uint32_t m = *n_meas;
    for (uint32_t c = threadIdx.x; c < n_chains; c += blockDim.x)
        if (m < cap) log[(size_t)c * cap + m] = obs[(size_t)c * 2];
    __syncthreads();
    if (threadIdx.x == 0) *n_meas = m + 1;
}

// DifferencePlot for f64 fields (outputdata.rs:241-257, kahan.rs:21-31).
__global__ void k_difference_update_f64(Geom g, const double* __restrict__ c0, const double* __restrict__ c1,
                                        double* __restrict__ sum, double* __restrict__ comp) {
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
    const double a = c0[row + k], b = c1[row + k];
    const double ax = p0 ? c1[row + kp] : b, bx = p0 ? a : c0[row + kp];
    const size_t n_slab = (size_t)g.Tl * g.plane * 2;
    const size_t lex_even = ((size_t)tl * g.plane + i) * 2, lex_odd = lex_even + 1;
    const size_t la = p0 ? lex_odd : lex_even, lb = p0 ? lex_even : lex_odd;
    const double d[3][2] = {{a - ax, b - bx}, {a - c1[yr], b - c0[yr]}, {a - c1[up], b - c0[up]}};
#pragma unroll
    for (int dir = 0; dir < 3; ++dir)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const size_t o = (size_t)dir * n_slab + (s ? lb : la);
            const double yv = d[dir][s] - comp[o];
            const double tv = sum[o] + yv;
            comp[o] = (tv - sum[o]) - yv;
            sum[o] = tv;
        }
}

}  // namespace lqft

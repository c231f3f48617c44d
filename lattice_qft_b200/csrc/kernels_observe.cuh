// kernels_observe.cuh — observables of the measurement schedule, computed and logged on the device.
//
// One pass over the field gives, per chain, the energy (fields.rs:728-737), the action (fields.rs:719-726) and —
// when the correlator is due — the per-slice moments S1[t] = sum h, S2[t] = sum h^2.  The correlator of
// CorrelationPlotOutputData::update (outputdata.rs:313-344),
//     C[tau] = sum_rep sum_{y : t_y - t_x = tau} (h_x - h_y)^2 = sum_rep ( A h_x^2 - 2 h_x S1[t_x + tau] + S2[t_x + tau] ),  A = X * Y,
// is then an O(reps * T) epilogue in exact int64 arithmetic.  All of these depend on height differences only, so they
// are evaluated on the STORED heights, whatever the pending normalisation offset or the bias of the 16-bit copy.
//
// Logging (Simulation::run pushes one value per due step, computation.rs:262-265): the last block to arrive at the
// launch's ticket appends the sums to the device-side logs and clears the accumulators, so a measurement is ONE
// launch (two with the correlator on many chains) and no host round trip.  The slot of a measurement follows from
// the step number alone (the number of due steps before it), so nothing is counted on the device.
#pragma once
#include <type_traits>

#include "kernels_v1.cuh"

namespace lqft {

constexpr uint32_t kObsEnergy = 1u, kObsAction = 2u, kObsCorr = 4u;
constexpr int kCorrTile = 256;
constexpr int kCorrSmemT = 1024;

struct ObsFin {
    unsigned int* ticket;   // blocks of the running launch that have flushed their sums (0 between launches)
    long long* elog;        // [n_chains][cap_e]
    long long* alog;        // [n_chains][cap_a]
    double* clog;           // [n_chains][cap_c][T]
    long long* result;      // [n_chains][2] = E, S of this launch (standalone lqft_energy / lqft_action), nullable
    uint32_t cap_e, cap_a, cap_c;
    uint32_t every_e, every_a, every_c, reps;
    const uint64_t* run;    // device: {sweep_base, first_step} of the running schedule (null: 0, 0);
                            // step of this launch = (sweep + *ctr) - sweep_base
    const uint64_t* sites;  // explicit reference sites of the correlator (standalone), nullable
    uint32_t fused;         // 1: the last block finalises inside the observe launch
};

__device__ __forceinline__ uint64_t due_before(uint64_t step, uint64_t first, uint32_t every) {
    // multiples of `every` in [first, step)
    return (step + every - 1) / every - (first + every - 1) / every;
}

// E and S of chains [chain0, chain0 + n) into the logs / the result array; accumulators cleared.  Whole block.
__device__ __forceinline__ void obs_log_sums(const ObsFin& f, uint32_t due, uint64_t sweep, uint32_t chain0, uint32_t n,
                                             unsigned long long* __restrict__ obs) {
    const uint64_t step = sweep - (f.run ? f.run[0] : 0ull), first = f.run ? f.run[1] : 0ull;
    const uint64_t me = (due & kObsEnergy) ? due_before(step, first, f.every_e) : 0;
    const uint64_t ma = (due & kObsAction) ? due_before(step, first, f.every_a) : 0;
    for (uint32_t c = chain0 + threadIdx.x; c < chain0 + n; c += blockDim.x) {
        const long long e = (long long)__ldcg(obs + (size_t)c * 2), s = (long long)__ldcg(obs + (size_t)c * 2 + 1);
        if ((due & kObsEnergy) && me < f.cap_e) f.elog[(size_t)c * f.cap_e + me] = e;
        if ((due & kObsAction) && ma < f.cap_a) f.alog[(size_t)c * f.cap_a + ma] = s;
        if (f.result) {
            f.result[(size_t)c * 2] = e;
            f.result[(size_t)c * 2 + 1] = s;
        }
        obs[(size_t)c * 2] = 0;
        obs[(size_t)c * 2 + 1] = 0;
    }
}

template <typename H>
__device__ __forceinline__ long long stored_height(const Geom& g, const H* __restrict__ c0, const H* __restrict__ c1, uint64_t site,
                                                   uint32_t& tg_out) {
    const uint64_t hi = site >> 1;
    const uint32_t tg = (uint32_t)(hi / g.plane);
    const uint32_t i = (uint32_t)(hi - (uint64_t)tg * g.plane);
    tg_out = tg;
    const uint32_t tl = (tg + g.T - g.t0) % g.T;
    if (tl >= g.Tl) return 0;  // another rank owns the site
    const uint32_t y = i / g.Xh;
    const uint32_t x = 2 * (i - y * g.Xh) + (uint32_t)(site & 1u);
    const uint32_t col = (x + y + tg) & 1u;
    return (long long)(col ? c1 : c0)[(size_t)plane_self(g, tl) * g.plane + i];
}

// Correlator epilogue of one chain, whole block: out[tau] = sum_rep (A v^2 - 2 v S1[t+tau] + S2[t+tau]); the
// chain's moments are cleared afterwards.  Reference sites: explicit list, or stream LQFT_STREAM_CORRELATOR at
// counter `counter` (lqft_host_pick_site).  `gathered` (slabs): heights collected over the ranks beforehand.
template <typename H>
__device__ void corr_fold_chain(const Geom& g, const H* __restrict__ c0, const H* __restrict__ c1, const ChainDev cd,
                                const uint64_t* __restrict__ sites, const long long* __restrict__ gathered, uint32_t reps,
                                uint64_t counter, unsigned long long* __restrict__ mom, double* __restrict__ out) {
    __shared__ long long s_v[kCorrTile];
    __shared__ uint32_t s_t[kCorrTile];
    __shared__ long long s_mom[2 * kCorrSmemT];  // the chain's moments, when they fit: the epilogue reads each 2 * reps times
    const uint64_t n_sites = (uint64_t)g.X * g.Y * g.T;
    const long long A = (long long)g.X * g.Y;
    const bool staged = g.T <= (uint32_t)kCorrSmemT;
    if (staged)
        for (uint32_t t = threadIdx.x; t < 2 * g.T; t += blockDim.x) s_mom[t] = (long long)__ldcg(mom + t);
    for (uint32_t tau0 = 0; tau0 < g.T; tau0 += blockDim.x) {
        const uint32_t tau = tau0 + threadIdx.x;
        long long acc = 0;
        for (uint32_t r0 = 0; r0 < reps; r0 += kCorrTile) {
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < kCorrTile && r0 + j < reps; j += blockDim.x) {
                const uint32_t r = r0 + j;
                const uint64_t site = sites ? sites[r] : pick_site(cd.k0, cd.k1, 4u, counter, r, n_sites);
                uint32_t tg;
                const long long v = stored_height(g, c0, c1, site, tg);
                s_v[j] = gathered ? gathered[r] : v;
                s_t[j] = tg;
            }
            __syncthreads();
            if (tau < g.T) {
                const uint32_t nr = min((uint32_t)kCorrTile, reps - r0);
                for (uint32_t j = 0; j < nr; ++j) {
                    uint32_t t = s_t[j] + tau;
                    if (t >= g.T) t -= g.T;
                    const long long v = s_v[j];
                    const long long m1 = staged ? s_mom[2 * t] : (long long)__ldcg(mom + (size_t)t * 2);
                    const long long m2 = staged ? s_mom[2 * t + 1] : (long long)__ldcg(mom + (size_t)t * 2 + 1);
                    acc += A * v * v - 2 * v * m1 + m2;
                }
            }
        }
        if (tau < g.T && out) out[tau] = (double)acc;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < g.T; t += blockDim.x) {
        mom[(size_t)t * 2] = 0;
        mom[(size_t)t * 2 + 1] = 0;
    }
}

__device__ __forceinline__ double* corr_slot(const ObsFin& f, const Geom& g, uint32_t chain, uint64_t sweep) {
    if (f.every_c == 0) return f.clog;  // standalone: one row
    const uint64_t step = sweep - (f.run ? f.run[0] : 0ull), first = f.run ? f.run[1] : 0ull;
    const uint64_t m = due_before(step, first, f.every_c);
    return m < f.cap_c ? f.clog + ((size_t)chain * f.cap_c + m) * g.T : nullptr;
}

// Energy, action and (MOM) slice moments of chains [chain0, chain0 + gridDim.z): one block per (y chunk, plane,
// chain), a quad of 4 + 4 sites per thread and iteration.  X % 8 == 0.  grid = (n_ychunks, Tl, n_chains).
template <typename H, bool MOM>
__global__ void __launch_bounds__(256)
k_observe_rows(const Geom g, const H* __restrict__ c0_base, const H* __restrict__ c1_base, const uint32_t chain0,
               unsigned long long* __restrict__ obs, unsigned long long* __restrict__ mom, const ObsFin fin, const uint32_t due,
               uint64_t sweep, const uint64_t* __restrict__ ctr, const ChainDev* __restrict__ chains) {
    __shared__ bool s_last;
    const uint32_t chain = chain0 + blockIdx.z, tl = blockIdx.y;
    const uint32_t tg = (g.t0 + tl) % g.T;
    const uint32_t Qx = g.Xh >> 2;
    const uint32_t y0 = blockIdx.x * g.Y / gridDim.x, y1 = (blockIdx.x + 1u) * g.Y / gridDim.x;
    const H* c0 = c0_base + (size_t)chain * g.chain_stride;
    const H* c1 = c1_base + (size_t)chain * g.chain_stride;
    const size_t ps = (size_t)plane_self(g, tl) * g.plane, pa = (size_t)plane_above(g, tl) * g.plane;
    long long q[4] = {0, 0, 0, 0};
    // differences in i32 like the reference (fields.rs:728-737: `value - neighbour` in T).  i32 storage: squares and
    // sums in i64 (one IMAD.WIDE with a 64-bit addend per term).  16-bit storage: heights lie in [0, 4096), a square
    // is below 2^24 and the 24 terms of one iteration stay below 2^31, so they are summed in 32 bits (one IMAD per
    // term) and only the iteration's totals are widened.
    constexpr bool NARROW = sizeof(H) == 2;
    using W = typename std::conditional<NARROW, int, long long>::type;
    auto sq = [](int u, int v) { const int d = u - v; return (W)d * (W)d; };
    for (uint32_t i = threadIdx.x; i < (y1 - y0) * Qx; i += blockDim.x) {
        const uint32_t yy = i / Qx, y = y0 + yy, k0 = (i - yy * Qx) << 2;
        const uint32_t p0 = (y + tg) & 1u;  // x parity of the colour-0 sites of this row
        const size_t row = ps + (size_t)y * g.Xh;
        const uint32_t yp = (y + 1 == g.Y) ? 0 : y + 1;
        const uint32_t kn = (k0 + 4 == g.Xh) ? 0 : k0 + 4;
        const size_t up = pa + (size_t)y * g.Xh + k0, yr = ps + (size_t)yp * g.Xh + k0;
        const int4 a = load_h4(c0 + row + k0), b = load_h4(c1 + row + k0);
        const int4 ay = load_h4(c1 + yr), by = load_h4(c0 + yr);
        const int4 at = load_h4(c1 + up), bt = load_h4(c0 + up);
        const int e = (int)(p0 ? c1 : c0)[row + kn];  // +x neighbour across the quad edge
        const int4 ax = p0 ? make_int4(b.y, b.z, b.w, e) : b;
        const int4 bx = p0 ? a : make_int4(a.y, a.z, a.w, e);
        const W dx = sq(a.x, ax.x) + sq(a.y, ax.y) + sq(a.z, ax.z) + sq(a.w, ax.w) +
                     sq(b.x, bx.x) + sq(b.y, bx.y) + sq(b.z, bx.z) + sq(b.w, bx.w);
        const W dy = sq(a.x, ay.x) + sq(a.y, ay.y) + sq(a.z, ay.z) + sq(a.w, ay.w) +
                     sq(b.x, by.x) + sq(b.y, by.y) + sq(b.z, by.z) + sq(b.w, by.w);
        const W dt = sq(a.x, at.x) + sq(a.y, at.y) + sq(a.z, at.z) + sq(a.w, at.w) +
                     sq(b.x, bt.x) + sq(b.y, bt.y) + sq(b.z, bt.z) + sq(b.w, bt.w);
        q[0] += (long long)(dx + dy - dt);
        q[1] += (long long)(dx + dy + dt);
        if (MOM) {
            q[2] += (long long)((W)a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w);
            q[3] += (long long)(sq(a.x, 0) + sq(a.y, 0) + sq(a.z, 0) + sq(a.w, 0) + sq(b.x, 0) + sq(b.y, 0) + sq(b.z, 0) + sq(b.w, 0));
        }
    }
    unsigned long long* const dst[4] = {
        obs + (size_t)chain * 2, obs + (size_t)chain * 2 + 1,
        MOM ? mom + ((size_t)chain * g.T + tg) * 2 : nullptr,
        MOM ? mom + ((size_t)chain * g.T + tg) * 2 + 1 : nullptr};
    block_sum_flush<4>(q, dst);
    if (!fin.fused) return;
    // last block of the launch: log, fold, clear
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        s_last = atomicAdd(fin.ticket, 1u) + 1u == total;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (ctr) sweep += *ctr;
    obs_log_sums(fin, due, sweep, chain0, gridDim.z, obs);
    if (MOM && (due & kObsCorr)) {
        for (uint32_t c = chain0; c < chain0 + gridDim.z; ++c)
            corr_fold_chain<H>(g, c0_base + (size_t)c * g.chain_stride, c1_base + (size_t)c * g.chain_stride, chains[c], fin.sites,
                               nullptr, fin.reps, sweep, mom + (size_t)c * g.T * 2, corr_slot(fin, g, c, sweep));
    }
    if (threadIdx.x == 0) *fin.ticket = 0;
}

// Separate finalisers: after the scalar observe kernel, or after the all-reduce of slab partial sums.
__global__ void __launch_bounds__(256)
k_obs_finalize(const ObsFin fin, const uint32_t due, uint64_t sweep, const uint64_t* __restrict__ ctr, uint32_t chain0, uint32_t n,
               unsigned long long* __restrict__ obs) {
    if (ctr) sweep += *ctr;
    obs_log_sums(fin, due, sweep, chain0, n, obs);
}

// grid = n chains (from chain0), block = 128.
template <typename H>
__global__ void __launch_bounds__(128)
k_corr_fold(const Geom g, const H* __restrict__ c0_base, const H* __restrict__ c1_base, const ObsFin fin, uint64_t sweep,
            const uint64_t* __restrict__ ctr, uint32_t chain0, const ChainDev* __restrict__ chains,
            const long long* __restrict__ gathered, unsigned long long* __restrict__ mom) {
    if (ctr) sweep += *ctr;
    const uint32_t c = chain0 + blockIdx.x;
    corr_fold_chain<H>(g, c0_base + (size_t)c * g.chain_stride, c1_base + (size_t)c * g.chain_stride, chains[c], fin.sites,
                       gathered ? gathered + (size_t)c * fin.reps : nullptr, fin.reps, sweep, mom + (size_t)c * g.T * 2,
                       corr_slot(fin, g, c, sweep));
}

// Slabs: stored heights of the reference sites this rank owns (0 elsewhere), to be summed over the ranks.
// grid = (ceil(reps / 128), n chains from chain0).
template <typename H>
__global__ void __launch_bounds__(128)
k_corr_gather(const Geom g, const H* __restrict__ c0_base, const H* __restrict__ c1_base, const ObsFin fin, uint64_t sweep,
              const uint64_t* __restrict__ ctr, uint32_t chain0, const ChainDev* __restrict__ chains, long long* __restrict__ gathered) {
    if (ctr) sweep += *ctr;
    const uint32_t c = chain0 + blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= fin.reps) return;
    const ChainDev cd = chains[c];
    const uint64_t site = fin.sites ? fin.sites[r] : pick_site(cd.k0, cd.k1, 4u, sweep, r, (uint64_t)g.X * g.Y * g.T);
    uint32_t tg;
    gathered[(size_t)c * fin.reps + r] =
        stored_height(g, c0_base + (size_t)c * g.chain_stride, c1_base + (size_t)c * g.chain_stride, site, tg);
}

// DifferencePlot on either storage type (outputdata.rs:241-276, kahan.rs:100-107): per-bond running sum of
// bond_difference in the 3 positive directions, Kahan-compensated like WelfordsAlgorithm64.sum.
// sum/comp are [3][Tl*plane*2] in lexicographic slab order.  grid = (ceil(plane/bd), Tl, 1)
template <typename H>
__global__ void k_difference_update_h(Geom g, const H* __restrict__ c0, const H* __restrict__ c1, double* __restrict__ sum,
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
    const int a = (int)c0[row + k], b = (int)c1[row + k];
    const int ax = p0 ? (int)c1[row + kp] : b, bx = p0 ? a : (int)c0[row + kp];
    const int ay = (int)c1[yr], by = (int)c0[yr], at = (int)c1[up], bt = (int)c0[up];
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

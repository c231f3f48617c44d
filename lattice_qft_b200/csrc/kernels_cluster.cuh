// kernels_cluster.cuh — chains of up to ~1.5 MB (64^3, 128 x 64 x 64 ...): a whole run inside one launch, the field
// resident in the shared memory of a thread-block cluster.
//
// The launch path (kernels_narrow.cuh) needs one kernel per half-sweep; at 64^3 a half-sweep is 131 072 updates per
// chain, a few microseconds of work, and launch latency, the ramp and the first loads from L2 are most of its time
// (BASELINE configs 2 and 4: coupling and Wilson scans of 64^3 chains; computation.rs:229-270, 313-340).  Here a
// cluster of 1 ... 8 CTAs owns one chain for the whole call: CTA r keeps the t-planes [r TP, (r+1) TP) of both
// colours of the 16-bit field in its shared memory, between two halo planes; a half-sweep updates the owned
// planes from shared memory (six 16-byte loads, one store and the decision table, all on chip), stores the first
// and last owned plane a second time into the neighbours' halo planes through distributed shared memory
// (st.shared::cluster), and ends with the hardware cluster barrier.  Energy and action of a measured step are
// accumulated by the colour-1 half-sweep of that step from the values it holds in registers anyway (below);
// normalize is one thread.  The field returns to the 16-bit arrays in global memory when the call ends.
// Same sites, same Philox counters, same packed arithmetic as k_sweep_n16 (n16_row_decide): bit-identical results.
//
// Energy / action without a pass of their own: n16_obs_pair (kernels_narrow.cuh).
#pragma once
#include "kernels_narrow.cuh"

namespace lqft {

constexpr int kClusterMaxThreads = 768;              // 24 warps at 80 registers: the run loop keeps more state live than a half-sweep launch
constexpr int kClusterMaxCtas = 8;                   // portable cluster size

struct ClusterArgs {
    Geom g;                // geometry of the 16-bit arrays (one GPU: ghost == 0)
    uint64_t sweep_base, first_step, n_steps;
    uint32_t energy_every, action_every, normalize_every;
    uint32_t elog_cap, alog_cap;
    uint32_t e_slot0, a_slot0;  // log slots of the first energy / action measurement of this launch
    uint32_t n_cta;        // CTAs per chain = cluster size
    uint32_t tp;           // planes per CTA
    uint32_t max_thr;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// grid = (n_cta, n_chains), cluster = (n_cta, 1, 1), block = 768 or 384 (a multiple of 32 and of XQ),
// dynamic smem = 2 * (tp + 2) * Y * XQ * 16 bytes.
template <int XQ, bool WILSON, bool COUNT>
__global__ void __launch_bounds__(kClusterMaxThreads, 1)
k_run_cluster(const ClusterArgs a, uint16_t* __restrict__ f0, uint16_t* __restrict__ f1, const ChainDev* __restrict__ chains,
              const uint32_t* __restrict__ thr_all, int32_t* __restrict__ offset, const int32_t* __restrict__ centre,
              long long* __restrict__ elog, long long* __restrict__ alog, unsigned long long* __restrict__ acc) {
    extern __shared__ __align__(16) unsigned char s_field[];
    __shared__ long long s_red[3][kClusterMaxThreads / 32];
    const Geom& g = a.g;
    const uint32_t chain = blockIdx.y, rank = blockIdx.x, n_cta = a.n_cta;
    const uint32_t TP = a.tp, Y = g.Y;
    constexpr uint32_t ROWB = XQ * 16u;                  // bytes per row of one colour
    const uint32_t planeB = Y * ROWB, colB = (TP + 2u) * planeB;
    const ChainDev cd = chains[chain];
    const N16Tab tab = build_interleaved(thr_all + (size_t)chain * kThrStride, a.max_thr, g.entry_bytes);
    const RoundKeys rk = make_round_keys(cd.k0, cd.k1);
    const uint32_t s_col0 = smem_u32(s_field), s_col1 = s_col0 + colB;

    // load: storage plane j of this CTA (j = 0 and TP + 1: the halo planes) is global plane rank * TP + j - 1, periodic
    {
        const uint4* gq[2] = {reinterpret_cast<const uint4*>(f0 + (size_t)chain * g.chain_stride),
                              reinterpret_cast<const uint4*>(f1 + (size_t)chain * g.chain_stride)};
        const uint32_t planeQ = Y * XQ;
        for (uint32_t i = threadIdx.x; i < (TP + 2u) * planeQ; i += blockDim.x) {
            const uint32_t j = i / planeQ, r = i - j * planeQ;
            const uint32_t tg = (rank * TP + j + g.T - 1u) % g.T;
            sts_v4(s_col0 + i * 16u, __ldg(gq[0] + (size_t)tg * planeQ + r));
            sts_v4(s_col1 + i * 16u, __ldg(gq[1] + (size_t)tg * planeQ + r));
        }
    }
    cluster_barrier();  // nobody's halo planes are overwritten by a neighbour before they have been loaded

    // my neighbours' halo planes, as seen from here: plane TP + 1 of the lower neighbour, plane 0 of the upper one
    const uint32_t rank_lo = (rank + n_cta - 1u) % n_cta, rank_hi = (rank + 1u) % n_cta;
    const uint32_t r_lo0 = mapa_u32(s_col0, rank_lo) + (TP + 1u) * planeB, r_lo1 = mapa_u32(s_col1, rank_lo) + (TP + 1u) * planeB;
    const uint32_t r_hi0 = mapa_u32(s_col0, rank_hi), r_hi1 = mapa_u32(s_col1, rank_hi);

    // the thread's rows: (pl_first, y_first), then every rows_per_pass-th row of the CTA's TP * Y rows.  Everything that
    // is linear in the row number advances by a constant per pass: the byte offset within a colour (storage plane
    // pl + 1, row y), the Philox counter group.
    const uint32_t k4 = threadIdx.x & (XQ - 1u), row0 = threadIdx.x / XQ, rows_per_pass = blockDim.x / XQ;
    const uint32_t pl_first = row0 / Y, y_first = row0 - pl_first * Y, dpl = rows_per_pass / Y, dy = rows_per_pass - dpl * Y;
    const uint32_t off_first = (Y + row0) * ROWB + k4 * 16u, off_step = rows_per_pass * ROWB, off_end = (TP + 1u) * planeB;
    const uint32_t grp_first = (rank * TP * Y + row0) * (2u * XQ) + k4 * 2u, grp_step = rows_per_pass * (2u * XQ);
    const uint32_t wrap = (Y - 1u) * ROWB;
    // the x neighbour beyond the thread's eight positions, relative to the thread's own 16 bytes: first word of the
    // column to the right (p) or last word of the column to the left, periodic in the row
    const uint32_t e_right = ((k4 + 1u) & (XQ - 1u)) * 16u - k4 * 16u, e_left = ((k4 - 1u) & (XQ - 1u)) * 16u + 12u - k4 * 16u;
    const uint32_t y_one = 1u % Y;
    int accepted = 0;
    uint32_t e_slot = a.e_slot0, a_slot = a.a_slot0;
    const uint64_t n_sites = (uint64_t)g.X * g.Y * g.T;

    for (uint64_t step = a.first_step; step < a.first_step + a.n_steps; ++step) {
        const uint64_t sweep = a.sweep_base + step;
        const bool due_e = a.energy_every && step % a.energy_every == 0;
        const bool due_a = a.action_every && step % a.action_every == 0;
        long long sum_e = 0, sum_a = 0;
#pragma unroll 1
        for (uint32_t colour = 0; colour < 2; ++colour) {
            const uint32_t own = colour ? s_col1 : s_col0, oth = colour ? s_col0 : s_col1;
            // where the rows of the first / last owned plane go as well: the neighbours' halo planes, minus the
            // offset of those planes here
            const uint32_t r_lo = (colour ? r_lo1 : r_lo0) - planeB, r_hi = (colour ? r_hi1 : r_hi0) - TP * planeB;
            const uint32_t par0 = rank * TP + colour;
            const bool measure = colour == 1u && (due_e || due_a);
            uint32_t y = y_first, pl = pl_first, off = off_first, grp = grp_first;
#pragma unroll 1
            while (off < off_end) {
                const uint32_t p = (y + pl + par0) & 1u;
                const uint32_t off_m = y == 0u ? off + wrap : off - ROWB;
                const uint32_t off_p = y + 1u == Y ? off - wrap : off + ROWB;
                const uint4 v = lds_v4(own + off);
                const uint4 oc = lds_v4(oth + off), om = lds_v4(oth + off_m), on = lds_v4(oth + off_p);
                const uint4 ta = lds_v4(oth + off + planeB), tb = lds_v4(oth + off - planeB);
                const uint32_t edge = lds_u32(oth + off + (p ? e_right : e_left));
                // position j+1 (p) or j-1 (!p) of the other colour's row, without a branch: rows of both parities
                // share a warp here
                const uint32_t wa = p ? oc.x : edge, wb = p ? oc.y : oc.x, wc = p ? oc.z : oc.y, wd = p ? oc.w : oc.z,
                               we = p ? edge : oc.w;
                const uint4 sh = make_uint4(__byte_perm(wa, wb, 0x5432), __byte_perm(wb, wc, 0x5432),
                                            __byte_perm(wc, wd, 0x5432), __byte_perm(wd, we, 0x5432));
                const u32x4 r0 = philox4x32_rk(grp, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
                const u32x4 r1 = philox4x32_rk(grp + 1u, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
                uint4 woff = make_uint4(0, 0, 0, 0);
                if (WILSON) {
                    if (rank * TP + pl < cd.wil_h && (y == 0u || y == y_one))
                        woff = n16_wilson(8u * k4, p, (y == 0u ? 1 : 0) - (y == y_one ? 1 : 0), cd.wil_w);
                }
                uint4 s4, s2;
                n16_row_sums(om, oc, on, ta, tb, sh, s4, s2);
                const uint4 w = n16_row_decide<WILSON, COUNT>(v, s4, s2, r0, r1, tab, woff, accepted);
                sts_v4(own + off, w);
                if (pl == 0u) st_cluster_v4(r_lo + off, w);
                if (pl + 1u == TP) st_cluster_v4(r_hi + off, w);
                if (measure) {
                    if (due_e) sum_e += n16_obs_row<1, 1>(w, s4, s2, oc);
                    if (due_a) sum_a += n16_obs_row<3, -1>(w, s4, s2, oc);
                }
                off += off_step;
                grp += grp_step;
                y += dy;
                pl += dpl;
                if (y >= Y) {
                    y -= Y;
                    ++pl;
                }
            }
            cluster_barrier();
        }
        if (due_e || due_a) {
            // this CTA's share of the sums into the chain's log slots (zeroed by the host before the launch)
            sum_e = warp_sum_ll(sum_e);
            sum_a = warp_sum_ll(sum_a);
            const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
            if (lane == 0u) {
                s_red[0][warp] = sum_e;
                s_red[1][warp] = sum_a;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                long long te = 0, tot_a = 0;
                for (uint32_t i = 0; i < (blockDim.x >> 5); ++i) {
                    te += s_red[0][i];
                    tot_a += s_red[1][i];
                }
                if (due_e && e_slot < a.elog_cap)
                    atomicAdd(reinterpret_cast<unsigned long long*>(elog + (size_t)chain * a.elog_cap + e_slot), (unsigned long long)(2 * te));
                if (due_a && a_slot < a.alog_cap)
                    atomicAdd(reinterpret_cast<unsigned long long*>(alog + (size_t)chain * a.alog_cap + a_slot), (unsigned long long)(2 * tot_a));
            }
            __syncthreads();
            e_slot += due_e ? 1u : 0u;
            a_slot += due_a ? 1u : 0u;
        }
        if (a.normalize_every && step % a.normalize_every == 0 && threadIdx.x == 0) {
            // normalize_random: offset := stored height of the picked site (kernels_v1.cuh, k_pick_shift); the CTA
            // that owns the site writes it
            const uint64_t site = pick_site(cd.k0, cd.k1, 2u, sweep, 0u, n_sites);
            const uint64_t hi = site >> 1;
            const uint32_t tg = (uint32_t)(hi / g.plane), i = (uint32_t)(hi - (uint64_t)tg * g.plane);
            if (tg / TP == rank) {
                const uint32_t yy = i / g.Xh, x = 2u * (i - yy * g.Xh) + (uint32_t)(site & 1u);
                const uint32_t col = (x + yy + tg) & 1u;
                const uint16_t* sp = reinterpret_cast<const uint16_t*>(s_field + (size_t)col * colB + (size_t)(tg - rank * TP + 1u) * planeB);
                offset[chain] = (int32_t)sp[i] - kN16Bias + centre[chain];
            }
        }
    }
    // store the owned planes
    {
        uint4* gq[2] = {reinterpret_cast<uint4*>(f0 + (size_t)chain * g.chain_stride),
                        reinterpret_cast<uint4*>(f1 + (size_t)chain * g.chain_stride)};
        const uint32_t planeQ = Y * XQ;
        for (uint32_t i = threadIdx.x; i < TP * planeQ; i += blockDim.x) {
            gq[0][(size_t)rank * TP * planeQ + i] = lds_v4(s_col0 + planeB + i * 16u);
            gq[1][(size_t)rank * TP * planeQ + i] = lds_v4(s_col1 + planeB + i * 16u);
        }
    }
    if (COUNT) {
        long long t = warp_sum_ll((long long)accepted);
        if ((threadIdx.x & 31u) == 0u) s_red[2][threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = 0;
            for (uint32_t i = 0; i < (blockDim.x >> 5); ++i) tot += s_red[2][i];
            atomicAdd(acc + (size_t)chain * kAccSlots + (rank % kAccSlots), (unsigned long long)tot);
        }
    }
    cluster_barrier();  // no CTA leaves while a neighbour may still store into its shared memory
}

// ---- host side ----------------------------------------------------------------------------------------------
struct ClusterPlan {
    bool enabled = false;
    int xq = 0;
    uint32_t n_cta = 0, tp = 0, threads = 0;
    size_t smem = 0;
};

template <int XQ>
inline const void* cluster_variant(bool w, bool c) {
    if (w) return c ? (const void*)k_run_cluster<XQ, true, true> : (const void*)k_run_cluster<XQ, true, false>;
    return c ? (const void*)k_run_cluster<XQ, false, true> : (const void*)k_run_cluster<XQ, false, false>;
}
inline const void* cluster_kernel(int xq, bool w, bool c) {
    switch (xq) {
        case 1: return cluster_variant<1>(w, c);
        case 2: return cluster_variant<2>(w, c);
        case 4: return cluster_variant<4>(w, c);
        case 8: return cluster_variant<8>(w, c);
    }
    return nullptr;
}

// One chain per cluster of 1, 2, 4 or 8 CTAs (T divisible), X = 16 ... 128, one 768-thread CTA per SM.  Where the path
// pays was measured against the launch path (k_sweep_n16 with the energy inside the half-sweep), chains of 64^3,
// site updates/s, cluster / launch:  1 chain 3.1 / 4.7e10,  4: 1.25 / 1.42e11,  8: 2.49 / 2.20e11,  12: 3.74 / 3.01e11,
// 16 (two 384-thread CTAs per SM, as only 15 clusters of 8 x 768 threads fit the GPCs): 2.97 / 3.96e11,  32: 5.05 /
// 5.23e11;  one chain of 128 x 64 x 64: 3.7 / 8.3e10.  A few chains leave most SMs idle in a cluster while the launch
// path spreads them over the device; many chains fill the device either way and the launch path has more warps per SM.
// So: every chain resident at once, the largest cluster that allows it, and the clusters together on at least 3/8 of
// the SMs; otherwise the handle stays on k_sweep_n16.
inline cudaError_t cluster_prepare(ClusterPlan& p, const Geom& g, uint32_t n_chains, int sm_count, uint32_t max_thr,
                                   size_t smem_optin, bool force_off) {
    p = ClusterPlan{};
    if (force_off || g.ghost != 0 || g.X < 16 || g.X > 128 || (g.X & (g.X - 1)) != 0 || max_thr > (uint32_t)kN16KcMax + 1 ||
        (uint64_t)g.T * g.Y * g.X / 8 >= (1ull << 32))
        return cudaSuccess;
    const size_t budget = smem_optin > (12u << 10) ? smem_optin - (12u << 10) : 0;  // table + reduction scratch + slack
    const size_t plane_bytes = (size_t)g.plane * sizeof(uint16_t);
    uint32_t pick = 0;
    if (const char* e = getenv("LQFT_CLUSTER_CTAS")) {  // tests: force the cluster size (must divide T and fit)
        const uint32_t c = (uint32_t)atoi(e);
        if (c >= 1 && c <= (uint32_t)kClusterMaxCtas && (c & (c - 1)) == 0 && g.T % c == 0 &&
            2 * (size_t)(g.T / c + 2) * plane_bytes <= budget)
            pick = c;
        else
            return cudaSuccess;
    }
    // clusters of c CTAs that the device holds at once (the GPC layout decides, so ask: 15 of 8 x 768 threads on a B200)
    auto resident = [&](uint32_t c, size_t smem) -> int {
        const void* fn = cluster_kernel((int)(g.X / 16), true, true);  // the instantiation with the most registers
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(c, n_chains, 1);
        cfg.blockDim = dim3(kClusterMaxThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = c;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return n;
    };
    for (uint32_t c = (uint32_t)kClusterMaxCtas; c >= 1 && !pick; c /= 2) {  // as many SMs per chain as stay resident together
        if (g.T % c) continue;
        const size_t need = 2 * (size_t)(g.T / c + 2) * plane_bytes;
        if (need > budget || (uint64_t)c * n_chains > (uint64_t)sm_count) continue;
        if ((uint64_t)c * n_chains * 8 < (uint64_t)sm_count * 3) return cudaSuccess;  // too few SMs busy: launch path
        if (resident(c, need) < (int)n_chains) return cudaSuccess;  // (smaller clusters of more chains measured slower still)
        pick = c;
    }
    if (!pick) return cudaSuccess;
    p.xq = (int)(g.X / 16);
    p.n_cta = pick;
    p.tp = g.T / pick;
    p.smem = 2 * (size_t)(p.tp + 2) * plane_bytes;
    p.threads = (uint32_t)kClusterMaxThreads;
    if (const char* e = getenv("LQFT_CLUSTER_THREADS")) p.threads = (uint32_t)atoi(e);  // experiments
    const uint64_t units = (uint64_t)p.tp * g.Y * p.xq;  // 8-site rows per colour and CTA
    while (p.threads > 96u && units < p.threads) p.threads /= 2;
    if (p.threads < 32u * (uint32_t)p.xq || p.threads % (uint32_t)p.xq) return cudaSuccess;
    for (int w = 0; w < 2; ++w)
        for (int c = 0; c < 2; ++c) {
            cudaError_t e = cudaFuncSetAttribute(cluster_kernel(p.xq, w, c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
            if (e != cudaSuccess) return e;
        }
    p.enabled = true;
    if (getenv("LQFT_DEBUG")) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(p.n_cta, n_chains, 1);
        cfg.blockDim = dim3(p.threads);
        cfg.dynamicSmemBytes = p.smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = p.n_cta;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, cluster_kernel(p.xq, false, false), &cfg) != cudaSuccess) cudaGetLastError();
        fprintf(stderr, "[lqft] cluster plan: %u CTAs x %u threads per chain, %zu bytes of shared memory per CTA, %d clusters resident\n",
                p.n_cta, p.threads, p.smem, n);
    }
    return cudaSuccess;
}

inline cudaError_t cluster_launch(const ClusterPlan& p, const ClusterArgs& a, uint32_t n_chains, bool wilson, bool count,
                                  uint16_t* f0, uint16_t* f1, const ChainDev* chains, const uint32_t* thr, int32_t* offset,
                                  const int32_t* centre, long long* elog, long long* alog, unsigned long long* acc,
                                  cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.n_cta, n_chains, 1);
    cfg.blockDim = dim3(p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.n_cta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ClusterArgs args = a;
    void* params[] = {&args, &f0, &f1, &chains, &thr, &offset, &centre, &elog, &alog, &acc};
    return cudaLaunchKernelExC(&cfg, cluster_kernel(p.xq, wilson, count), params);
}

}  // namespace lqft

// kernels_flow.cuh — a whole run of the 16-bit field in ONE persistent launch, half-sweeps ordered by data flow.
//
// The launch path (k_sweep_n16) puts a kernel boundary between two half-sweeps: every block of half-sweep h has to be
// finished, and the whole machine drained, before the first block of h + 1 may read.  At 256^3 that is ~2 us of ramp and
// drain on 11 us of work per half-sweep, at 128^3 (BASELINE config 3, sim.rs:78) three quarters of the time, and every
// block pays its prologue (decision table, round keys, index arithmetic) once per 7 rows.  But a tile of half-sweep h + 1
// needs only a handful of tiles of half-sweep h: the site update reads the six nearest neighbours, nothing else.
//
// Here the grid of k_sweep_n16 — one block per (y chunk, row segment, plane group and parity, chain), one wave, all
// blocks resident (cooperative launch) — stays on the machine for the whole call (computation.rs:229-270: sweep, energy /
// action when due, normalize when due), and each block walks through the half-sweeps of its own tile.  A tile publishes the
// number of half-sweeps it has completed (st.release.gpu after the block's barrier); before half-sweep h a block waits
// (ld.acquire.gpu, one polling thread per neighbour) until the tiles it exchanges rows with have completed h:
//     chunk - 1, chunk + 1      rows y0 - 1 and y1 of the same planes
//     the partner block         planes of the other parity of the same group: the t +- 1 neighbours of all its planes
//     one block of the next /   the t neighbour of the group's first even (last odd) plane
//       previous group
//     segment - 1, segment + 1  X = 1024: the x neighbour across the 512-site segment edge
// That orders reads after the writes they need (half-sweep h reads what h - 1 wrote) and writes after the reads they
// would destroy (half-sweep h overwrites what the neighbours read in h - 1).  No grid barrier, no drain: SMs always hold
// blocks in different phases, and a block's wait overlaps the rows of the three others on its SM.
// Field loads bypass L1 (ld.global.cg): another SM's stores are visible in L2 once its flag is.
// Same sites, same Philox counters, same packed arithmetic as k_sweep_n16 (n16_row_update): bit-identical results.
// Energy and action of a due step ride on its colour-1 half-sweep (n16_obs_row) and are added into the step's log slot;
// normalize is one thread of the block that owns the picked site.
#pragma once
#include "kernels_narrow.cuh"

namespace lqft {

struct FlowArgs {
    Geom g;                 // geometry of the 16-bit arrays (one GPU: ghost == 0)
    uint64_t sweep_base, first_step, n_steps;
    uint32_t energy_every, action_every, normalize_every;
    uint32_t e_phase, a_phase, n_phase;  // first_step mod the three frequencies (0 where a frequency is 0)
    uint32_t elog_cap, alog_cap;
    uint32_t e_slot0, a_slot0;  // log slots of the first energy / action measurement of this launch
    uint32_t n_chunks;
    uint32_t max_thr;
    uint32_t* prog;         // [gridDim.z][gridDim.y][gridDim.x] half-sweeps completed by each tile; zero at launch
};

__device__ unsigned long long g_flow_timeout_ns = 5000000000ull;

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until the tile at p has completed `need` half-sweeps.  All blocks are resident, so the wait is microseconds;
// a wait that lasts seconds is a bug or a dead device: trap rather than hang (the context is lost, every later call
// fails with LQFT_ERR_CUDA).
__device__ __forceinline__ void flow_wait(const uint32_t* p, uint32_t need) {
    if (ld_acquire_gpu_u32(p) >= need) return;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_gpu_u32(p) < need) {
        if (global_timer_ns() - t0 > *reinterpret_cast<volatile unsigned long long*>(&g_flow_timeout_ns)) __trap();
    }
}

#ifndef LQFT_FLOW_LD
#define LQFT_FLOW_LD(p) __ldcg(p)
#endif

constexpr int kFlowNbrs = 6;

// grid = (n_chunks * max(SEGS, 1), 2 * ceil(Tl / P), n_chains), P = 2 * warps * planes per warp; block = 32 * warps;
// cooperative launch (every block resident).  Tile geometry as in k_sweep_n16.
// SINGLE: a lone chain's parameters arrive as a kernel argument (round keys in uniform registers, see k_sweep_n16).
template <int XQ, bool WILSON, bool COUNT, bool SINGLE>
__global__ void __launch_bounds__(kN16Threads, 4)
k_run_flow(const FlowArgs a, uint16_t* __restrict__ f0, uint16_t* __restrict__ f1, const ChainDev* __restrict__ chains,
           const ChainDev chain0, const uint32_t* __restrict__ thr_all, int32_t* __restrict__ offset, const int32_t* __restrict__ centre,
           long long* __restrict__ elog, long long* __restrict__ alog, unsigned long long* __restrict__ acc) {
    constexpr uint32_t ROWQ = XQ;                        // uint4 per row of one colour
    constexpr uint32_t LPR = XQ < 32 ? XQ : 32;          // lanes of a warp on one row
    constexpr uint32_t PPW = 32u / LPR;                  // planes per warp
    constexpr int SEGS = XQ / 32;                        // 512-site row segments (0: a row is narrower than a warp)
    constexpr uint32_t NSEG = SEGS > 1 ? SEGS : 1;
    __shared__ unsigned int s_cnt;
    __shared__ long long s_red[2][kN16Warps];
    const Geom& g = a.g;
    const uint32_t chain = SINGLE ? 0u : blockIdx.z;
    const ChainDev cd = SINGLE ? chain0 : chains[chain];
    if (COUNT && threadIdx.x == 0) s_cnt = 0;
    const N16Tab tab = build_interleaved(thr_all + (size_t)chain * kThrStride, a.max_thr, g.entry_bytes);

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t seg = SEGS > 1 ? blockIdx.x % NSEG : 0u, chunk = SEGS > 1 ? blockIdx.x / NSEG : blockIdx.x;
    const uint32_t par = blockIdx.y & 1u, grp = blockIdx.y >> 1, n_grp = gridDim.y >> 1;
    const uint32_t P = 2u * PPW * n_warps;               // planes of a group (both parities)
    const uint32_t slot = PPW * warp + lane / LPR;       // plane slot within the block
    const uint32_t pl_raw = grp * P + 2u * slot + par;
    const uint32_t y0 = chunk * g.Y / a.n_chunks, y1 = (chunk + 1u) * g.Y / a.n_chunks;
    const bool live = pl_raw < g.Tl;                     // surplus plane slots: plane 0, stores off
    const uint32_t tl = live ? pl_raw : 0u;              // one GPU: local plane == global plane
    const uint32_t kq = SEGS ? seg * 32u + lane : lane & (LPR - 1u);  // uint4 column of this thread within the row
    const uint32_t planeq = g.Y * ROWQ;
    const uint32_t cq = chain * (uint32_t)(g.chain_stride / 8);
    const uint32_t r_own = cq + tl * planeq;
    const uint32_t b_own = r_own + kq;
    const uint32_t b_ta = cq + (tl + 1u == g.Tl ? 0u : tl + 1u) * planeq + kq, b_tb = cq + (tl == 0u ? g.Tl - 1u : tl - 1u) * planeq + kq;
    uint4* const q0 = reinterpret_cast<uint4*>(f0);
    uint4* const q1 = reinterpret_cast<uint4*>(f1);
    const uint32_t e_right = (r_own + (kq + 1u) % ROWQ) * 4u, e_left = (r_own + (kq + ROWQ - 1u) % ROWQ) * 4u + 3u;
    const RoundKeys rk = make_round_keys(cd.k0, cd.k1);
    const uint32_t src_r = (lane & ~(LPR - 1u)) | ((lane + 1u) & (LPR - 1u)), src_l = (lane & ~(LPR - 1u)) | ((lane - 1u) & (LPR - 1u));
    const uint32_t grp_base = tl * g.Y * (2u * ROWQ) + kq * 2u;
    const bool wil_plane = WILSON && tl < cd.wil_h;
    const uint32_t y_one = 1u % g.Y;
    // a warp without a live plane slot does one row per half-sweep (uniform trip count by vote, as in k_sweep_n16)
    const uint32_t n_rows = __all_sync(0xffffffffu, pl_raw >= g.Tl) ? 1u : y1 - y0;

    // the tiles this one exchanges rows with (duplicates and the tile itself are harmless)
    const uint32_t tiles_per_chain = gridDim.x * gridDim.y;
    uint32_t* const prog_chain = a.prog + (size_t)chain * tiles_per_chain;
    uint32_t* const prog_mine = prog_chain + blockIdx.y * gridDim.x + blockIdx.x;
    const uint32_t* prog_nbr = prog_mine;
    {
        const uint32_t n_ch = a.n_chunks;
        const uint32_t c_lo = (chunk + n_ch - 1u) % n_ch, c_hi = (chunk + 1u) % n_ch;
        const uint32_t g_adj = par ? (grp + 1u) % n_grp : (grp + n_grp - 1u) % n_grp;
        uint32_t bx = blockIdx.x, by = blockIdx.y;
        switch (threadIdx.x) {
            case 0: bx = c_lo * NSEG + seg; break;
            case 1: bx = c_hi * NSEG + seg; break;
            case 2: by = blockIdx.y ^ 1u; break;
            case 3: by = 2u * g_adj + (par ^ 1u); break;
            case 4: bx = chunk * NSEG + (seg + NSEG - 1u) % NSEG; break;
            case 5: bx = chunk * NSEG + (seg + 1u) % NSEG; break;
            default: break;
        }
        prog_nbr = prog_chain + by * gridDim.x + bx;
    }

    int accepted = 0;
    uint32_t e_slot = a.e_slot0, a_slot = a.a_slot0;
    uint32_t e_cnt = a.e_phase, a_cnt = a.a_phase, n_cnt = a.n_phase;  // step mod frequency, kept by counting
    const uint64_t n_sites = (uint64_t)g.X * g.Y * g.T;
    const uint32_t n_half = 2u * (uint32_t)a.n_steps;

#define LQFT_FLOW_STEP(OM, OC, ON, MEAS)                                                                                 \
    {                                                                                                                    \
        const uint32_t yn = (y + 1 == g.Y) ? 0 : y + 1;                                                                  \
        uint4* const po = q_own + (b_own + y * ROWQ);                                                                    \
        const uint4 v = LQFT_FLOW_LD(po);                                                                                \
        ON = LQFT_FLOW_LD(q_oth + (b_own + yn * ROWQ));                                                                  \
        const uint4 ta = LQFT_FLOW_LD(q_oth + (b_ta + y * ROWQ));                                                        \
        const uint4 tb = LQFT_FLOW_LD(q_oth + (b_tb + y * ROWQ));                                                        \
        uint32_t e_ld = 0;                                                                                               \
        const bool e_lane = SEGS > 1 && (p ? lane == 31u : lane == 0u);                                                  \
        if (SEGS > 1) {                                                                                                  \
            if (e_lane) e_ld = LQFT_FLOW_LD(w_oth + ((p ? e_right : e_left) + y * (ROWQ * 4u)));                         \
        }                                                                                                                \
        const uint32_t grp_ctr = grp_base + y * (2u * ROWQ);                                                             \
        const u32x4 r0 = philox4x32_rk(grp_ctr, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);                   \
        const u32x4 r1 = philox4x32_rk(grp_ctr + 1u, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);              \
        uint4 woff = make_uint4(0, 0, 0, 0);                                                                             \
        if (WILSON) {                                                                                                    \
            if (wil_plane && (y == 0u || y == y_one))                                                                    \
                woff = n16_wilson(8u * kq, p, (y == 0u ? 1 : 0) - (y == y_one ? 1 : 0), cd.wil_w);                       \
        }                                                                                                                \
        uint4 w, s4, s2;                                                                                                 \
        int acc_row = 0;                                                                                                 \
        if (p) {                                                                                                         \
            uint32_t edge = __shfl_sync(0xffffffffu, OC.x, src_r);                                                       \
            if (SEGS > 1) edge = e_lane ? e_ld : edge;                                                                   \
            w = n16_row_update<1, WILSON, COUNT>(v, OM, OC, ON, ta, tb, edge, r0, r1, tab, woff, acc_row, s4, s2);       \
        } else {                                                                                                         \
            uint32_t edge = __shfl_sync(0xffffffffu, OC.w, src_l);                                                       \
            if (SEGS > 1) edge = e_lane ? e_ld : edge;                                                                   \
            w = n16_row_update<0, WILSON, COUNT>(v, OM, OC, ON, ta, tb, edge, r0, r1, tab, woff, acc_row, s4, s2);       \
        }                                                                                                                \
        if (COUNT) accepted += live ? acc_row : 0;                                                                       \
        if (live) *po = w;                                                                                               \
        if (MEAS) {                                                                                                      \
            if (due_e) sum_e += live ? n16_obs_row<1, 1>(w, s4, s2, OC) : 0;                                             \
            if (due_a) sum_a += live ? n16_obs_row<3, -1>(w, s4, s2, OC) : 0;                                            \
        }                                                                                                                \
        y = yn;                                                                                                          \
        p ^= 1u;                                                                                                         \
    }

#pragma unroll 1
    for (uint32_t hs = 0; hs < n_half; ++hs) {
        const uint32_t colour = hs & 1u;
        const uint64_t sweep = a.sweep_base + a.first_step + (hs >> 1);
        const bool due_e = a.energy_every && e_cnt == 0u;
        const bool due_a = a.action_every && a_cnt == 0u;
        const bool measure = colour == 1u && (due_e || due_a);
#ifndef LQFT_FLOW_NOWAIT
        if (threadIdx.x < (uint32_t)kFlowNbrs) flow_wait(prog_nbr, hs);
#endif
        __syncthreads();  // the acquire of the polling threads reaches the block; the compiler sees it converge
        uint4* const q_own = colour ? q1 : q0;
        const uint4* const q_oth = colour ? q0 : q1;
        const uint32_t* const w_oth = reinterpret_cast<const uint32_t*>(q_oth);
        long long sum_e = 0, sum_a = 0;
        uint32_t y = y0;
        const uint32_t ym = (y == 0) ? g.Y - 1 : y - 1;
        uint4 wa = LQFT_FLOW_LD(q_oth + (b_own + ym * ROWQ)), wb = LQFT_FLOW_LD(q_oth + (b_own + y * ROWQ)), wc;
        uint32_t p = (y0 + par + colour) & 1u;
        uint32_t n = n_rows;
        if (measure) {
            for (;;) {
                LQFT_FLOW_STEP(wa, wb, wc, true)
                if (--n == 0) break;
                LQFT_FLOW_STEP(wb, wc, wa, true)
                if (--n == 0) break;
                LQFT_FLOW_STEP(wc, wa, wb, true)
                if (--n == 0) break;
            }
            sum_e = warp_sum_ll(sum_e);
            sum_a = warp_sum_ll(sum_a);
            if (lane == 0u) {
                s_red[0][warp] = sum_e;
                s_red[1][warp] = sum_a;
            }
        } else {
            for (;;) {
                LQFT_FLOW_STEP(wa, wb, wc, false)
                if (--n == 0) break;
                LQFT_FLOW_STEP(wb, wc, wa, false)
                if (--n == 0) break;
                LQFT_FLOW_STEP(wc, wa, wb, false)
                if (--n == 0) break;
            }
        }
        __syncthreads();  // every row of the tile is stored
        if (threadIdx.x == 0) {
            st_release_gpu_u32(prog_mine, hs + 1u);
            if (measure) {
                // this block's share of the sums into the chain's log slots (zeroed by the host before the launch)
                long long te = 0, tot_a = 0;
                for (uint32_t i = 0; i < n_warps; ++i) {
                    te += s_red[0][i];
                    tot_a += s_red[1][i];
                }
                if (due_e && e_slot < a.elog_cap && te != 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(elog + (size_t)chain * a.elog_cap + e_slot), (unsigned long long)(2 * te));
                if (due_a && a_slot < a.alog_cap && tot_a != 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(alog + (size_t)chain * a.alog_cap + a_slot), (unsigned long long)(2 * tot_a));
            }
            if (colour == 1u && a.normalize_every && n_cnt == 0u) {
                // normalize_random: offset := stored height of the picked site (kernels_v1.cuh, k_pick_shift); the block
                // whose tile holds the site has just stored both colours of it
                const uint64_t site = pick_site(cd.k0, cd.k1, 2u, sweep, 0u, n_sites);
                const uint64_t hi = site >> 1;
                const uint32_t tg = (uint32_t)(hi / g.plane), i = (uint32_t)(hi - (uint64_t)tg * g.plane);
                const uint32_t yy = i / g.Xh, k = i - yy * g.Xh, x = 2u * k + (uint32_t)(site & 1u);
                const bool mine = tg / P == grp && (tg & 1u) == par && yy >= y0 && yy < y1 && (SEGS > 1 ? k / 256u == seg : true);
                if (mine) {
                    const uint32_t col = (x + yy + tg) & 1u;
                    const uint16_t* sp = (col ? f1 : f0) + (size_t)chain * g.chain_stride + (size_t)tg * g.plane + i;
                    offset[chain] = (int32_t)__ldcg(sp) - kN16Bias + centre[chain];
                }
            }
        }
        if (colour == 1u) {
            e_slot += due_e ? 1u : 0u;
            a_slot += due_a ? 1u : 0u;
            e_cnt = e_cnt + 1u == a.energy_every ? 0u : e_cnt + 1u;
            a_cnt = a_cnt + 1u == a.action_every ? 0u : a_cnt + 1u;
            n_cnt = n_cnt + 1u == a.normalize_every ? 0u : n_cnt + 1u;
        }
    }
#undef LQFT_FLOW_STEP
    if (COUNT)
        block_count_flush(&s_cnt, (unsigned int)accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
}

// ---- host side ----------------------------------------------------------------------------------------------
struct FlowPlan {
    bool enabled = false;
    uint32_t n_chunks = 0, warps = 0, groups = 0, blocks_per_sm = 0;
    dim3 grid;
    uint32_t* prog = nullptr;
};

template <int XQ, bool S>
inline const void* flow_variant(bool w, bool c) {
    if (w) return c ? (const void*)k_run_flow<XQ, true, true, S> : (const void*)k_run_flow<XQ, true, false, S>;
    return c ? (const void*)k_run_flow<XQ, false, true, S> : (const void*)k_run_flow<XQ, false, false, S>;
}
inline const void* flow_kernel(int xq, bool w, bool c, bool s) {
    switch (xq) {
        case 1: return s ? flow_variant<1, true>(w, c) : flow_variant<1, false>(w, c);
        case 2: return s ? flow_variant<2, true>(w, c) : flow_variant<2, false>(w, c);
        case 4: return s ? flow_variant<4, true>(w, c) : flow_variant<4, false>(w, c);
        case 8: return s ? flow_variant<8, true>(w, c) : flow_variant<8, false>(w, c);
        case 16: return s ? flow_variant<16, true>(w, c) : flow_variant<16, false>(w, c);
        case 32: return s ? flow_variant<32, true>(w, c) : flow_variant<32, false>(w, c);
        case 64: return s ? flow_variant<64, true>(w, c) : flow_variant<64, false>(w, c);
    }
    return nullptr;
}

// The path is taken when the whole grid is resident at once: as many y chunks per plane group as the block slots of
// the device allow (at least two rows per chunk), every chain of the handle in the same launch.
inline cudaError_t flow_prepare(FlowPlan& f, NarrowPlan& p, const Geom& g, uint32_t n_chains, bool force_off) {
    f.enabled = false;
    if (force_off || !p.enabled || g.ghost != 0) return cudaSuccess;
    if (const char* e = getenv("LQFT_FLOW")) {
        if (atoi(e) == 0) return cudaSuccess;
    }
    f.warps = narrow_warps(p, g.Tl);
    f.groups = narrow_groups(p, g.Tl);
    int occ = 1 << 30;
    for (int w = 0; w < 2; ++w)
        for (int c = 0; c < 2; ++c) {
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, flow_kernel(p.xq, w, c, n_chains == 1), (int)(32 * f.warps), 0);
            if (e != cudaSuccess) return e;
            occ = n < occ ? n : occ;
        }
    if (occ < 1) return cudaSuccess;
    f.blocks_per_sm = (uint32_t)occ;
    const uint64_t slots = (uint64_t)p.sm_count * occ;
    const uint64_t base = (uint64_t)f.groups * (p.segs() ? p.segs() : 1) * n_chains;  // blocks per chunk
    uint64_t nc = slots / base;
    const uint32_t max_chunks = g.Y / 2 ? g.Y / 2 : 1;
    if (nc > max_chunks) nc = max_chunks;
    if (p.forced_chunks) nc = std::min<uint64_t>(nc, p.forced_chunks);
    if (nc < 1) return cudaSuccess;
    f.n_chunks = (uint32_t)nc;
    f.grid = dim3(f.n_chunks * (p.segs() ? p.segs() : 1), f.groups, n_chains);
    const size_t tiles = (size_t)f.grid.x * f.grid.y * f.grid.z;
    if (f.prog) cudaFree(f.prog);
    f.prog = nullptr;
    cudaError_t e = cudaMalloc(&f.prog, sizeof(uint32_t) * tiles);
    if (e != cudaSuccess) return e;
    f.enabled = true;
    return cudaSuccess;
}

inline cudaError_t flow_launch(const FlowPlan& f, const NarrowPlan& p, FlowArgs a, bool wilson, bool count, uint16_t* f0, uint16_t* f1,
                               const ChainDev* chains, ChainDev chain0, const uint32_t* thr, int32_t* offset, const int32_t* centre, long long* elog,
                               long long* alog, unsigned long long* acc, cudaStream_t stream) {
    const size_t tiles = (size_t)f.grid.x * f.grid.y * f.grid.z;
    cudaError_t e = cudaMemsetAsync(f.prog, 0, sizeof(uint32_t) * tiles, stream);
    if (e != cudaSuccess) return e;
    a.n_chunks = f.n_chunks;
    a.prog = f.prog;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = f.grid;
    cfg.blockDim = dim3(32 * f.warps);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    a.e_phase = a.energy_every ? (uint32_t)(a.first_step % a.energy_every) : 0u;
    a.a_phase = a.action_every ? (uint32_t)(a.first_step % a.action_every) : 0u;
    a.n_phase = a.normalize_every ? (uint32_t)(a.first_step % a.normalize_every) : 0u;
    void* params[] = {&a, &f0, &f1, &chains, &chain0, &thr, &offset, &centre, &elog, &alog, &acc};
    return cudaLaunchKernelExC(&cfg, flow_kernel(p.xq, wilson, count, f.grid.z == 1), params);
}

}  // namespace lqft

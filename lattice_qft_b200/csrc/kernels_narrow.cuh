// kernels_narrow.cuh — red-black Metropolis half-sweep on 16-bit heights (X a power of two, 16 ... 1024).
//
// The reference keeps heights in i32 (Field<i32>, fields.rs:63-67) and so does the ABI.  On the device the field of an
// admitted lattice is STORED as u16: heights biased by kN16Bias relative to a per-chain centre (the chain's
// normalisation offset when the arrays were last centred); hot start, upload, download and every observable work on
// these arrays, i32 arrays exist only for a field that has left the band (lqft.cu).  Half the bytes per site, and —
// what matters more on this instruction-bound kernel — half the registers per site: one thread owns 8 same-colour
// sites (one 16-byte load), 16 lanes a row segment of 128 positions.
//
// Same decision spec as every other kernel family (spec.cuh: Philox keying by global index, integer
// threshold table), so results are bit-identical to the i32 kernels and to the oracle; what differs is the
// arithmetic used to get there:
//   * neighbour sums are formed on packed pairs: with heights in [0, 4096) a 6-term sum stays below 2^16, so
//     plain 32-bit adds never carry from the low half into the high half;
//   * u = 6 v + 0x6000 - sum holds d + 24576 per half, again without borrow; one VIADDMNMX.S16x2.RELU
//     (add.s16x2 + min.s16x2.relu) turns both halves into the clamped table index d + kc in [0, 2 kc];
//   * the coin bit of the site's Philox word is funnel-shifted into the index, so one LDS.64 fetches
//     {threshold | coin << 31, +-1} for the proposed direction.
//
// Slabs (world_size > 1): the narrow copy carries `depth` (up to kN16GhostDepth) ghost planes per side instead of
// one, in TWO sets (Geom::gbuf selects the current one).  While a set is current, each half-sweep also updates the
// ghost planes that still have valid t-neighbours — redundantly, with the same Philox words as their owner, hence
// identically — and the valid depth shrinks by one; after `depth` half-sweeps it is used up.  Meanwhile the last two
// half-sweeps of such a period (one per colour) store their first and last `depth` owned planes a second time:
// straight into the OTHER ghost set of the two neighbours, over NVLink (PUSH instantiation, peer memory mapped
// through CUDA IPC).  A period therefore ends with both sets complete and needs no copy kernel, no second stream
// and no system-scope fence: the first half-sweep of the next period raises this rank's flag at the neighbours in
// its prologue (its predecessor's completion has flushed the pushed rows), waits for theirs, and reads the other
// set.  Writing into the set that is NOT current is race free: a neighbour can only be in the same period (it
// cannot flip ahead without this rank's flag), where it reads and redundantly updates its current set only.
// k_n16_fill copies the boundary planes for the first period after the copy was (re)built.
//
// The ghost planes a half-sweep updates redundantly are dealt to the warps as ghost tasks (k_sweep_n16, below).
//
// Validity: a conversion or re-centring measures the largest |height - centre| = w; a sweep moves a site by at most
// one, so the arrays are good for kN16Bias - 1 - w more sweeps (the sweep budget, host logic in lqft.cu: narrow_enter /
// narrow_recheck / narrow_exit); they are entered or re-centred only with every height within kN16Guard of the centre.
#pragma once
#include "kernels_common.cuh"
#include "kernels_observe.cuh"

namespace lqft {

constexpr int kN16Warps = 8;                     // per block, at most (9 warps at 56 registers: 6 % slower, measured)
constexpr int kN16Threads = kN16Warps * 32;
constexpr int kN16Bias = 2048;
constexpr int kN16Guard = 1023;
constexpr uint32_t kN16Segment = 1000;           // sweeps a refused field stays on i32 before the band is tried again
constexpr int kN16KcMax = 255;                   // threshold tables up to 256 entries (beta >= ~0.045)

constexpr uint32_t kN16GhostDepth = 8;           // ghost planes per side of a slab's narrow copy (even); fewer for large planes
constexpr uint32_t kN16FillBlocks = 32;          // per (side, colour, chain)

// Decision table.  Entry (i << 1) | coin serves d = i - kc: coin 1 proposes +1 (m = d), coin 0 proposes -1 (m = -d); same
// contents as build_two_sided (kernels_common.cuh), interleaved so that the coin can be shifted into the index; an entry
// is {threshold | coin << 31, +-1}.  (The lanes of a warp look up different entries: a 4.4-way bank conflict per LDS.64
// on average, ncu round 2.  Storing the table [entry][copy] with one copy per lane of a half-warp removes the conflicts
// and made the kernel 2 % slower — 24 KB of table per block instead of 8 cost more L1 and prologue than the conflicts
// cost LSU time, measured with 1, 8 and 16 copies.  One copy it is.)
__shared__ uint2 s_n16_tab[2 * (2 * kN16KcMax + 1)];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct N16Tab {
    uint32_t base;     // shared-memory byte address of entry 0
    uint32_t stride;   // bytes per entry, as a value the compiler cannot fold (n16_lookup)
    uint32_t c_off;    // packed kc - 24576: u + c_off = d + kc per half
    uint32_t c_max;    // packed 2 kc
};

__device__ __forceinline__ N16Tab build_interleaved(const uint32_t* __restrict__ thr, uint32_t n_thr, uint32_t entry_bytes) {
    const int kc = max((int)n_thr - 1, 3);
    for (int i = threadIdx.x; i < 2 * (2 * kc + 1); i += blockDim.x) {
        const int up = i & 1, d = (i >> 1) - kc;
        const int m = up ? d : -d;
        const int k = min(max(m + 3, 0), (int)n_thr - 1);
        s_n16_tab[i] = make_uint2(thr[k] | (up ? 0x80000000u : 0u), up ? 1u : 0xffffffffu);
    }
    __syncthreads();
    N16Tab t;
    t.base = smem_u32(s_n16_tab);
    t.stride = entry_bytes;
    t.c_off = (uint32_t)((kc - 24576) & 0xffff) * 0x10001u;
    t.c_max = (uint32_t)(2 * kc) * 0x10001u;
    return t;
}

// Entry i.  With the entry size arriving as a kernel argument the address is an integer multiply-add (FMA pipe) where a
// compile-time 8 gives a LEA: the ALU pipe is the loaded one (ncu: ALU 55 %, FMA 20 % of their peaks before the change).
__device__ __forceinline__ uint2 n16_lookup(const N16Tab& tab, const uint32_t i) {
    uint2 t;
    const uint32_t a = i * tab.stride + tab.base;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(a));
    return t;
}

// w += (r <= t.x) ? t.y << SH : 0 as a compare and a predicated (multiply-)add; the compiler's own choice for the C
// expression is two selects, a shift and a three-input add per pair.
template <int SH>
__device__ __forceinline__ void n16_step_if(uint32_t& w, const uint32_t r, const uint2 t) {
    if (SH == 0)
        asm("{\n\t.reg .pred p;\n\tsetp.ls.u32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(w) : "r"(r), "r"(t.x), "r"(t.y));
    else
        asm("{\n\t.reg .pred p;\n\tsetp.ls.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, 65536, %0;\n\t}" : "+r"(w) : "r"(r), "r"(t.x), "r"(t.y));
}

// Decisions of one row of 8 sites per thread.  s4 = packed sum of the four x / y neighbours, s2 = of the two t neighbours.
template <bool WILSON, bool COUNT>
__device__ __forceinline__ uint4 n16_row_decide(const uint4 v, const uint4 s4, const uint4 s2, const u32x4 r0, const u32x4 r1,
                                                const N16Tab& tab, const uint4 woff, int& accepted) {
    constexpr uint32_t C = 0x60006000u;
    uint4 w;
#define LQFT_N16_PAIR(F, RA, RB)                                                                      \
    {                                                                                                 \
        uint32_t u; /* 6 v + C - s4 - s2, the constant folded into the subtraction (one add less than the compiler's form) */ \
        asm("{\n\t.reg .u32 t;\n\tsub.u32 t, %3, %1;\n\tsub.u32 t, t, %2;\n\tmad.lo.u32 %0, %4, 6, t;\n\t}"   \
            : "=r"(u) : "r"(s4.F), "r"(s2.F), "n"(C), "r"(v.F));                                      \
        if (WILSON) u += woff.F;                                                                      \
        const uint32_t idx = __viaddmin_s16x2_relu(u, tab.c_off, tab.c_max);                          \
        const uint2 ta = n16_lookup(tab, __funnelshift_l(RA, idx & 0xffffu, 1));                      \
        const uint2 tb = n16_lookup(tab, __funnelshift_l(RB, idx >> 16, 1));                          \
        if (COUNT) {                                                                                  \
            const bool aa = (RA) <= ta.x, ab = (RB) <= tb.x;                                          \
            accepted += (int)aa + (int)ab;                                                            \
            w.F = v.F + (aa ? ta.y : 0u) + ((ab ? tb.y : 0u) << 16);                                  \
        } else {                                                                                      \
            w.F = v.F;                                                                                \
            n16_step_if<0>(w.F, RA, ta);                                                              \
            n16_step_if<16>(w.F, RB, tb);                                                             \
        }                                                                                             \
    }
    LQFT_N16_PAIR(x, r0.x, r0.y)
    LQFT_N16_PAIR(y, r0.z, r0.w)
    LQFT_N16_PAIR(z, r1.x, r1.y)
    LQFT_N16_PAIR(w, r1.z, r1.w)
#undef LQFT_N16_PAIR
    return w;
}

// Packed neighbour sums of a row: sh = the other x neighbour of site j (position j+1 or j-1 of the other colour's row,
// by the row's x parity).  Heights below 4096: a half never carries into the other.
__device__ __forceinline__ void n16_row_sums(const uint4 om, const uint4 oc, const uint4 on, const uint4 to, const uint4 ti,
                                             const uint4 sh, uint4& s4, uint4& s2) {
    s4 = make_uint4(oc.x + sh.x + om.x + on.x, oc.y + sh.y + om.y + on.y, oc.z + sh.z + om.z + on.z, oc.w + sh.w + om.w + on.w);
    s2 = make_uint4(to.x + ti.x, to.y + ti.y, to.z + ti.z, to.w + ti.w);
}

// The same with the x parity P of the row's sites known at compile time (warp uniform in k_sweep_n16).
template <int P, bool WILSON, bool COUNT>
__device__ __forceinline__ uint4 n16_row_update(const uint4 v, const uint4 om, const uint4 oc, const uint4 on, const uint4 to,
                                                const uint4 ti, const uint32_t edge, const u32x4 r0, const u32x4 r1,
                                                const N16Tab& tab, const uint4 woff, int& accepted,
                                                uint4& s4, uint4& s2) {
    uint4 sh;
    if (P) {
        sh.x = __byte_perm(oc.x, oc.y, 0x5432); sh.y = __byte_perm(oc.y, oc.z, 0x5432);
        sh.z = __byte_perm(oc.z, oc.w, 0x5432); sh.w = __byte_perm(oc.w, edge, 0x5432);
    } else {
        sh.x = __byte_perm(edge, oc.x, 0x5432); sh.y = __byte_perm(oc.x, oc.y, 0x5432);
        sh.z = __byte_perm(oc.y, oc.z, 0x5432); sh.w = __byte_perm(oc.z, oc.w, 0x5432);
    }
    n16_row_sums(om, oc, on, to, ti, sh, s4, s2);
    return n16_row_decide<WILSON, COUNT>(v, s4, s2, r0, r1, tab, woff, accepted);
}

// Energy / action from inside a colour-1 half-sweep.  Every bond joins a colour-1 site and a colour-0 site, so
//     sum_bonds s_b (h_x - h_y)^2 = sum_{x in colour 1} sum_{6 nbrs n} s (w - n)^2        (s = +1, or -1 for t bonds
//                                                                                          of the energy, fields.rs:728-737)
//   = sum_x [ c w^2 - 2 w (S4 +- S2) ] + c sum_{colour 0} h^2,   S4 = sum of the x, y neighbours, S2 = of the t ones,
// c = 4 - 2 for the energy and 4 + 2 for the action (fields.rs:719-726): w, S4, S2 and the colour-0 row of the same
// positions are in registers when a colour-1 row has been decided.  An exact integer identity, evaluated on the
// stored (biased) heights; the eight terms of a row are summed in 32 bits (|.| < 2^31), then widened.
// acc += sum over the two 16-bit halves of  w (K w - s4 + SGN s2) + K oc^2:  K = 1, SGN = +1 gives half the energy share
// of two colour-1 sites, K = 3, SGN = -1 half their action share.
template <int K, int SGN>
__device__ __forceinline__ void n16_obs_pair(uint32_t w, uint32_t s4, uint32_t s2, uint32_t oc, int& acc) {
    const int w0 = (int)(w & 0xffffu), w1 = (int)(w >> 16);
    const int o0 = (int)(oc & 0xffffu), o1 = (int)(oc >> 16);
    const int t0 = K * w0 - (int)(s4 & 0xffffu) + SGN * (int)(s2 & 0xffffu);
    const int t1 = K * w1 - (int)(s4 >> 16) + SGN * (int)(s2 >> 16);
    acc += w0 * t0 + K * o0 * o0 + w1 * t1 + K * o1 * o1;
}
template <int K, int SGN>
__device__ __forceinline__ int n16_obs_row(const uint4 w, const uint4 s4, const uint4 s2, const uint4 oc) {
    int e = 0;
    n16_obs_pair<K, SGN>(w.x, s4.x, s2.x, oc.x, e);
    n16_obs_pair<K, SGN>(w.y, s4.y, s2.y, oc.y, e);
    n16_obs_pair<K, SGN>(w.z, s4.z, s2.z, oc.z, e);
    n16_obs_pair<K, SGN>(w.w, s4.w, s2.w, oc.w, e);
    return e;
}

// Packed Wilson offsets of the 8 sites x = 2 (k0 + j) + p of a row next to the sheet (spec.cuh, wilson_offset).
__device__ __forceinline__ uint4 n16_wilson(uint32_t k0, uint32_t p, int sgn, uint32_t wil_w) {
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t xa = 2u * (k0 + 2u * i) + p, xb = xa + 2u;
        const int lo = xa < wil_w ? sgn : 0, hi = xb < wil_w ? sgn : 0;
        o[i] = (uint32_t)lo + ((uint32_t)hi << 16);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- the half-sweep: X = 16 ... 1024 (powers of two); XQ = X / 16 = uint4 per row of one colour ---------------------
// A thread owns 8 same-colour sites of a row; LPR = min(XQ, 32) lanes cover a row (or a 512-site segment of it), so a
// warp carries 32 / LPR planes of the same parity: 1 for X >= 512, 2 for X = 256, 8 for X = 64.  X = 1024 has two
// segments per row; the x-neighbour across the segment edge is one predicated 4-byte load.  Per step and
// thread: own row (read + write), the other colour's rows y+1 (y-1 and y stay in a rotating register window), t+1
// and t-1: 4 LDG.128 + 1 STG.128, two Philox blocks, 8 decisions.
//
// Updates n_planes local planes: the first n_first of them start at tl_first, the rest at tl_second (two wedges
// next to the slab boundaries in one launch; one range when n_first == n_planes).  Negative plane numbers reach
// into the lower ghost planes, numbers >= Tl into the upper ones.
//
// Control flow is kept visibly uniform: ptxas uses the uniform datapath (Philox round keys in uniform registers)
// and plain shuffles only where it can prove convergence, and it does not know that threadIdx.x >> 5 is warp
// uniform.  Hence (a) no guard around the body — surplus warps past the end of the plane range walk plane 0 with
// their stores predicated off, (b) the branch on the x parity depends on blockIdx only: blocks 2m and 2m+1 share
// P consecutive planes, the even ones in block 2m, the odd ones in 2m+1 (P = 2 * warps per block * planes per
// warp), (c) SINGLE: a lone chain's parameters arrive as a kernel argument, so that the key is a
// parameter-bank constant.  Without these the same source compiles to 1400 instead of 1024 instructions and runs
// 18 % slower.
// grid = (n_chunks * max(SEGS, 1), 2 * ceil(n_planes / P), n_chains); block = 32 * warps, warps <= kN16Warps chosen by
// narrow_warps so that small lattices do not carry idle warps.
// Slab launches: where the boundary planes go (PUSH) and the flip of the ghost sets in the prologue (sync).
struct SlabArgs {
    uint16_t* lo;            // lower neighbour's array of the colour being updated (chain 0), as mapped here
    uint16_t* hi;            // upper neighbour's
    uint32_t lo_stride8, hi_stride8;  // their uint4 per chain
    uint32_t lo_plane, hi_plane;      // storage plane there of my local plane 0 / of my local plane Tl - depth
    uint32_t depth;          // planes pushed per side
    uint32_t extra;          // ghost planes per side this half-sweep updates as well (ghost tasks, below)
    uint32_t sync;           // 1: raise my flag at both neighbours and wait for theirs before touching a ghost plane
    HaloSync* mine;
    HaloSync* lo_sync;
    HaloSync* hi_sync;
    const unsigned long long* seq_base;  // flag value = (seq_base ? *seq_base : 0) + seq_off: the number of this flip
    unsigned long long seq_off;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Flip of the ghost sets.  One thread of the launch tells both neighbours that this rank has finished the period (its
// pushes were flushed when the previous kernel completed: the flag store needs no fence of its own).  A block that
// is going to touch a ghost plane — `waits`: its planes include a ghost plane or a boundary plane, whose t
// neighbour is one — spins with acquire loads until both neighbours have said the same, i.e. until the other ghost
// set holds their planes; every other block goes straight on.  No fence: the acquire load orders the block's later
// reads (bar.sync extends it to the other threads), and a system-scope fence per block costs microseconds.
__device__ __forceinline__ void slab_flip(const SlabArgs& sl, bool first_block, bool waits) {
    if (threadIdx.x == 0 && (first_block || waits)) {
        const unsigned long long seq = (sl.seq_base ? *sl.seq_base : 0ull) + sl.seq_off;
        if (first_block) {
            st_relaxed_sys(&sl.lo_sync->from_hi, seq);
            st_relaxed_sys(&sl.hi_sync->from_lo, seq);
        }
        if (waits) {
            halo_wait(&sl.mine->from_lo, seq, &sl.mine->error);
            halo_wait(&sl.mine->from_hi, seq, &sl.mine->error);
        }
    }
    __syncthreads();  // also where the compiler sees the block converge again: without it the sweep loop is compiled as divergent code
}

// The wait of slab_flip on its own: for a block that touches ghost planes only in its ghost tasks, after its own rows.
__device__ __forceinline__ void slab_wait(const SlabArgs& sl) {
    if (threadIdx.x == 0) {
        const unsigned long long seq = (sl.seq_base ? *sl.seq_base : 0ull) + sl.seq_off;
        halo_wait(&sl.mine->from_lo, seq, &sl.mine->error);
        halo_wait(&sl.mine->from_hi, seq, &sl.mine->error);
    }
    __syncthreads();
}

template <int XQ, bool WILSON, bool COUNT, bool SINGLE, bool SLAB, bool MEAS>
__global__ void __launch_bounds__(kN16Threads, (COUNT || WILSON) ? 3 : 4)
k_sweep_n16(const Geom g, const int colour, uint64_t sweep, const uint64_t* __restrict__ ctr, const int32_t tl_first,
            const uint32_t n_first, const int32_t tl_second, const uint32_t n_planes, const uint32_t n_chunks,
            uint16_t* __restrict__ own_base, const uint16_t* __restrict__ oth_base,
            const ChainDev* __restrict__ chains, const ChainDev chain0, const uint32_t* __restrict__ thr_all,
            unsigned long long* __restrict__ acc, const uint32_t max_thr, const SlabArgs sl,
            unsigned long long* __restrict__ obs_e, const ObsFin fin, const uint32_t rows_base, const uint32_t rows_rem) {
    constexpr uint32_t ROWQ = XQ;                        // uint4 per row of one colour
    constexpr uint32_t LPR = XQ < 32 ? XQ : 32;          // lanes of a warp on one row
    constexpr uint32_t PPW = 32u / LPR;                  // planes per warp
    constexpr int SEGS = XQ / 32;                        // 512-site row segments (0: a row is narrower than a warp)
    __shared__ unsigned int s_cnt;
    __shared__ unsigned long long s_obs;
    __shared__ bool s_last;
    pdl_launch_dependents();
    const uint32_t chain = SINGLE ? 0u : blockIdx.z;
    // MEAS: this launch is the colour-1 half-sweep of a step whose energy is due: accumulate it on the way (n16_obs_pair)
    // into obs_e.  An instantiation of its own: the plain kernel must not carry the second loop (3.5 % slower, measured).
    if (MEAS && threadIdx.x == 0) s_obs = 0;
    long long sum_e = 0;
    const ChainDev cd = SINGLE ? chain0 : chains[chain];
    if (COUNT && threadIdx.x == 0) s_cnt = 0;
    const N16Tab tab = build_interleaved(thr_all + (size_t)chain * kThrStride, max_thr, g.entry_bytes);
    // The wait for the previous launch: a slab launch that flips the ghost sets needs it before the flip; every other
    // launch takes it as late as possible, after the index arithmetic and right before the first load of the field —
    // what precedes it runs under the previous launch's drain (128^3: 3.03 -> 3.49e11, 256^3: 7.59 -> 7.73e11,
    // profiles/r2_late_wait.txt)
    const bool flips = SLAB && sl.sync != 0u;
    if (flips) {
        pdl_wait();
        if (ctr) sweep += ld_after_wait(ctr);
    }
    // planes of this block's group: local numbers [grp_lo, grp_hi) of a launch with one range (every SLAB launch)
    const int32_t grp_lo = tl_first + (int32_t)((blockIdx.y >> 1) * (2u * PPW * (blockDim.x >> 5)));
    const int32_t grp_hi = grp_lo + (int32_t)(2u * PPW * (blockDim.x >> 5));
    if (SLAB) {
        if (sl.sync)
            slab_flip(sl, blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0, grp_lo <= 0 || grp_hi >= (int32_t)g.Tl);
    }

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t seg = SEGS > 1 ? blockIdx.x % SEGS : 0u, chunk = SEGS > 1 ? blockIdx.x / SEGS : blockIdx.x;
    // all planes of a block have the parity of blockIdx.y (see above)
    const uint32_t slot = PPW * warp + lane / LPR;   // plane slot within the block
    const uint32_t pl_raw = (blockIdx.y >> 1) * (2u * PPW * (blockDim.x >> 5)) + 2u * slot + (blockIdx.y & 1u);
    // rows [y0, y1) of this y chunk: Y = n_chunks * rows_base + rows_rem, the first rows_rem chunks one row longer
    const uint32_t y0 = chunk * rows_base + min(chunk, rows_rem), y1 = y0 + rows_base + (chunk < rows_rem ? 1u : 0u);
    int accepted = 0;
    // (An early exit of warps without a live plane slot — `if (__all_sync(~0u, pl_raw >= n_planes)) return;` — compiles
    // without divergence handling but moves the Philox keys out of the uniform registers: 3.6 % slower at 256^3.)
    bool live = pl_raw < n_planes;                    // surplus plane slots: plane 0 of the range, stores off
    const uint32_t pl = live ? pl_raw : 0u;
    {
        // The plane of a thread and what follows from it are variables set by set_plane: a SLAB launch first runs its
        // ghost tasks on other planes (below); everywhere else they are assigned once.
        const uint32_t kq = SEGS ? seg * 32u + lane : lane & (LPR - 1u);  // uint4 column of this thread within the row
        const uint32_t planeq = g.Y * ROWQ;
        const uint32_t cq = chain * (uint32_t)(g.chain_stride / 8);
        uint4* const q_own = reinterpret_cast<uint4*>(own_base);
        const uint4* const q_oth = reinterpret_cast<const uint4*>(oth_base);
        const uint32_t* const w_oth = reinterpret_cast<const uint32_t*>(oth_base);
        uint32_t tl, tg, r_own, b_own, b_ta, b_tb, e_right, e_left, grp_base;
        bool owned, wil_plane;
        auto set_plane = [&](const uint32_t t) {
            tl = t;
            if (SLAB) {
                tg = (g.t0 + g.T + tl) % g.T;  // ghost planes: tl < 0 or >= Tl
            } else {
                tg = g.t0 + tl;                // own planes only
                tg = tg >= g.T ? tg - g.T : tg;
            }
            r_own = cq + plane_self(g, tl) * planeq;  // row 0 of this plane
            b_own = r_own + kq;
            b_ta = cq + plane_above(g, tl) * planeq + kq;
            b_tb = cq + plane_below(g, tl) * planeq + kq;
            // 32-bit word of the other-colour row that holds the x neighbour across the segment edge (SEGS > 1):
            // first word of the column to the right / last word of the column to the left, periodic in the row
            e_right = (r_own + (kq + 1u) % ROWQ) * 4u;
            e_left = (r_own + (kq + ROWQ - 1u) % ROWQ) * 4u + 3u;
            grp_base = tg * g.Y * (2u * ROWQ) + kq * 2u;
            wil_plane = WILSON && tg < cd.wil_h;
        };
        // PUSH: this plane's rows also go to a neighbour's other ghost set (a thin slab may owe both neighbours)
        uint4* q_lo = nullptr;
        uint4* q_hi = nullptr;
        const RoundKeys rk = make_round_keys(cd.k0, cd.k1);
        uint32_t y, p;
        uint4 wa, wb, wc;
        const uint32_t src_r = (lane & ~(LPR - 1u)) | ((lane + 1u) & (LPR - 1u)), src_l = (lane & ~(LPR - 1u)) | ((lane - 1u) & (LPR - 1u));
        const uint32_t y_one = 1u % g.Y;
        uint4 v_pre, on_pre, ta_pre, tb_pre;  // the first row's loads, issued right behind the wait (PRE)
#define LQFT_N16_STEP(OM, OC, ON, PUSH, MEAS, PRE)                                                                          \
    {                                                                                                                    \
        const uint32_t yn = (y + 1 == g.Y) ? 0 : y + 1;                                                                  \
        uint4* const po = q_own + (b_own + y * ROWQ);                                                                    \
        const uint4 v = PRE ? v_pre : *po;                                                                               \
        ON = PRE ? on_pre : __ldg(q_oth + (b_own + yn * ROWQ));                                                          \
        const uint4 ta = PRE ? ta_pre : __ldg(q_oth + (b_ta + y * ROWQ));                                                \
        const uint4 tb = PRE ? tb_pre : __ldg(q_oth + (b_tb + y * ROWQ));                                                \
        uint32_t e_ld = 0;                                                                                               \
        const bool e_lane = SEGS > 1 && (p ? lane == 31u : lane == 0u);                                                  \
        if (SEGS > 1) {                                                                                                  \
            if (e_lane) e_ld = PRE ? ld_after_wait(w_oth + ((p ? e_right : e_left) + y * (ROWQ * 4u)))                   \
                                   : __ldg(w_oth + ((p ? e_right : e_left) + y * (ROWQ * 4u)));                          \
        }                                                                                                                \
        const uint32_t grp = grp_base + y * (2u * ROWQ);                                                                 \
        const u32x4 r0 = philox4x32_rk(grp, (uint32_t)colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);             \
        const u32x4 r1 = philox4x32_rk(grp + 1u, (uint32_t)colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);        \
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
            w = n16_row_update<1, WILSON, COUNT>(v, OM, OC, ON, ta, tb, edge, r0, r1, tab, woff, acc_row, s4, s2); \
        } else {                                                                                                         \
            uint32_t edge = __shfl_sync(0xffffffffu, OC.w, src_l);                                                       \
            if (SEGS > 1) edge = e_lane ? e_ld : edge;                                                                   \
            w = n16_row_update<0, WILSON, COUNT>(v, OM, OC, ON, ta, tb, edge, r0, r1, tab, woff, acc_row, s4, s2); \
        }                                                                                                                \
        if (COUNT) accepted += owned ? acc_row : 0;                                                                      \
        if (live) *po = w;                                                                                               \
        if (MEAS) sum_e += owned ? n16_obs_row<1, 1>(w, s4, s2, OC) : 0;                                                 \
        if (PUSH) {                                                                                                      \
            if (q_lo) q_lo[y * ROWQ] = w;                                                                                \
            if (q_hi) q_hi[y * ROWQ] = w;                                                                                \
        }                                                                                                                \
        y = yn;                                                                                                          \
        p ^= 1u;                                                                                                         \
    }
        // the thread's own plane
        live = pl_raw < n_planes;
        set_plane(pl < n_first ? (uint32_t)tl_first + pl : (uint32_t)tl_second + (pl - n_first));
        owned = live && tl < g.Tl;
        if (SLAB) {
            if (owned && sl.depth) {
                if (tl < sl.depth && sl.lo)
                    q_lo = reinterpret_cast<uint4*>(sl.lo) + (chain * sl.lo_stride8 + (sl.lo_plane + tl) * planeq + kq);
                if (tl + sl.depth >= g.Tl && sl.hi)
                    q_hi = reinterpret_cast<uint4*>(sl.hi) + (chain * sl.hi_stride8 + (sl.hi_plane + (tl + sl.depth - g.Tl)) * planeq + kq);
            }
        }
        y = y0;
        // x parity of the first row, no thread-dependent term; the second range starts at the parity the first one
        // would have continued with (launch_sweep: Tl is even)
        p = (y0 + g.t0 + (uint32_t)tl_first + (blockIdx.y & 1u) + (uint32_t)colour) & 1u;
        // a warp without a live plane slot leaves after one row (the vote keeps the trip count warp uniform for the
        // compiler; an outright early exit costs the uniform registers, see above)
        uint32_t n = __all_sync(0xffffffffu, pl_raw >= n_planes) ? 1u : y1 - y0;
        if (!flips) {
            pdl_wait();
            if (ctr) sweep += ld_after_wait(ctr);
        }
        {
            // the window rows and the four loads of the first row step (in the step itself they would be scheduled behind
            // the round keys and the rest of the set-up, a hundred instructions after the wait); coherent loads: ld_after_wait
            const uint32_t ym = (y == 0) ? g.Y - 1 : y - 1, yn = (y + 1 == g.Y) ? 0 : y + 1;
            wa = ld_after_wait(q_oth + (b_own + ym * ROWQ));
            wb = ld_after_wait(q_oth + (b_own + y * ROWQ));
            v_pre = ld_after_wait(q_own + (b_own + y * ROWQ));
            on_pre = ld_after_wait(q_oth + (b_own + yn * ROWQ));
            ta_pre = ld_after_wait(q_oth + (b_ta + y * ROWQ));
            tb_pre = ld_after_wait(q_oth + (b_tb + y * ROWQ));
        }
        // only the blocks whose group holds one of the first / last `depth` owned planes carry the push stores
        if (SLAB && sl.depth != 0u && grp_hi > 0 && grp_lo < (int32_t)g.Tl &&
            (grp_lo < (int32_t)sl.depth || grp_hi > (int32_t)(g.Tl - sl.depth))) {
            LQFT_N16_STEP(wa, wb, wc, true, false, true)
            if (--n != 0)
                for (;;) {
                    LQFT_N16_STEP(wb, wc, wa, true, false, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wc, wa, wb, true, false, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wa, wb, wc, true, false, false)
                    if (--n == 0) break;
                }
        } else if (MEAS) {
            LQFT_N16_STEP(wa, wb, wc, false, true, true)
            if (--n != 0)
                for (;;) {
                    LQFT_N16_STEP(wb, wc, wa, false, true, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wc, wa, wb, false, true, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wa, wb, wc, false, true, false)
                    if (--n == 0) break;
                }
        } else {
            LQFT_N16_STEP(wa, wb, wc, false, false, true)
            if (--n != 0)
                for (;;) {
                    LQFT_N16_STEP(wb, wc, wa, false, false, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wc, wa, wb, false, false, false)
                    if (--n == 0) break;
                    LQFT_N16_STEP(wa, wb, wc, false, false, false)
                    if (--n == 0) break;
                }
        }
        // Ghost tasks (slabs).  The ghost planes this half-sweep updates redundantly (sl.extra per side) do not get
        // blocks of their own: 14 planes beside 256 owned ones made a ninth, mostly empty pair of plane groups — 576
        // blocks of 8 rows where the owned planes alone fill 592 blocks of 7 (and 130 planes of 1024^2 nine pairs
        // where 128 fill eight).  Instead their rows are dealt, one (plane slot, row) at a time, to the warps that
        // hold the same y chunk and plane parity: task k goes to warp k mod W of the W warps of this chunk and parity,
        // which runs it as a row step with a cold register window (rows y - 1 and y loaded instead of kept) after its
        // own rows.  (Before its own rows: 2 - 3 % slower, measured on 2 GPUs — every block then waits for the
        // neighbours' flags at the start of a period instead of after its rows.)
        if (SLAB) {
            if (sl.extra != 0u) {
                const bool at_edge = grp_lo <= 0 || grp_hi >= (int32_t)g.Tl;
                if (sl.sync && !at_edge) slab_wait(sl);  // (edge blocks have waited in the prologue)
                const uint32_t E = sl.extra, par = blockIdx.y & 1u;
                const uint32_t n_lo = (E + par) >> 1;                      // ghost planes of this parity below plane 0
                const uint32_t nw = blockDim.x >> 5, W = (gridDim.y >> 1) * nw;
                const uint32_t wid = (blockIdx.y >> 1) * nw + __reduce_min_sync(0xffffffffu, warp);  // uniform for the compiler
                const uint32_t rows = y1 - y0, n_tasks = ((E + PPW - 1u) / PPW) * rows;
                owned = false;
#pragma unroll 1
                for (uint32_t k = wid; k < n_tasks; k += W) {
                    const uint32_t s = k / rows, q = s * PPW + lane / LPR;   // ghost plane number q of this parity
                    live = q < E;
                    // local plane: -(2 q + 2 - par) below, Tl + 2 (q - n_lo) + par above (Tl is even)
                    set_plane(!live ? 0u : (q < n_lo ? (uint32_t)(-(int32_t)(2u * q + 2u - par)) : g.Tl + 2u * (q - n_lo) + par));
                    y = y0 + (k - s * rows);
                    p = (y + g.t0 + par + (uint32_t)colour) & 1u;
                    const uint32_t yb = (y == 0) ? g.Y - 1 : y - 1;
                    wa = __ldg(q_oth + (b_own + yb * ROWQ));
                    wb = __ldg(q_oth + (b_own + y * ROWQ));
                    LQFT_N16_STEP(wa, wb, wc, false, false, false)
                }
            }
        }
#undef LQFT_N16_STEP
    }
    if (COUNT)
        block_count_flush(&s_cnt, (unsigned int)accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
    if (MEAS) {
        // the block's share of the chain's energy sum (obs[chain][0])
        sum_e = warp_sum_ll(sum_e);
        __syncthreads();
        if ((threadIdx.x & 31u) == 0u && sum_e != 0) atomicAdd(&s_obs, (unsigned long long)sum_e);
        __syncthreads();
        // the last block of the launch to get here logs the energies of all chains into their slots and clears the
        // sums (as the observe kernel does, kernels_observe.cuh): no finalising launch after the half-sweep
        if (threadIdx.x == 0) {
            if (s_obs != 0) atomicAdd(obs_e + (size_t)chain * 2, 2ull * s_obs);
            __threadfence();
            s_last = atomicAdd(fin.ticket, 1u) + 1u == gridDim.x * gridDim.y * gridDim.z;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            obs_log_sums(fin, kObsEnergy, sweep, 0u, SINGLE ? 1u : gridDim.z, obs_e);
            if (threadIdx.x == 0) *fin.ticket = 0;
        }
    }
}

// ---- first fill of a ghost set ---------------------------------------------------------------------------------
// My first `depth` owned planes go to the lower neighbour's upper ghost planes, my last `depth` owned planes to the
// upper neighbour's lower ghost planes — of the set that is NOT current there; both colours, all chains.  No flags:
// the flip that follows (slab_flip / k_n16_flip, after this kernel has completed) tells the neighbours.
// grid = (kN16FillBlocks, 4 = side * 2 + colour, n_chains), block = 256.
struct FillArgs {
    uint16_t* mine[2];       // my narrow arrays (colour 0, 1)
    uint16_t* lo[2];         // lower neighbour's narrow arrays, as mapped here
    uint16_t* hi[2];         // upper neighbour's
    uint64_t stride_mine, stride_lo, stride_hi;  // elements per chain
    uint32_t plane;          // elements per plane per colour
    uint32_t depth;          // planes to send per side == ghost planes per side
    uint32_t tl_mine;        // owned planes here
    uint32_t lo_plane, hi_plane;  // storage plane at the lower / upper neighbour where the first sent plane goes
};

__global__ void __launch_bounds__(256)
k_n16_fill(const FillArgs a) {
    const uint32_t side = blockIdx.y >> 1, colour = blockIdx.y & 1u, chain = blockIdx.z;
    const uint64_t n = (uint64_t)a.depth * a.plane / 8;  // uint4 per (side, colour, chain)
    // storage planes here: [0, depth) lower ghosts, [depth, depth + Tl) owned
    const uint16_t* src = a.mine[colour] + (size_t)chain * a.stride_mine + (size_t)(side ? a.tl_mine : a.depth) * a.plane;
    uint16_t* dst = side ? a.hi[colour] + (size_t)chain * a.stride_hi + (size_t)a.hi_plane * a.plane
                         : a.lo[colour] + (size_t)chain * a.stride_lo + (size_t)a.lo_plane * a.plane;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = __ldcg(reinterpret_cast<const uint4*>(src) + i);
}

// The flip on its own (one thread), for operations other than a half-sweep that need the next ghost set.
__global__ void k_n16_flip(const SlabArgs sl) { slab_flip(sl, true, true); }

// ---- i32 <-> narrow ---------------------------------------------------------------------------------------
// Largest |height - centre| a conversion met, into *worst: one atomic per block, and none once the word is at
// least as large (capped at 2^24).
__device__ __forceinline__ void n16_report_worst(uint32_t far, uint32_t* __restrict__ worst) {
    __shared__ uint32_t s_far;
    if (threadIdx.x == 0) s_far = 0;
    __syncthreads();
    far = __reduce_max_sync(0xffffffffu, far);
    if ((threadIdx.x & 31u) == 0u && far) atomicMax(&s_far, far);
    __syncthreads();
    if (threadIdx.x == 0 && s_far > *reinterpret_cast<volatile uint32_t*>(worst)) atomicMax(worst, s_far);
}

// Owned planes of one colour, 4 heights per thread.  grid = (ceil(n/4/256), n_chains); n = Tl * plane elements per
// chain, src/dst point at the first owned plane of chain 0.  worst = max |height - centre| over everything converted.
__global__ void __launch_bounds__(256)
k_narrow(const int32_t* __restrict__ f, uint64_t f_stride, uint16_t* __restrict__ n, uint64_t n_stride, uint64_t count,
         const int32_t* __restrict__ centre, uint32_t* __restrict__ worst) {
    const uint32_t chain = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t far = 0;
    if (i < count / 4) {
        const int c = centre[chain];
        const int4 v = reinterpret_cast<const int4*>(f + (size_t)chain * f_stride)[i];
        // wide arithmetic: an uploaded field may hold any i32
        const long long dx = (long long)v.x - c, dy = (long long)v.y - c, dz = (long long)v.z - c, dw = (long long)v.w - c;
        const long long m = max(max(llabs(dx), llabs(dy)), max(llabs(dz), llabs(dw)));
        far = (uint32_t)min(m, (long long)(1 << 24));
        uint2 o;
        o.x = ((uint32_t)(dx + kN16Bias) & 0xffffu) | ((uint32_t)(dy + kN16Bias) << 16);
        o.y = ((uint32_t)(dz + kN16Bias) & 0xffffu) | ((uint32_t)(dw + kN16Bias) << 16);
        reinterpret_cast<uint2*>(n + (size_t)chain * n_stride)[i] = o;
    }
    n16_report_worst(far, worst);
}

__global__ void __launch_bounds__(256)
k_widen(const uint16_t* __restrict__ n, uint64_t n_stride, int32_t* __restrict__ f, uint64_t f_stride, uint64_t count,
        const int32_t* __restrict__ centre) {
    const uint32_t chain = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count / 4) return;
    const int c = centre[chain] - kN16Bias;
    const uint2 v = reinterpret_cast<const uint2*>(n + (size_t)chain * n_stride)[i];
    reinterpret_cast<int4*>(f + (size_t)chain * f_stride)[i] =
        make_int4((int)(v.x & 0xffffu) + c, (int)(v.x >> 16) + c, (int)(v.y & 0xffffu) + c, (int)(v.y >> 16) + c);
}

// centre[chain] := offset[chain] (the chain's pending normalisation shift: identical on every rank, and the
// stored height of a site of the field whenever a normalize has run); *worst := 0.
__global__ void k_narrow_begin(const int32_t* __restrict__ offset, int32_t* __restrict__ centre, uint32_t n_chains,
                               uint32_t* __restrict__ worst) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chains) centre[c] = offset[c];
    if (c == 0) *worst = 0;
}

// ---- the 16-bit copy as the storage between calls: re-centring -----------------------------------------------------
// A sweep moves a height by at most one, so after a conversion that measured `worst` the copy is good for
// kN16Bias - 1 - worst sweeps (host: narrow_budget).  When they are used up the heights are measured against the
// chain's CURRENT normalisation offset (k_n16_worst) and, if they lie within kN16Guard of it, moved there in place
// (k_n16_rebias; centre := offset afterwards, k_narrow_begin); otherwise the field goes back to i32 (k_widen).
// Same grid as k_narrow.  centre = current centres, target = the offsets.
__global__ void __launch_bounds__(256)
k_n16_worst(const uint16_t* __restrict__ n, uint64_t n_stride, uint64_t count, const int32_t* __restrict__ centre,
            const int32_t* __restrict__ target, uint32_t* __restrict__ worst) {
    const uint32_t chain = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t far = 0;
    if (i < count / 4) {
        const long long d = (long long)centre[chain] - target[chain] - kN16Bias;
        const uint2 v = reinterpret_cast<const uint2*>(n + (size_t)chain * n_stride)[i];
        const long long a = llabs((long long)(v.x & 0xffffu) + d), b = llabs((long long)(v.x >> 16) + d);
        const long long c = llabs((long long)(v.y & 0xffffu) + d), e = llabs((long long)(v.y >> 16) + d);
        far = (uint32_t)min(max(max(a, b), max(c, e)), (long long)(1 << 24));
    }
    n16_report_worst(far, worst);
}

__global__ void __launch_bounds__(256)
k_n16_rebias(uint16_t* __restrict__ n, uint64_t n_stride, uint64_t count, const int32_t* __restrict__ centre,
             const int32_t* __restrict__ target) {
    const uint32_t chain = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count / 4) return;
    // |delta| <= 2 * kN16Guard + 2 (k_n16_worst has passed): the packed halves cannot carry into each other
    const uint32_t d = (uint32_t)(centre[chain] - target[chain]) & 0xffffu;
    uint2* q = reinterpret_cast<uint2*>(n + (size_t)chain * n_stride) + i;
    uint2 v = *q;
    v.x = ((v.x + d) & 0xffffu) | ((v.x + (d << 16)) & 0xffff0000u);
    v.y = ((v.y + d) & 0xffffu) | ((v.y + (d << 16)) & 0xffff0000u);
    *q = v;
}

// Every height of the 16-bit arrays := kN16Bias (a field of zeros around centre 0).  n = uint32 words.
__global__ void k_n16_zero(uint32_t* __restrict__ words, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        words[i] = (uint32_t)kN16Bias * 0x10001u;
}

// ---- host side ----------------------------------------------------------------------------------------------
struct NarrowPlan {
    bool enabled = false;
    int xq = 0;            // X / 16: uint4 per row of one colour (1 ... 64)
    int sm_count = 0;
    uint32_t forced_chunks = 0;  // LQFT_N16_CHUNKS (tests): y chunks per plane group instead of the cost model's choice
    int occ[2][2][2][kN16Warps + 1];  // resident blocks per SM by (WILSON, COUNT, SLAB, warps per block); 0 = not asked yet
    int segs() const { return xq / 32; }                       // 512-site row segments
    uint32_t ppw() const { return xq < 32 ? 32u / xq : 1u; }   // planes per warp
};

struct NarrowArgs {
    Geom g;               // geometry of the narrow arrays (ghost = their ghost depth)
    int colour;
    uint64_t sweep;
    const uint64_t* ctr;
    int32_t tl_first;
    uint32_t n_first;     // planes of the first range (== n_planes for one range)
    int32_t tl_second;
    uint32_t n_planes;
    uint32_t n_chains;
    uint16_t* own;
    const uint16_t* oth;
    const ChainDev* chains;
    ChainDev chain0;      // host copy of chain 0's parameters (used when n_chains == 1)
    const uint32_t* thr;
    unsigned long long* acc;
    uint32_t max_thr;
    bool wilson, count;
    cudaStream_t stream;
    bool pdl;
    bool slab = false;     // SLAB instantiation: push boundary planes and / or flip the ghost sets (sl)
    SlabArgs sl{};
    unsigned long long* obs_e = nullptr;  // colour-1 half-sweep of a measured step: energy sums accumulate here ([chain][2])
    ObsFin fin{};                         //   ... and are logged by the launch's last block as `fin` says
};

template <int XQ, bool SLAB>
inline const void* n16_variant_s(bool w, bool c, bool s) {
    if (w) {
        if (c) return s ? (const void*)k_sweep_n16<XQ, true, true, true, SLAB, false> : (const void*)k_sweep_n16<XQ, true, true, false, SLAB, false>;
        return s ? (const void*)k_sweep_n16<XQ, true, false, true, SLAB, false> : (const void*)k_sweep_n16<XQ, true, false, false, SLAB, false>;
    }
    if (c) return s ? (const void*)k_sweep_n16<XQ, false, true, true, SLAB, false> : (const void*)k_sweep_n16<XQ, false, true, false, SLAB, false>;
    return s ? (const void*)k_sweep_n16<XQ, false, false, true, SLAB, false> : (const void*)k_sweep_n16<XQ, false, false, false, SLAB, false>;
}
// measuring instantiations: one GPU, no acceptance counts
template <int XQ>
inline const void* n16_variant_m(bool w, bool s) {
    if (w) return s ? (const void*)k_sweep_n16<XQ, true, false, true, false, true> : (const void*)k_sweep_n16<XQ, true, false, false, false, true>;
    return s ? (const void*)k_sweep_n16<XQ, false, false, true, false, true> : (const void*)k_sweep_n16<XQ, false, false, false, false, true>;
}
template <int XQ>
inline const void* n16_variant(bool w, bool c, bool s, bool slab, bool meas) {
    if (meas) return n16_variant_m<XQ>(w, s);
    return slab ? n16_variant_s<XQ, true>(w, c, s) : n16_variant_s<XQ, false>(w, c, s);
}
inline const void* n16_kernel(int xq, bool w, bool c, bool s, bool slab, bool meas = false) {
    switch (xq) {
        case 1: return n16_variant<1>(w, c, s, slab, meas);
        case 2: return n16_variant<2>(w, c, s, slab, meas);
        case 4: return n16_variant<4>(w, c, s, slab, meas);
        case 8: return n16_variant<8>(w, c, s, slab, meas);
        case 16: return n16_variant<16>(w, c, s, slab, meas);
        case 32: return n16_variant<32>(w, c, s, slab, meas);
        case 64: return n16_variant<64>(w, c, s, slab, meas);
    }
    return nullptr;
}

inline bool narrow_geometry_ok(const Geom& g, uint32_t max_thr, uint32_t n_chains) {
    return g.X >= 16 && g.X <= 1024 && (g.X & (g.X - 1)) == 0 && g.Tl >= 2 && g.Tl % 2 == 0 &&
           max_thr <= (uint32_t)kN16KcMax + 1 &&
           (uint64_t)g.plane / 8 * (g.Tl + 4 * kN16GhostDepth) * n_chains < (1ull << 30) &&  // uint4 and word offsets in 32 bits
           (uint64_t)g.T * g.Y * (g.X / 8) < (1ull << 32);
}

inline cudaError_t narrow_prepare(NarrowPlan& p, const Geom& g, uint32_t n_chains, int sm_count, uint32_t max_thr, bool force_off) {
    p.enabled = false;
    if (force_off || !narrow_geometry_ok(g, max_thr, n_chains)) return cudaSuccess;
    p.sm_count = sm_count;
    p.xq = (int)(g.X / 16);
    memset(p.occ, 0, sizeof p.occ);
    p.enabled = true;
    p.forced_chunks = 0;
    if (const char* e = getenv("LQFT_N16_CHUNKS")) p.forced_chunks = (uint32_t)atoi(e);
    return cudaSuccess;
}

// Blocks along the planes come in pairs (even / odd planes): as few pairs as kN16Warps warps per block allow, then as
// few warps per block as those pairs need.  (Choosing the block size that leaves the fewest surplus warps — 17 pairs
// of 4-warp blocks for the 270 planes of a slab with ghost planes in flight instead of 9 pairs of 8 — executes 5 %
// fewer instructions and still runs 2 % slower: measured.  So does compiling the slab launches for 9-warp blocks — 270
// planes in 8 pairs, 592 blocks of 7 rows instead of 576 of 8 — at the 56 registers that keeps 4 blocks per SM: 5 %
// slower on 2 GPUs, the spills cost more than the eighth row.)
inline uint32_t narrow_plane_warps(const NarrowPlan& p, uint32_t n_planes) {  // warps one parity needs
    const uint32_t w = ((n_planes + 1) / 2 + p.ppw() - 1) / p.ppw();
    return w < 1 ? 1 : w;
}
inline uint32_t narrow_pairs(const NarrowPlan& p, uint32_t n_planes) {
    return (narrow_plane_warps(p, n_planes) + kN16Warps - 1) / kN16Warps;
}
inline uint32_t narrow_warps(const NarrowPlan& p, uint32_t n_planes) {
    const uint32_t pairs = narrow_pairs(p, n_planes);
    return (narrow_plane_warps(p, n_planes) + pairs - 1) / pairs;
}

// Resident blocks per SM of the instantiation a launch will use (each has its own register count).
inline int narrow_occupancy(NarrowPlan& p, bool wilson, bool count, bool slab, uint32_t warps) {
    int& o = p.occ[wilson][count][slab][warps];
    if (o == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, n16_kernel(p.xq, wilson, count, true, slab), (int)(32 * warps), 0) != cudaSuccess) {
            cudaGetLastError();
            n = 1;
        }
        o = n > 0 ? n : 1;
    }
    return o;
}

// grid.y
inline uint32_t narrow_groups(const NarrowPlan& p, uint32_t n_planes) { return 2 * narrow_pairs(p, n_planes); }

// y chunks per plane group: one wave if possible, chunks as short as that allows (cost model of stream_chunks;
// the per-chunk prologue is worth about two rows here).
inline uint32_t narrow_chunks(NarrowPlan& p, const Geom& g, uint32_t n_planes, uint32_t n_chains, bool wilson = false,
                              bool count = false, bool slab = false) {
    const uint64_t slots = (uint64_t)p.sm_count * narrow_occupancy(p, wilson, count, slab, narrow_warps(p, n_planes));
    const uint64_t base = narrow_groups(p, n_planes) * (uint64_t)(p.segs() ? p.segs() : 1) * n_chains;  // blocks per chunk
    // down to one row per chunk: a small lattice (128^3: 4 096 warp-rows per half-sweep) is short of warps, not of
    // prologue instructions — 128 chunks of one row run 25 % faster there than 64 chunks of two (tools/probe_chunks.py)
    const uint32_t max_chunks = g.Y;
    if (p.forced_chunks) return p.forced_chunks < g.Y ? p.forced_chunks : g.Y;
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t nc = 1; nc <= max_chunks; ++nc) {
        const uint64_t blocks = base * nc;
        const uint64_t waves = (blocks + slots - 1) / slots;
        // an SM is throughput bound over its resident warps rather than waiting for the slowest one: chunks of 7 and 8
        // rows cost about their mean, the longest one counts only half
        const double rows = 0.5 * ((double)g.Y / nc + (double)((g.Y + nc - 1) / nc));
        const double cost = (double)waves * (rows + 2.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = nc;
        }
    }
    return best;
}

inline cudaError_t narrow_half(NarrowPlan& p, const NarrowArgs& a) {
    const uint32_t n_groups = narrow_groups(p, a.n_planes);
    uint32_t n_chunks = narrow_chunks(p, a.g, a.n_planes, a.n_chains, a.wilson, a.count, a.slab);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_chunks * (p.segs() ? p.segs() : 1), n_groups, a.n_chains);
    cfg.blockDim = dim3(32 * narrow_warps(p, a.n_planes));
    cfg.dynamicSmemBytes = 0;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
    // parameter order of k_sweep_n16
    Geom g = a.g;
    int colour = a.colour;
    uint64_t sweep = a.sweep;
    const uint64_t* ctr = a.ctr;
    int32_t tl_first = a.tl_first, tl_second = a.tl_second;
    uint32_t n_first = a.n_first, n_planes = a.n_planes, max_thr = a.max_thr;
    uint16_t* own = a.own;
    const uint16_t* oth = a.oth;
    const ChainDev* chains = a.chains;
    ChainDev chain0 = a.chain0;
    const uint32_t* thr = a.thr;
    unsigned long long* acc = a.acc;
    SlabArgs sl = a.sl;
    unsigned long long* obs_e = a.obs_e;
    ObsFin fin = a.fin;
    uint32_t rows_base = a.g.Y / n_chunks, rows_rem = a.g.Y % n_chunks;
    void* args[] = {&g, &colour, &sweep, &ctr, &tl_first, &n_first, &tl_second, &n_planes, &n_chunks, &own, &oth, &chains,
                    &chain0, &thr, &acc, &max_thr, &sl, &obs_e, &fin, &rows_base, &rows_rem};
    return cudaLaunchKernelExC(&cfg, n16_kernel(p.xq, a.wilson, a.count, a.n_chains == 1, a.slab, a.obs_e != nullptr), args);
}

}  // namespace lqft

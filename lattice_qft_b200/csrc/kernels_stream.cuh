// kernels_stream.cuh — the streaming half-sweep for wide rows (X = 256 * SEGS).
//
// The marching kernel (kernels_tiled.cuh) cut the instruction count 3x and left the sweep latency
// bound: at 256^3 the two colour arrays (64 MiB) do not stay in the two-die L2, one launch still
// reads ~34 MB from HBM, and a warp sits on LDG round trips (profiles/r1_march_*: long scoreboard
// 7-8 of every issue slot, issue 36-50 % busy, 2 waves of 8-step blocks).  This kernel keeps the
// register window and the shuffle trick and adds
//   * a per-thread asynchronous prefetch ring in shared memory: the four row loads of step s + D
//     (own row, other colour row y+1, planes t-1 / t+1) are issued with cp.async (LDGSTS) D steps
//     ahead and read back with LDS.128 — depth without registers, no block barriers;
//   * one balanced wave: the y range of every plane group is cut into chunks so that the grid just
//     fills the resident-block slots of the 148 SMs, and each thread streams its whole chunk with a
//     runtime loop (prologue paid once per chunk, no tail wave);
//   * rows wider than one warp (X = 512, 1024): a row is SEGS warp-wide segments; the neighbour
//     across a segment edge is prefetched by the edge lanes into the same ring.
// Same sites, same Philox counters, same decisions as every other path: bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_tiled.cuh"

namespace lqft {

#ifndef LQFT_STREAM_WARPS
#define LQFT_STREAM_WARPS 4
#endif
#ifndef LQFT_STREAM_DEPTH
#define LQFT_STREAM_DEPTH 2
#endif
#ifndef LQFT_STREAM_MINBLOCKS
#define LQFT_STREAM_MINBLOCKS 7
#endif
#ifndef LQFT_STREAM_HALO_MINBLOCKS
#define LQFT_STREAM_HALO_MINBLOCKS 6
#endif
constexpr int kStreamWarps = LQFT_STREAM_WARPS;
constexpr int kStreamDepth = LQFT_STREAM_DEPTH;   // prefetch distance in rows (even)
constexpr int kStreamThreads = kStreamWarps * 32;
static_assert(kStreamDepth % 2 == 0 && kStreamDepth >= 2, "the ring depth must be even");

__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4_ca(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---- peer-memory halo (world_size > 1): the sweep kernel pushes its boundary rows straight into the
// neighbours' ghost planes over NVLink and publishes a phase flag when its boundary blocks are done;
// the warps that read a ghost plane wait for the neighbour's flag of the previous phase.  No separate
// pack / send / receive launches and no collective on the critical path.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag >= need.  Bounded (g_halo_timeout_ns, 30 s unless LQFT_HALO_TIMEOUT_S says otherwise: ranks are
// separate processes and may be seconds apart — module load, graph instantiation, a rank writing CSVs): on
// expiry the error word is set and the kernel traps.  A lost neighbour therefore never yields a result computed
// from stale ghost planes: the launch fails, every later call on the handle returns LQFT_ERR_CUDA, and
// lqft_synchronize / lqft_download report the halo timeout.  The handle has to be destroyed after that.
__device__ unsigned long long g_halo_timeout_ns = 30000000000ull;
__device__ __forceinline__ void halo_wait(const unsigned long long* flag, unsigned long long need, unsigned int* error) {
    if (ld_acquire_sys(flag) >= need) return;
    const unsigned long long t0 = global_timer_ns();
    const unsigned long long limit = *reinterpret_cast<volatile unsigned long long*>(&g_halo_timeout_ns);
    while (ld_acquire_sys(flag) < need) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > limit) {
            atomicExch(error, 1u);
            __threadfence_system();
            __trap();
        }
    }
}

// One thread: wait for both neighbours' flags (before an operation that reads or rewrites ghost planes).
__global__ void k_halo_wait(HaloSync* mine, unsigned long long need) {
    halo_wait(&mine->from_lo, need, &mine->error);
    halo_wait(&mine->from_hi, need, &mine->error);
}
// One thread: publish `phase` to both neighbours (after such an operation).
__global__ void k_halo_signal(HaloSync* lo_sync, HaloSync* hi_sync, unsigned long long phase) {
    st_release_sys(&lo_sync->from_hi, phase);
    st_release_sys(&hi_sync->from_lo, phase);
}

// Everything a thread needs to issue the loads of one row.
struct StreamPtrs {
    const int4* own;   // own colour, this plane, row 0 of the chunk's plane (quad column kq)
    const int4* oc;    // other colour, this plane
    const int4* ta;    // other colour, plane t+1
    const int4* tb;    // other colour, plane t-1
    const int32_t* edge;  // other colour, this plane, element across the segment edge (row 0)
};

template <int SEGS, bool P>
__device__ __forceinline__ void stream_issue(const StreamPtrs& q, uint32_t y, uint32_t Y, uint32_t ring, uint32_t edge_ring,
                                             uint32_t stage, bool edge_lane) {
    constexpr uint32_t G = 32u * SEGS;
    constexpr uint32_t STAGE_BYTES = 4u * kStreamThreads * 16u;
    const uint32_t dst = ring + stage * STAGE_BYTES;
    const uint32_t yn = (y + 1 == Y) ? 0u : y + 1;
    cp_async16_cg(dst, q.own + y * G);
    cp_async16_ca(dst + kStreamThreads * 16u, q.oc + yn * G);
    cp_async16_ca(dst + 2u * kStreamThreads * 16u, q.ta + y * G);
    cp_async16_ca(dst + 3u * kStreamThreads * 16u, q.tb + y * G);
    if (SEGS > 1) {
        if (edge_lane) cp_async4_ca(edge_ring + stage * (kStreamThreads * 4u), q.edge + y * (4u * G));
    }
}

template <int SEGS, bool WILSON, bool COUNT, bool HALO, bool P>
__device__ __forceinline__ void stream_step(int4* __restrict__ own_row, const bool push, const long long* s_delta, uint32_t ring, uint32_t edge_ring, uint32_t stage,
                                            int4& om, int4& oc, const uint32_t src_lane, const bool edge_lane,
                                            const uint32_t grp, const uint32_t colour, const uint64_t sweep,
                                            const RoundKeys& rk, const uint2* __restrict__ s_dn,
                                            const uint2* __restrict__ s_up, const int kc, const uint32_t y,
                                            const uint32_t Y, const uint32_t kq, const bool wil_plane,
                                            const uint32_t wil_w, int& accepted) {
    constexpr uint32_t STAGE_BYTES = 4u * kStreamThreads * 16u;
    const uint32_t src = ring + stage * STAGE_BYTES;
    const int4 v = lds128(src);
    const int4 on = lds128(src + kStreamThreads * 16u);
    const int4 ta = lds128(src + 2u * kStreamThreads * 16u);
    const int4 tb = lds128(src + 3u * kStreamThreads * 16u);
    // x neighbour outside the quad: P -> element k0+4 (next lane's .x), else k0-1 (previous lane's .w)
    int e = __shfl_sync(0xffffffffu, P ? oc.x : oc.w, src_lane);
    if (SEGS > 1) {
        if (edge_lane) e = lds32(edge_ring + stage * (kStreamThreads * 4u));
    }
    int4 ns;
    ns.x = oc.x + (P ? oc.y : e);
    ns.y = oc.y + (P ? oc.z : oc.x);
    ns.z = oc.z + (P ? oc.w : oc.y);
    ns.w = oc.w + (P ? e : oc.z);
    ns.x += om.x + on.x; ns.y += om.y + on.y; ns.z += om.z + on.z; ns.w += om.w + on.w;
    ns.x += ta.x + tb.x; ns.y += ta.y + tb.y; ns.z += ta.z + tb.z; ns.w += ta.w + tb.w;
    const u32x4 r = philox4x32_rk(grp, colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);
    int4 off = make_int4(0, 0, 0, 0);
    if (WILSON) {
        if (wil_plane && (y == 0u || y == 1u % Y)) {
            const int sgn = (y == 0u ? 1 : 0) - (y == 1u % Y ? 1 : 0);
            const uint32_t x0 = 8u * kq + (P ? 1u : 0u);
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
    *own_row = w;
    if (HALO) {
        // boundary plane: the same row goes to the neighbour's ghost plane over NVLink.  The byte offset from
        // the own row to the peer row sits in shared memory so that it costs no register in the hot loop.
        if (push) *reinterpret_cast<int4*>(reinterpret_cast<char*>(own_row) + *s_delta) = w;
    }
    om = oc;
    oc = on;
}

// The chunk [y_begin, y_end) of one plane, parity P0 at y_begin.
template <int SEGS, bool WILSON, bool COUNT, bool HALO, bool P0>
__device__ __forceinline__ void stream_chunk(int4* __restrict__ own_w, const bool push, const long long* s_delta, const StreamPtrs& q,
                                             const uint32_t y_begin,
                                             const uint32_t y_end, const uint32_t Y, const uint32_t ring,
                                             const uint32_t edge_ring, const uint32_t lane, const uint32_t kq,
                                             const uint32_t grp_row0, const uint32_t colour, const uint64_t sweep,
                                             const RoundKeys& rk, const uint2* __restrict__ s_dn,
                                             const uint2* __restrict__ s_up, const int kc, const bool wil_plane,
                                             const uint32_t wil_w, int& accepted) {
    constexpr uint32_t G = 32u * SEGS;
    constexpr int D = kStreamDepth;
    const uint32_t n_steps = y_end - y_begin;  // even
    const uint32_t src_up = (lane + 1u) & 31u, src_dn = (lane - 1u) & 31u;
    // edge lanes of a multi-segment row fetch their neighbour from the adjacent segment
    const bool edge_p = SEGS > 1 && lane == 31u, edge_n = SEGS > 1 && lane == 0u;
    StreamPtrs qp = q, qn = q;  // parity-true rows look right (+4), parity-false rows look left (-1)
    if (SEGS > 1) {
        const uint32_t k_right = (kq + 1u == G) ? 0u : 4u * (kq + 1u);
        const uint32_t k_left = (kq == 0u) ? 4u * G - 1u : 4u * kq - 1u;
        qp.edge = q.edge + k_right;
        qn.edge = q.edge + k_left;
    }
#pragma unroll
    for (int s = 0; s < D; ++s) {
        if ((uint32_t)s < n_steps) {
            if ((s & 1) ? !P0 : P0) stream_issue<SEGS, true>(qp, y_begin + s, Y, ring, edge_ring, (uint32_t)s, edge_p);
            else stream_issue<SEGS, false>(qn, y_begin + s, Y, ring, edge_ring, (uint32_t)s, edge_n);
        }
        cp_async_commit();
    }
    const uint32_t ym = (y_begin == 0u) ? Y - 1u : y_begin - 1u;
    int4 om = __ldg(q.oc + ym * G);
    int4 oc = __ldg(q.oc + y_begin * G);
    // Running pointers to the row that is D steps ahead (the one being prefetched); the row being updated is
    // D rows behind pf_own, a compile-time offset.  No per-step index arithmetic.
    const int4* pf_own = q.own + (y_begin + D) * G;
    const int4* pf_ta = q.ta + (y_begin + D) * G;
    const int4* pf_tb = q.tb + (y_begin + D) * G;
    const int4* pf_on = q.oc + (y_begin + D + 1u) * G;  // other colour, one row further (wraps to row 0)
    const int32_t* pf_ep = qp.edge + (y_begin + D) * (4u * G);
    const int32_t* pf_en = qn.edge + (y_begin + D) * (4u * G);
    constexpr uint32_t STAGE_BYTES = 4u * kStreamThreads * 16u;
    constexpr uint32_t SLOT = kStreamThreads * 16u;
    uint32_t grp = grp_row0 + y_begin * G;
    uint32_t y = y_begin;
    for (uint32_t s = 0; s < n_steps; s += 2, y += 2) {
        const uint32_t st0 = s & (uint32_t)(D - 1), st1 = (s + 1u) & (uint32_t)(D - 1);
        cp_async_wait<D - 1>();
        stream_step<SEGS, WILSON, COUNT, HALO, P0>(const_cast<int4*>(pf_own) - D * (int)G, push, s_delta, ring, edge_ring, st0,
                                                   om, oc, P0 ? src_up : src_dn, P0 ? edge_p : edge_n, grp, colour, sweep,
                                                   rk, s_dn, s_up, kc, y, Y, kq, wil_plane, wil_w, accepted);
        if (s + D < n_steps) {
            const uint32_t dst = ring + st0 * STAGE_BYTES;
            cp_async16_cg(dst, pf_own);
            cp_async16_ca(dst + SLOT, (y + D + 1u == Y) ? q.oc : pf_on);
            cp_async16_ca(dst + 2u * SLOT, pf_ta);
            cp_async16_ca(dst + 3u * SLOT, pf_tb);
            if (SEGS > 1) {
                if (P0 ? edge_p : edge_n) cp_async4_ca(edge_ring + st0 * (kStreamThreads * 4u), P0 ? pf_ep : pf_en);
            }
        }
        cp_async_commit();
        cp_async_wait<D - 1>();
        stream_step<SEGS, WILSON, COUNT, HALO, !P0>(const_cast<int4*>(pf_own) - (D - 1) * (int)G, push, s_delta, ring, edge_ring,
                                                    st1, om, oc, P0 ? src_dn : src_up, P0 ? edge_n : edge_p, grp + G, colour,
                                                    sweep, rk, s_dn, s_up, kc, y + 1u, Y, kq, wil_plane, wil_w, accepted);
        if (s + 1u + D < n_steps) {
            const uint32_t dst = ring + st1 * STAGE_BYTES;
            cp_async16_cg(dst, pf_own + G);
            cp_async16_ca(dst + SLOT, (y + D + 2u == Y) ? q.oc : pf_on + G);
            cp_async16_ca(dst + 2u * SLOT, pf_ta + G);
            cp_async16_ca(dst + 3u * SLOT, pf_tb + G);
            if (SEGS > 1) {
                if (P0 ? edge_n : edge_p)
                    cp_async4_ca(edge_ring + st1 * (kStreamThreads * 4u), (P0 ? pf_en : pf_ep) + 4u * G);
            }
        }
        cp_async_commit();
        pf_own += 2u * G;
        pf_on += 2u * G;
        pf_ta += 2u * G;
        pf_tb += 2u * G;
        if (SEGS > 1) {
            pf_ep += 8u * G;
            pf_en += 8u * G;
        }
        grp += 2u * G;
    }
    cp_async_wait<0>();
}

// grid = (n_chunks * SEGS, ceil(n_planes / kStreamWarps), n_chains), block = kStreamThreads.
template <int SEGS, bool WILSON, bool COUNT, bool HALO>
__global__ void __launch_bounds__(kStreamThreads, HALO ? LQFT_STREAM_HALO_MINBLOCKS : LQFT_STREAM_MINBLOCKS)
k_sweep_stream(const Geom g, const int colour, uint64_t sweep, const uint64_t* __restrict__ ctr,
               const uint32_t tl_first, const uint32_t n_planes, const uint32_t n_chunks, const uint32_t n_bchunks,
               int32_t* __restrict__ own_base, const int32_t* __restrict__ oth_base,
               const ChainDev* __restrict__ chains, const uint32_t* __restrict__ thr_all,
               unsigned long long* __restrict__ acc, const uint32_t max_thr, const HaloP2P halo) {
    constexpr uint32_t G = 32u * SEGS;
    constexpr uint32_t RING_BYTES = (uint32_t)kStreamDepth * 4u * kStreamThreads * 16u;
    constexpr uint32_t EDGE_BYTES = SEGS > 1 ? (uint32_t)kStreamDepth * kStreamThreads * 4u : 0u;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned int s_cnt;
    __shared__ long long s_delta[kStreamWarps];
    pdl_launch_dependents();
    const uint32_t chain = blockIdx.z;
    // Block prologue.  The chain parameters and the threshold table are read independently of each other (the
    // table is built for the handle's longest table; shorter chains' tails are zero, exactly what their clamp
    // would select) and BEFORE the dependency wait, so with programmatic dependent launch this part overlaps
    // the predecessor's tail.
    const ChainDev cd = chains[chain];
    if (COUNT && threadIdx.x == 0) s_cnt = 0;
    uint2* s_tab2 = reinterpret_cast<uint2*>(s_raw + RING_BYTES + EDGE_BYTES);
    const int kc = build_two_sided(s_tab2, thr_all + (size_t)chain * kThrStride, max_thr);
    const uint2* s_dn = s_tab2 + kc;
    const uint2* s_up = s_tab2 + 3 * kc + 1;
    pdl_wait();  // from here on the predecessor has completed
    if (ctr) sweep += *ctr;

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t seg = blockIdx.x % SEGS;
    uint32_t chunk = blockIdx.x / SEGS, group = blockIdx.y, chunks_here = n_chunks;
    bool boundary_block = false;
    if (HALO) {
        // peer-memory halo launch, grid.y == 1: the first n_bchunks blocks take plane group 0 — rotated so
        // that it holds the slab's two boundary planes {Tl-2, Tl-1, 0, 1} — in short chunks: they are
        // scheduled first and finish early, so pushes and flag are out long before the phase ends.
        if (chunk < n_bchunks) {
            group = 0;
            chunks_here = n_bchunks;
            boundary_block = true;
        } else {
            chunk -= n_bchunks;
            group = 1u + chunk / n_chunks;
            chunk %= n_chunks;
        }
    }
    uint32_t pl = group * kStreamWarps + warp;
    const bool plane_ok = pl < n_planes;
    if (HALO) pl = (pl + n_planes - (uint32_t)(kStreamWarps / 2) % n_planes) % n_planes;
    const uint32_t half = g.Y >> 1;
    const uint32_t y_begin = 2u * (uint32_t)(((uint64_t)chunk * half) / chunks_here);
    const uint32_t y_end = 2u * (uint32_t)(((uint64_t)(chunk + 1u) * half) / chunks_here);
    int accepted = 0;
    if (plane_ok && y_begin < y_end) {  // warp uniform
        const uint32_t tl = tl_first + pl;
        const uint32_t tg = g.t0 + tl;
        const size_t cbase = (size_t)chain * g.chain_stride;
        const uint32_t kq = seg * 32u + lane;
        const uint32_t planeq = g.plane / 4;
        const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_raw) + threadIdx.x * 16u;
        const uint32_t edge_ring = (uint32_t)__cvta_generic_to_shared(s_raw) + RING_BYTES + threadIdx.x * 4u;
        int4* own_w = reinterpret_cast<int4*>(own_base + cbase) + (plane_self(g, tl) * planeq + kq);
        const int4* othq = reinterpret_cast<const int4*>(oth_base + cbase);
        StreamPtrs q;
        q.own = own_w;
        q.oc = othq + (plane_self(g, tl) * planeq + kq);
        q.ta = othq + (plane_above(g, tl) * planeq + kq);
        q.tb = othq + (plane_below(g, tl) * planeq + kq);
        q.edge = oth_base + cbase + (size_t)plane_self(g, tl) * g.plane;
        const bool wil_plane = WILSON && tg < cd.wil_h;
        const RoundKeys rk = make_round_keys(cd.k0, cd.k1);
        const uint32_t grp_row0 = tg * g.Y * G + kq;  // Philox counter word 0 of this quad column at y = 0
        bool push = false;
        if (HALO && halo.enabled) {  // warp uniform
            const int4* peer_w = nullptr;
            if (tl == 0u) {
                if (lane == 0u) halo_wait(&halo.mine->from_lo, halo.wait_for, &halo.mine->error);
                peer_w = reinterpret_cast<const int4*>(halo.lo_ghost + cbase) + kq;
            } else if (tl + 1u == g.Tl) {
                if (lane == 0u) halo_wait(&halo.mine->from_hi, halo.wait_for, &halo.mine->error);
                peer_w = reinterpret_cast<const int4*>(halo.hi_ghost + cbase) + kq;
            }
            push = peer_w != nullptr;
            // lanes differ only by kq on both sides, so one offset per warp serves all of them
            if (push && lane == 0u)
                s_delta[warp] = reinterpret_cast<const char*>(peer_w) - reinterpret_cast<const char*>(own_w);
            __syncwarp();
        }
        if ((y_begin + tg + (uint32_t)colour) & 1u)
            stream_chunk<SEGS, WILSON, COUNT, HALO, true>(own_w, push, s_delta + warp, q, y_begin, y_end, g.Y, ring, edge_ring, lane, kq, grp_row0,
                                                    (uint32_t)colour, sweep, rk, s_dn, s_up, kc, wil_plane, cd.wil_w,
                                                    accepted);
        else
            stream_chunk<SEGS, WILSON, COUNT, HALO, false>(own_w, push, s_delta + warp, q, y_begin, y_end, g.Y, ring, edge_ring, lane, kq, grp_row0,
                                                     (uint32_t)colour, sweep, rk, s_dn, s_up, kc, wil_plane, cd.wil_w,
                                                     accepted);
    }
    if (COUNT)
        block_count_flush(&s_cnt, (unsigned int)accepted, acc + (size_t)chain * kAccSlots + (blockIdx.x % kAccSlots));
    if (HALO && halo.enabled) {
        // the last boundary block to finish publishes this launch's phase to both neighbours
        __syncthreads();
        if (boundary_block && threadIdx.x == 0) {
            // release chain instead of sequentially-consistent fences (a MEMBAR.SC.SYS here costs ~10 us):
            // the block's stores happen-before the bar.sync above, thread 0 releases them with the ticket
            // increment, the last block acquires every ticket and releases the flag at system scope.
            const unsigned int done = atom_add_acq_rel_gpu(&halo.mine->ticket, 1u);
            if (done + 1u == halo.n_boundary_blocks) {
                st_relaxed_gpu(&halo.mine->ticket, 0u);
                st_release_sys(&halo.lo_sync->from_hi, halo.phase);
                st_release_sys(&halo.hi_sync->from_lo, halo.phase);
            }
        }
    }
}

struct StreamPlan {
    bool enabled = false;
    int segs = 0;
    int blocks_per_sm = 0;
    int sm_count = 0;
};

inline size_t stream_smem_bytes(int segs, uint32_t max_thr) {
    const uint32_t kc = max_thr > 4 ? max_thr - 1 : 3;
    return (size_t)kStreamDepth * 4 * kStreamThreads * 16 + (segs > 1 ? (size_t)kStreamDepth * kStreamThreads * 4 : 0) +
           sizeof(uint2) * 2 * (2 * kc + 1);
}

template <int SEGS>
inline cudaError_t stream_occupancy(int* blocks, size_t smem, bool halo) {
    cudaError_t e = cudaFuncSetAttribute(k_sweep_stream<SEGS, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k_sweep_stream<SEGS, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sweep_stream<SEGS, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (halo) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k_sweep_stream<SEGS, false, false, true>, kStreamThreads, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, k_sweep_stream<SEGS, false, false, false>, kStreamThreads, smem);
}

// Applies to X in {256, 512, 1024}.
inline cudaError_t stream_prepare(StreamPlan& p, const Geom& g, int sm_count, uint32_t max_thr, bool force_off, bool halo) {
    p.enabled = false;
    if (force_off) return cudaSuccess;
    if (g.X != 256 && g.X != 512 && g.X != 1024) return cudaSuccess;
    if ((uint64_t)g.X * g.Y * g.T >= (1ull << 35) || g.chain_stride >= (1ull << 31)) return cudaSuccess;
    p.segs = (int)(g.X / 256);
    p.sm_count = sm_count;
    const size_t smem = stream_smem_bytes(p.segs, max_thr);
    if (smem > 200 * 1024) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    switch (p.segs) {
        case 1: e = stream_occupancy<1>(&p.blocks_per_sm, smem, halo); break;
        case 2: e = stream_occupancy<2>(&p.blocks_per_sm, smem, halo); break;
        default: e = stream_occupancy<4>(&p.blocks_per_sm, smem, halo); break;
    }
    if (e != cudaSuccess) return e;
    p.enabled = p.blocks_per_sm > 0;
    return cudaSuccess;
}

// Number of y chunks per plane group.  Blocks are resident for their whole chunk, so a launch lasts about
// ceil(waves) * (rows per chunk + a per-chunk prologue worth ~3 rows): pick the count that minimises it.
inline uint32_t stream_chunks(const StreamPlan& p, const Geom& g, uint32_t n_planes, uint32_t n_chains) {
    const uint64_t slots = (uint64_t)p.sm_count * p.blocks_per_sm;
    const uint64_t base = (uint64_t)p.segs * ((n_planes + kStreamWarps - 1) / kStreamWarps) * n_chains;
    const uint32_t max_chunks = g.Y / 4 ? g.Y / 4 : 1;  // at least 4 rows per chunk
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t nc = 1; nc <= max_chunks; ++nc) {
        const uint64_t blocks = base * nc;
        const uint64_t waves = (blocks + slots - 1) / slots;
        const double rows = (double)((g.Y / 2 + nc - 1) / nc) * 2.0;  // longest chunk
        const double cost = (double)waves * (rows + 3.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = nc;
        }
    }
    return best;
}

template <int SEGS>
inline cudaError_t stream_launch(const StreamPlan& p, const MarchArgs& a) {
    const uint32_t n_chunks = stream_chunks(p, a.g, a.n_planes, a.n_chains);
    const uint32_t n_groups = (a.n_planes + kStreamWarps - 1) / kStreamWarps;
    dim3 grid(n_chunks * SEGS, n_groups, a.n_chains);
    const size_t smem = stream_smem_bytes(SEGS, a.max_thr);
    HaloP2P halo = a.halo;
    uint32_t n_bchunks = 0;
    if (halo.enabled) {
        // boundary plane group in chunks a fraction as long as the interior ones (at least 4 rows each): the
        // earlier they finish, the more slack the flag has to cross NVLink before the next phase starts
        constexpr uint32_t factor = 4;
        const uint32_t max_chunks = a.g.Y / 4 ? a.g.Y / 4 : 1;
        n_bchunks = factor * n_chunks < max_chunks ? factor * n_chunks : max_chunks;
        if (n_bchunks == 0) n_bchunks = 1;
        grid = dim3(SEGS * (n_bchunks + (n_groups - 1) * n_chunks), 1, a.n_chains);
        halo.n_boundary_blocks = SEGS * n_bchunks * a.n_chains;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kStreamThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
#define LQFT_GO(W, C, H)                                                                                                 \
    return cudaLaunchKernelEx(&cfg, k_sweep_stream<SEGS, W, C, H>, a.g, a.colour, a.sweep, a.ctr, a.tl_first, a.n_planes, \
                              n_chunks, n_bchunks, a.own, a.oth, a.chains, a.thr, a.acc, a.max_thr, halo)
    if (halo.enabled) {
        if (a.wilson) {
            if (a.count) LQFT_GO(true, true, true); else LQFT_GO(true, false, true);
        } else {
            if (a.count) LQFT_GO(false, true, true); else LQFT_GO(false, false, true);
        }
    }
    if (a.wilson) {
        if (a.count) LQFT_GO(true, true, false); else LQFT_GO(true, false, false);
    } else {
        if (a.count) LQFT_GO(false, true, false); else LQFT_GO(false, false, false);
    }
#undef LQFT_GO
    return cudaGetLastError();
}

inline cudaError_t stream_half(const StreamPlan& p, const MarchArgs& a) {
    switch (p.segs) {
        case 1: return stream_launch<1>(p, a);
        case 2: return stream_launch<2>(p, a);
        default: return stream_launch<4>(p, a);
    }
}

}  // namespace lqft

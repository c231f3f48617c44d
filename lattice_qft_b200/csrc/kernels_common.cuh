// kernels_common.cuh — what the 16-bit kernels (kernels_narrow.cuh, kernels_cluster.cuh) and the one-CTA run kernel
// (kernels_resident.cuh) share: programmatic dependent launch, Philox with hoisted round keys, the two-sided decision
// table of the i32 resident kernel, and the flag words of the peer-memory halo.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_v1.cuh"

namespace lqft {

// Programmatic dependent launch: a sweep kernel lets its successor start (and run its prologue) at
// once, and touches field data / the device sweep counter only after its predecessor has completed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Loads that follow the wait for the previous launch directly.  A non-coherent load (__ldg, or any load through a
// const __restrict__ pointer) may be hoisted above griddepcontrol.wait — ptxas does it (seen in the SASS of a slab
// instantiation: five LDG.E.128.CONSTANT in front of ACQBULK, wrong acceptance counts) — so the first loads behind the wait
// are plain coherent ones, issued by volatile statements; the loads of the later row steps sit behind a branch.
__device__ __forceinline__ uint4 ld_after_wait(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_after_wait(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_after_wait(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

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

// ---- peer-memory halo (world_size > 1): flag words ------------------------------------------------------------
// Each rank keeps one HaloSync in device memory, mapped into both neighbours through CUDA IPC.  The sweep kernel
// of the 16-bit storage (kernels_narrow.cuh) raises the neighbours' words when a period of its ghost planes is used up
// and waits for theirs before it reads the other set.
struct HaloSync {  // lives in each rank's device memory, mapped into both neighbours
    unsigned long long from_lo;   // last phase published by the lower neighbour (rank - 1)
    unsigned long long from_hi;   // last phase published by the upper neighbour (rank + 1)
    unsigned int ticket;          // boundary blocks of the running launch that have finished
    unsigned int error;           // set when a flag wait timed out
};


__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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


}  // namespace lqft

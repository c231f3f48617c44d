// i16_proto.cu — is 16-bit height storage worth building?  (round-2 candidate)
//
// Same red-black Metropolis half-sweep and the same decision spec as the product (Philox keying, two-sided
// threshold table), but the colour arrays hold int16 heights and one thread owns 8 same-colour sites (one
// 16-byte load).  A half-warp covers one row of X = 256 (128 positions), a warp two adjacent t-planes, so the
// inner t-neighbour comes from the partner half-warp by SHFL; neighbour sums are formed on packed pairs
// (heights biased by +4096 so that 6-term sums do not carry across the halves).  The result is compared bit
// for bit with the product's scalar kernel (k_sweep_generic on an i32 copy).
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I lattice_qft_b200/csrc \
//              -o experiments/i16_proto experiments/i16_proto.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "kernels_tiled.cuh"

using namespace lqft;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kBias = 2048;  // heights within +-2047: 6-term packed sums and 6 v - sum + 0x6000 stay inside 16 bits
#ifndef WARPS
#define WARPS 4  // per block: 2 * WARPS planes
#endif
#ifndef MINB
#define MINB 7
#endif

// packed helpers: a register holds two biased heights (lo = even element, hi = odd element)
__device__ __forceinline__ int lo16(uint32_t p) { return (int)(p & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t p) { return (int)(p >> 16); }

// interleaved table: entry (i << 1) | coin for d = i - kc; coin 1 proposes +1 (m = d), coin 0 proposes -1 (m = -d)
__device__ __forceinline__ int build_interleaved(uint2* s_tab, const uint32_t* __restrict__ thr, uint32_t n_thr) {
    const int kc = max((int)n_thr - 1, 3);
    for (int i = threadIdx.x; i < 2 * (2 * kc + 1); i += blockDim.x) {
        const int up = i & 1, d = (i >> 1) - kc;
        const int m = up ? d : -d;
        const int k = min(max(m + 3, 0), (int)n_thr - 1);
        s_tab[i] = make_uint2(thr[k] | (up ? 0x80000000u : 0u), up ? 1u : 0xffffffffu);
    }
    __syncthreads();
    return kc;
}

// two sites at once: v, ns packed biased pairs; returns the updated packed pair
__device__ __forceinline__ uint32_t decide2(uint32_t v, uint32_t ns, uint32_t r_lo, uint32_t r_hi, const uint2* s_tab, uint32_t c_off, uint32_t c_max) {
    const uint32_t u = 6u * v + (0x60006000u - ns);                 // halves: d + 24576, no borrow
    const uint32_t idx = __viaddmin_s16x2_relu(u, c_off, c_max);    // halves: clamp(d + kc, 0, 2 kc)
    const uint32_t i_lo = __funnelshift_l(r_lo, idx & 0xffffu, 1);  // (i << 1) | coin
    const uint32_t i_hi = __funnelshift_l(r_hi, idx >> 16, 1);
    const uint2 t_lo = s_tab[i_lo], t_hi = s_tab[i_hi];
    const uint32_t d_lo = r_lo <= t_lo.x ? t_lo.y : 0u, d_hi = r_hi <= t_hi.x ? t_hi.y : 0u;
    return v + d_lo + (d_hi << 16);
}

// one site: d from unbiased values
__device__ __forceinline__ int decide(int v, int nsum, uint32_t r, const uint2* s_dn, const uint2* s_up, int kc) {
    const int d = min(max(6 * v - nsum, -kc), kc);
    const uint2 t = ((int)r < 0 ? s_up : s_dn)[d];
    return r <= t.x ? v + (int)t.y : v;
}

constexpr int kKcMax = 255;
__shared__ uint2 s_tab[2 * (2 * kKcMax + 1)];

// sums and decision for one row; P = x parity of the row's sites (uniform over the warp: diagonal pairing)
template <int P>
__device__ __forceinline__ uint4 row_update(const uint4 v, const uint4 om, const uint4 oc, const uint4 on, const uint4 to, const uint4 ti,
                                            uint32_t edge, const u32x4 r0, const u32x4 r1, uint32_t c_off, uint32_t c_max) {
    uint4 sh;
    if (P) {
        sh.x = __byte_perm(oc.x, oc.y, 0x5432); sh.y = __byte_perm(oc.y, oc.z, 0x5432);
        sh.z = __byte_perm(oc.z, oc.w, 0x5432); sh.w = __byte_perm(oc.w, edge, 0x5432);
    } else {
        sh.x = __byte_perm(edge, oc.x, 0x5432); sh.y = __byte_perm(oc.x, oc.y, 0x5432);
        sh.z = __byte_perm(oc.y, oc.z, 0x5432); sh.w = __byte_perm(oc.z, oc.w, 0x5432);
    }
    constexpr uint32_t C = 0x60006000u;
    uint4 w;
#define LQ_PAIR(F, RA, RB)                                                                 \
    {                                                                                      \
        const uint32_t ns = oc.F + sh.F + om.F + on.F + to.F + ti.F;                       \
        const uint32_t idx = __viaddmin_s16x2_relu(6u * v.F + C - ns, c_off, c_max);       \
        const uint2 ta = s_tab[__funnelshift_l(RA, idx & 0xffffu, 1)];                     \
        const uint2 tb = s_tab[__funnelshift_l(RB, idx >> 16, 1)];                         \
        w.F = v.F + ((RA) <= ta.x ? ta.y : 0u) + (((RB) <= tb.x ? tb.y : 0u) << 16);       \
    }
    LQ_PAIR(x, r0.x, r0.y)
    LQ_PAIR(y, r0.z, r0.w)
    LQ_PAIR(z, r1.x, r1.y)
    LQ_PAIR(w, r1.z, r1.w)
#undef LQ_PAIR
    return w;
}

// grid = (nch, T / (2 * WARPS)), block = WARPS * 32.  X = 256 only.  Lanes 0-15: plane t, row y; lanes 16-31: plane t+1, row y+1
// (same x parity), the inner t-neighbour row comes from the partner half-warp's window.
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_sweep_i16(uint32_t Y, uint32_t T, uint32_t nch, int colour, uint64_t sweep, uint16_t* __restrict__ own, const uint16_t* __restrict__ oth,
            uint32_t k0, uint32_t k1, const uint32_t* __restrict__ thr, uint32_t n_thr) {
    pdl_launch_dependents();
    const int kc = build_interleaved(s_tab, thr, n_thr);
    const uint32_t c_off = (uint32_t)((kc - 24576) & 0xffff) * 0x10001u, c_max = (uint32_t)(2 * kc) * 0x10001u;
    pdl_wait();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t half = lane >> 4, l16 = lane & 15u;
    const uint32_t t = (blockIdx.y * WARPS + warp) * 2u + half;
    const uint32_t y0 = blockIdx.x * Y / nch, y1 = (blockIdx.x + 1u) * Y / nch;
    constexpr uint32_t ROWQ = 16;
    const uint32_t planeq = Y * ROWQ;
    const uint32_t tp = (t + 1 == T) ? 0 : t + 1, tm = (t == 0) ? T - 1 : t - 1;
    // offsets in uint4 units from the array bases (arrays < 2^32 uint4)
    const uint32_t b_own = t * planeq + l16, b_out = (half ? tp : tm) * planeq + l16;
    uint4* const q_own = reinterpret_cast<uint4*>(own);
    const uint4* const q_oth = reinterpret_cast<const uint4*>(oth);
    const RoundKeys rk = make_round_keys(k0, k1);
    uint32_t y = y0 + half; if (y == Y) y = 0;           // this lane's first row
    const uint32_t ym = (y == 0) ? Y - 1 : y - 1;
    uint4 wa = __ldg(q_oth + b_own + ym * ROWQ), wb = __ldg(q_oth + b_own + y * ROWQ), wc;
    uint32_t p = (y0 + (blockIdx.y * WARPS + warp) * 2u + (uint32_t)colour) & 1u;   // parity of the lower half's row = the upper half's
    const uint32_t src_r = (lane & 16u) | ((lane + 1u) & 15u), src_l = (lane & 16u) | ((lane - 1u) & 15u);
    const uint32_t grp_base = t * Y * 32u + l16 * 2u;
    uint32_t n = y1 - y0;
#ifndef PREFETCH
#define LQ_STEP(OM, OC, ON)                                                                                            \
    {                                                                                                                  \
        const uint32_t yn = (y + 1 == Y) ? 0 : y + 1;                                                                  \
        uint4* const po = q_own + (b_own + y * ROWQ);                                                                  \
        const uint4 v = *po;                                                                                           \
        ON = __ldg(q_oth + (b_own + yn * ROWQ));                                                                       \
        const uint4 to = __ldg(q_oth + (b_out + y * ROWQ));                                                            \
        LQ_BODY(OM, OC, ON)                                                                                            \
    }
#else
    // one step ahead: row y's own / t-outer rows and row y+1's other-colour row are already in registers
    uint4 nv = q_own[b_own + y * ROWQ], nto = __ldg(q_oth + (b_out + y * ROWQ));
    wc = __ldg(q_oth + (b_own + ((y + 1 == Y) ? 0 : y + 1) * ROWQ));
#define LQ_STEP(OM, OC, ON)                                                                                            \
    {                                                                                                                  \
        const uint32_t yn = (y + 1 == Y) ? 0 : y + 1;                                                                  \
        const uint32_t ynn = (yn + 1 == Y) ? 0 : yn + 1;                                                               \
        uint4* const po = q_own + (b_own + y * ROWQ);                                                                  \
        const uint4 v = nv, to = nto;                                                                                  \
        nv = q_own[b_own + yn * ROWQ];                                                                                 \
        nto = __ldg(q_oth + (b_out + yn * ROWQ));                                                                      \
        OM##_next = __ldg(q_oth + (b_own + ynn * ROWQ));                                                               \
        LQ_BODY(OM, OC, ON)                                                                                            \
        OM = OM##_next;                                                                                                \
    }
    uint4 wa_next, wb_next, wc_next;
#endif
#define LQ_BODY(OM, OC, ON)                                                                                            \
        uint4 snd = half ? OM : ON, ti;                                                                                \
        ti.x = __shfl_xor_sync(0xffffffffu, snd.x, 16); ti.y = __shfl_xor_sync(0xffffffffu, snd.y, 16);                \
        ti.z = __shfl_xor_sync(0xffffffffu, snd.z, 16); ti.w = __shfl_xor_sync(0xffffffffu, snd.w, 16);                \
        const uint32_t grp = grp_base + y * 32u;                                                                       \
        const u32x4 r0 = philox4x32_rk(grp, (uint32_t)colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);           \
        const u32x4 r1 = philox4x32_rk(grp + 1u, (uint32_t)colour, (uint32_t)sweep, (uint32_t)(sweep >> 32), rk);      \
        uint4 w;                                                                                                       \
        if (p) {                                                                                                       \
            const uint32_t edge = __shfl_sync(0xffffffffu, OC.x, src_r);                                               \
            w = row_update<1>(v, OM, OC, ON, to, ti, edge, r0, r1, c_off, c_max);                                      \
        } else {                                                                                                       \
            const uint32_t edge = __shfl_sync(0xffffffffu, OC.w, src_l);                                               \
            w = row_update<0>(v, OM, OC, ON, to, ti, edge, r0, r1, c_off, c_max);                                      \
        }                                                                                                              \
        *po = w;                                                                                                       \
        y = yn;                                                                                                        \
        p ^= 1u;
    for (;;) {
        LQ_STEP(wa, wb, wc)
        if (--n == 0) break;
        LQ_STEP(wb, wc, wa)
        if (--n == 0) break;
        LQ_STEP(wc, wa, wb)
        if (--n == 0) break;
    }
#undef LQ_BODY
#undef LQ_STEP
}

__global__ void k_to16(const int32_t* a, uint16_t* b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) b[i] = (uint16_t)(a[i] + kBias);
}
__global__ void k_cmp(const int32_t* a, const uint16_t* b, size_t n, unsigned long long* bad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n && a[i] + kBias != (int)b[i]) atomicAdd(bad, 1ull);
}

int main(int argc, char** argv) {
    const uint32_t X = 256, Y = 256, T = 256;
    const double beta = 0.25;
    const uint64_t seed = 0xC0FFEE;
    Geom g{};
    g.X = X; g.Y = Y; g.T = T; g.Xh = X / 2; g.Tl = T; g.t0 = 0; g.ghost = 0; g.plane = g.Xh * Y; g.chain_stride = (uint64_t)g.plane * T;
    const size_t half = g.chain_stride;
    int32_t* f32[2];
    uint16_t* f16[2];
    for (int c = 0; c < 2; ++c) { CK(cudaMalloc(&f32[c], half * 4)); CK(cudaMalloc(&f16[c], half * 2)); }
    // threshold table (same as lqft_host_threshold_table)
    std::vector<uint32_t> thr;
    for (int k = 0; k < 2048; ++k) {
        const double prob = exp((double)(-(6 + 2 * (k - 3))) * beta);
        const double sc = floor(prob * 2147483648.0);
        thr.push_back(sc >= 2147483647.0 ? 2147483647u : (uint32_t)sc);
        if (thr.back() == 0) break;
    }
    uint32_t* thr_d; CK(cudaMalloc(&thr_d, 2048 * 4)); CK(cudaMemset(thr_d, 0, 2048 * 4));
    CK(cudaMemcpy(thr_d, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice));
    ChainDev cd{}; cd.k0 = (uint32_t)seed; cd.k1 = (uint32_t)(seed >> 32); cd.n_thr = (uint32_t)thr.size();
    ChainDev* cd_d; CK(cudaMalloc(&cd_d, sizeof cd)); CK(cudaMemcpy(cd_d, &cd, sizeof cd, cudaMemcpyHostToDevice));
    unsigned long long* acc; CK(cudaMalloc(&acc, 8 * kAccSlots)); CK(cudaMemset(acc, 0, 8 * kAccSlots));
    dim3 gi((g.plane + 255) / 256, T, 1);
    k_init_hot<<<gi, 256>>>(g, f32[0], f32[1], cd_d);
    // thermalise with the product's scalar kernel, then copy to i16
    const size_t smem_v1 = 4 * thr.size();
    for (int s = 0; s < 50; ++s)
        for (int c = 0; c < 2; ++c)
            k_sweep_generic<false><<<gi, 256, smem_v1>>>(g, c, (uint64_t)s, nullptr, 0, f32[c], f32[c ^ 1], cd_d, thr_d, acc);
    for (int c = 0; c < 2; ++c) k_to16<<<(half + 255) / 256, 256>>>(f32[c], f16[c], half);
    CK(cudaDeviceSynchronize());
    const uint32_t kc = (uint32_t)thr.size() - 1;
    uint32_t nch = argc > 1 ? atoi(argv[1]) : 27;
    dim3 grid(nch, T / (2 * WARPS));
    // correctness: 3 sweeps both ways
    for (int s = 50; s < 53; ++s)
        for (int c = 0; c < 2; ++c) {
            k_sweep_generic<false><<<gi, 256, smem_v1>>>(g, c, (uint64_t)s, nullptr, 0, f32[c], f32[c ^ 1], cd_d, thr_d, acc);
            k_sweep_i16<<<grid, WARPS * 32>>>(Y, T, nch, c, (uint64_t)s, f16[c], f16[c ^ 1], cd.k0, cd.k1, thr_d, cd.n_thr);
        }
    unsigned long long* bad; CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    for (int c = 0; c < 2; ++c) k_cmp<<<(half + 255) / 256, 256>>>(f32[c], f16[c], half, bad);
    unsigned long long nbad = 0; CK(cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost));
    printf("mismatching sites after 3 sweeps: %llu of %zu\n", nbad, 2 * half);
    // timing
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < (argc > 2 ? argc - 2 : 1) * 2; ++rep) {
        if (argc > 2) { nch = atoi(argv[2 + rep / 2]); grid.x = nch; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaEventRecord(a, st);
        for (int s = 100; s < 300; ++s)
            for (int c = 0; c < 2; ++c)
                CK(cudaLaunchKernelEx(&cfg, k_sweep_i16, Y, T, nch, c, (uint64_t)s, f16[c], (const uint16_t*)f16[c ^ 1], cd.k0, cd.k1, (const uint32_t*)thr_d, cd.n_thr));
        cudaEventRecord(b, st); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("i16 proto nch %u: %.2f us per half-sweep, %.3g site updates/s\n", nch, ms * 1e3 / 400, (double)X * Y * T * 200 / (ms * 1e-3));
    }
    return nbad != 0;
}

// lqft.cu — the C ABI of include/lqft.h: handle management, launch schedules, CUDA graphs,
// NCCL / peer-memory halo exchange.  All arithmetic that touches the field runs in the CUDA
// kernels of the kernels_*.cuh headers; this file contains no CPU implementation of the sweep or of
// the observables.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/lqft.h"
#include "kernels_f64.cuh"
#include "kernels_resident.cuh"
#include "kernels_narrow.cuh"
#include "kernels_cluster.cuh"
#include "kernels_observe.cuh"
#include "kernels_v1.cuh"
#include "spec.cuh"

using namespace lqft;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static lqft_status fail(lqft_status s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return s;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(LQFT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define ST(call)                        \
    do {                                \
        lqft_status s_ = (call);        \
        if (s_ != LQFT_OK) return s_;   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that the library loads (and single-GPU handles work) without it
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static lqft_status nccl_load() {
    if (g_nccl.lib) return LQFT_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(LQFT_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                          \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                  \
    if (!g_nccl.field) return fail(LQFT_ERR_NCCL, "libnccl lacks symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return LQFT_OK;
}

#define NC(call)                                                                                 \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            return fail(LQFT_ERR_NCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), \
                        __FILE__, __LINE__);                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct GraphKey {
    uint32_t energy_every, normalize_every, action_every, correlator_every, correlator_reps, difference_every, cycle;
    bool count, narrow;
    uint32_t gbuf;  // slabs: the ghost set that is current when a replay begins (kernel arguments bake it)
    bool operator<(const GraphKey& o) const {
        return std::tie(energy_every, normalize_every, action_every, correlator_every, correlator_reps, difference_every, cycle,
                        count, narrow, gbuf) <
               std::tie(o.energy_every, o.normalize_every, o.action_every, o.correlator_every, o.correlator_reps,
                        o.difference_every, o.cycle, o.count, o.narrow, o.gbuf);
    }
};
struct GraphVal {
    cudaGraphExec_t exec = nullptr;
    uint64_t kernels = 0;
    uint64_t phases = 0;  // halo phases one replay goes through (deep-halo exchanges of the narrow copy)
};

struct lqft_handle {
    lqft_config cfg{};
    Geom g{};
    uint64_t n_slab = 0;       // sites of this rank's slab (per chain)
    uint64_t n_global = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_bnd = nullptr, ev_comm = nullptr;
    bool timed = false;
    int32_t* field[2] = {nullptr, nullptr};
    // LQFT_F64 handles (scalar path, one GPU)
    bool is64 = false;
    double* field64[2] = {nullptr, nullptr};
    double* beta_d = nullptr;      // [n_chains]
    double* offset64_d = nullptr;  // [n_chains]
    double* obs64_d = nullptr;     // [n_chains][2]
    double* mom64_d = nullptr;     // [n_chains][T][2]
    double2* part_d = nullptr;     // per-block compensated partial sums
    double* elog64_d = nullptr;    // energy log of lqft_run
    double* gathered64_d = nullptr;
    std::vector<ChainDev> chains_h;
    std::vector<double> beta;
    std::vector<uint8_t> configured;
    std::vector<uint32_t> thr_h;
    ChainDev* chains_d = nullptr;
    uint32_t* thr_d = nullptr;
    bool params_dirty = true;
    bool any_wilson = false;
    uint32_t max_thr = 0;
    unsigned long long* acc_d = nullptr;   // [n_chains][kAccSlots]
    unsigned long long* obs_d = nullptr;   // [n_chains][2]
    unsigned long long* mom_d = nullptr;   // [n_chains][T][2]
    int32_t* shift_d = nullptr;            // [n_chains] scratch: picked heights
    int32_t* offset_d = nullptr;           // [n_chains] true height = stored - offset (lazy normalisation)
    int32_t* staging_d = nullptr;          // [n_slab] lexicographic
    uint64_t* sweep_ctr_d = nullptr;       // device sweep counter for graph replays
    uint32_t* n_meas_d = nullptr;
    long long* elog_d = nullptr;           // [n_chains][elog_cap] energies logged by lqft_run
    uint32_t elog_cap = 0;
    long long* alog_d = nullptr;           // [n_chains][alog_cap] actions
    uint32_t alog_cap = 0;
    double* clog_d = nullptr;              // [n_chains][clog_cap][T] correlators
    uint32_t clog_cap = 0;
    unsigned int* ticket_d = nullptr;      // arrival counter of the observe launch (kernels_observe.cuh)
    // results of standalone measurements: pinned host memory mapped into the device, written by the finaliser itself
    // (no copy after the launch, one synchronisation)
    long long* result_h = nullptr;         // [n_chains][2] E, S
    long long* result_d = nullptr;         //   its device address
    double* cres_h = nullptr;              // [T] correlator
    double* cres_d = nullptr;
    uint64_t* run_d = nullptr;             // {sweep_base, first_step} of the running schedule, read by the finalisers
    ObsFin run_fin{};                      // where the running schedule logs its observables
    uint64_t* sites_d = nullptr;
    long long* gathered_d = nullptr;
    size_t gathered_cap = 0;
    uint32_t sites_cap = 0;
    std::vector<double*> diff_sum, diff_comp;
    std::vector<uint64_t> diff_count;
    bool ghosts_valid = false;
    bool count_accepts = false;  // current call wants acceptance counts
    uint64_t launches = 0;
    ncclComm_t comm = nullptr;
    int sm_count = 148;
    std::map<GraphKey, GraphVal> graphs;
    bool resident_ok = false;   // small lattice: whole runs inside one launch (kernels_resident.cuh)
    size_t smem_optin = 0;
    // peer-memory halo (world_size > 1)
    bool p2p = false;
    HaloSync* halo_d = nullptr;                 // mine
    void* peer_base[2][5] = {{nullptr, nullptr, nullptr, nullptr, nullptr},
                             {nullptr, nullptr, nullptr, nullptr, nullptr}};  // [lo/hi][field0, field1, sync, narrow0, narrow1]
    bool peer_same = false;                     // lo and hi are the same rank (world_size == 2)
    bool self_halo = false;                     // LQFT_FLAG_SELF_HALO: this rank is its own neighbour
    unsigned long long phase = 0;               // operations that touch ghost planes so far (same on every rank)
    // 16-bit copy of the field (kernels_narrow.cuh): the storage of admitted lattices, also between calls; the i32
    // arrays are then allocated only if a field ever leaves the band (single GPU) or for the i32 halo path (slabs)
    uint16_t* narrow[2] = {nullptr, nullptr};
    int32_t* ncentre_d = nullptr;               // [n_chains] centre of the bias while a narrow segment runs
    uint32_t* nworst_d = nullptr;               // validity word of k_narrow
    NarrowPlan narrowp{};
    ClusterPlan clusterp{};                     // chains resident in the shared memory of a cluster (kernels_cluster.cuh)
    Geom gn{};                                  // geometry of the narrow arrays: slabs carry a deep halo (ghost > 1)
    uint32_t nvalid = 0;                        // ghost planes per side of the narrow copy that are valid right now
    unsigned long long* phase_d = nullptr;      // device copy of `phase` at the start of a graph replay (slabs, narrow copy)
    bool next_ready = false;                    // the OTHER ghost set of the narrow copy holds this rank's pushed planes
    bool capturing = false;                     // launches are being captured: flip numbers are relative to *phase_d
    unsigned long long phase_capture0 = 0;      // `phase` when the capture began
    uint32_t tl_lo = 0, tl_hi = 0;              // owned planes of the lower / upper neighbour
    bool narrow_active = false;                 // the live copy of the field is the 16-bit one (it stays so between calls)
    uint32_t narrow_budget = 0;                 // sweeps the 16-bit copy may still take before it has to be re-measured
    uint32_t narrow_retry = 0;                  // sweeps to stay on i32 before the band is tried again (after a refusal)
    uint64_t narrow_segments = 0;
};

static lqft_status use_device(lqft_handle* h) {
    CU(cudaSetDevice(h->cfg.device));
    return LQFT_OK;
}

static lqft_status narrow_exit(lqft_handle* h);
static lqft_status need_i32(lqft_handle* h);

static lqft_status upload_params(lqft_handle* h) {
    if (!h->params_dirty) return LQFT_OK;
    const uint32_t nc = h->cfg.n_chains;
    for (uint32_t c = 0; c < nc; ++c)
        if (!h->configured[c])
            return fail(LQFT_ERR_STATE, "chain %u has no parameters: call lqft_set_chain first", c);
    CU(cudaMemcpyAsync(h->chains_d, h->chains_h.data(), sizeof(ChainDev) * nc,
                       cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->thr_d, h->thr_h.data(), sizeof(uint32_t) * (size_t)nc * kThrStride,
                       cudaMemcpyHostToDevice, h->stream));
    // the async copies read pageable std::vector storage: fence before it can change
    CU(cudaStreamSynchronize(h->stream));
    h->any_wilson = false;
    h->max_thr = 0;
    for (uint32_t c = 0; c < nc; ++c) {
        if (h->chains_h[c].wil_w && h->chains_h[c].wil_h) h->any_wilson = true;
        h->max_thr = std::max(h->max_thr, h->chains_h[c].n_thr);
    }
    h->params_dirty = false;
    if (h->is64) {
        CU(cudaMemcpyAsync(h->beta_d, h->beta.data(), sizeof(double) * nc, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    h->resident_ok = !h->is64 && !(h->cfg.flags & (LQFT_FLAG_FORCE_GENERIC | LQFT_FLAG_NO_RESIDENT)) &&
                     resident_fits(h->g, h->max_thr, h->smem_optin);
    if (h->resident_ok) {
        const int smem = (int)resident_smem_bytes(h->g, h->max_thr);
        CU(cudaFuncSetAttribute(k_run_resident<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(k_run_resident<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(k_run_resident<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(k_run_resident<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(k_run_resident<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CU(cudaFuncSetAttribute(k_run_resident<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    CU(narrow_prepare(h->narrowp, h->g, nc, h->sm_count, h->max_thr,
                      !h->narrow[0] || h->resident_ok || (h->g.ghost != 0 && !h->p2p)));
    CU(cluster_prepare(h->clusterp, h->g, nc, h->sm_count, h->max_thr, h->smem_optin,
                       !h->narrowp.enabled || (h->cfg.flags & LQFT_FLAG_NO_RESIDENT) != 0));
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
    if (h->narrow_active && !h->narrowp.enabled) ST(narrow_exit(h));  // e.g. a coupling whose table the 16-bit kernel cannot hold
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// host spec functions
// ---------------------------------------------------------------------------------------------
extern "C" uint32_t lqft_version(void) { return (LQFT_VERSION_MAJOR << 16) | LQFT_VERSION_MINOR; }
extern "C" const char* lqft_last_error(void) { return g_err.c_str(); }

extern "C" int32_t lqft_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void lqft_host_philox4x32(uint64_t seed, const uint32_t ctr[4], uint32_t out[4]) {
    u32x4 r = philox4x32(ctr[0], ctr[1], ctr[2], ctr[3], (uint32_t)seed, (uint32_t)(seed >> 32));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

extern "C" uint32_t lqft_host_site_word(uint64_t seed, uint64_t idx, uint32_t colour, uint64_t sweep) {
    const uint64_t grp = idx >> 3;
    u32x4 r = philox4x32((uint32_t)grp, colour | ((uint32_t)(grp >> 32) << 8), (uint32_t)sweep,
                         (uint32_t)(sweep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    return word_of(r, (uint32_t)((idx >> 1) & 3u));
}

extern "C" uint32_t lqft_host_threshold_table(double beta, uint32_t* thr, uint32_t capacity) {
    if (!(beta > 0.0) || !thr) return 0;
    for (uint32_t k = 0; k < capacity; ++k) {
        const int m = (int)k - 3;
        // (old_action - new_action) as f64 * temp, then exp: algorithm.rs:152
        const double prob = exp((double)(-(6 + 2 * m)) * beta);
        const double scaled = floor(prob * 2147483648.0);
        const uint32_t t = scaled >= 2147483647.0 ? 2147483647u : (uint32_t)scaled;
        thr[k] = t;
        if (t == 0u) return k + 1;
    }
    return 0;  // did not reach a zero entry within capacity
}

extern "C" uint64_t lqft_host_pick_site(uint64_t seed, uint32_t stream, uint64_t counter,
                                        uint32_t rep, uint64_t n_sites) {
    return pick_site((uint32_t)seed, (uint32_t)(seed >> 32), stream, counter, rep, n_sites);
}

extern "C" int32_t lqft_host_hot_value(uint64_t seed, uint64_t idx) {
    return hot_value((uint32_t)seed, (uint32_t)(seed >> 32), idx);
}

extern "C" uint64_t lqft_host_index(uint32_t X, uint32_t Y, uint32_t T, uint32_t x, uint32_t y,
                                    uint32_t t) {
    return (uint64_t)(x % X) + (uint64_t)X * ((uint64_t)(y % Y) + (uint64_t)Y * (uint64_t)(t % T));
}

extern "C" void lqft_host_coords(uint32_t X, uint32_t Y, uint32_t T, uint64_t idx, uint32_t xyz[3]) {
    xyz[0] = (uint32_t)(idx % X);
    xyz[1] = (uint32_t)((idx / X) % Y);
    xyz[2] = (uint32_t)((idx / ((uint64_t)X * Y)) % T);
}

extern "C" void lqft_host_neighbours(uint32_t X, uint32_t Y, uint32_t T, uint64_t idx, uint64_t nbr[6]) {
    uint32_t c[3];
    lqft_host_coords(X, Y, T, idx, c);
    nbr[0] = lqft_host_index(X, Y, T, c[0] + 1, c[1], c[2]);
    nbr[1] = lqft_host_index(X, Y, T, c[0], c[1] + 1, c[2]);
    nbr[2] = lqft_host_index(X, Y, T, c[0], c[1], c[2] + 1);
    nbr[3] = lqft_host_index(X, Y, T, c[0] + X - 1, c[1], c[2]);
    nbr[4] = lqft_host_index(X, Y, T, c[0], c[1] + Y - 1, c[2]);
    nbr[5] = lqft_host_index(X, Y, T, c[0], c[1], c[2] + T - 1);
}

extern "C" int32_t lqft_host_wilson_offset(uint32_t Y, uint32_t x, uint32_t y, uint32_t t,
                                           uint32_t width, uint32_t height) {
    return wilson_offset(Y, x, y, t, width, height);
}

extern "C" lqft_status lqft_host_slab(uint32_t T, uint32_t ws, uint32_t rank, uint32_t* t0,
                                      uint32_t* nt) {
    if (ws == 0 || rank >= ws) return fail(LQFT_ERR_INVALID, "rank %u outside world of %u", rank, ws);
    if (T % ws != 0 || (ws > 1 && (T / ws) % 2 != 0))
        return fail(LQFT_ERR_INVALID, "T=%u does not split into %u slabs of an even plane count", T, ws);
    if (t0) *t0 = rank * (T / ws);
    if (nt) *nt = T / ws;
    return LQFT_OK;
}

extern "C" uint64_t lqft_run_measurements(uint64_t first_step, uint64_t n_steps, uint32_t every) {
    if (every == 0 || n_steps == 0) return 0;
    // multiples of `every` in [first_step, first_step + n_steps)
    const uint64_t last = first_step + n_steps - 1;
    const uint64_t lo = (first_step + every - 1) / every;
    const uint64_t hi = last / every;
    return hi >= lo ? hi - lo + 1 : 0;
}

extern "C" lqft_status lqft_comm_unique_id(uint8_t id[128]) {
    ST(nccl_load());
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return LQFT_OK;
}


// ---------------------------------------------------------------------------------------------
// peer-memory halo: map the neighbours' fields and HaloSync words through CUDA IPC
// ---------------------------------------------------------------------------------------------
static lqft_status setup_p2p(lqft_handle* h) {
    h->p2p = false;
    const int ws = (int)h->cfg.world_size, r = (int)h->cfg.rank;
    struct Handles { cudaIpcMemHandle_t f0, f1, sync, n0, n1; int has_narrow; };
    Handles mine{};
    bool ok = cudaIpcGetMemHandle(&mine.f0, h->field[0]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.f1, h->field[1]) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.sync, h->halo_d) == cudaSuccess;
    if (!ok) cudaGetLastError();
    mine.has_narrow = h->narrow[0] && cudaIpcGetMemHandle(&mine.n0, h->narrow[0]) == cudaSuccess &&
                      cudaIpcGetMemHandle(&mine.n1, h->narrow[1]) == cudaSuccess;
    if (h->narrow[0] && !mine.has_narrow) cudaGetLastError();
    // every rank takes part in the exchange (and in the vote below) even if its own export failed
    Handles* all_d = nullptr;
    CU(cudaMalloc(&all_d, sizeof(Handles) * ws));
    CU(cudaMemcpyAsync(all_d + r, &mine, sizeof(Handles), cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllGather(all_d + r, all_d, sizeof(Handles), ncclInt8, h->comm, h->stream));
    std::vector<Handles> all(ws);
    CU(cudaMemcpyAsync(all.data(), all_d, sizeof(Handles) * ws, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(all_d);
    const int lo = (r + ws - 1) % ws, hi = (r + 1) % ws;
    h->peer_same = lo == hi;
    bool narrow_ok = true;  // the 16-bit storage needs every rank to have it (slabs may differ in shape)
    for (int k = 0; k < ws; ++k) narrow_ok = narrow_ok && all[k].has_narrow;
    if (ok && !(h->cfg.flags & LQFT_FLAG_HALO_NCCL)) {
        for (int side = 0; side < 2 && ok; ++side) {
            if (side == 1 && h->peer_same) {
                for (int k = 0; k < 5; ++k) h->peer_base[1][k] = h->peer_base[0][k];
                break;
            }
            const Handles& ph = all[side == 0 ? lo : hi];
            const cudaIpcMemHandle_t* hs[5] = {&ph.f0, &ph.f1, &ph.sync, &ph.n0, &ph.n1};
            for (int k = 0; k < (narrow_ok ? 5 : 3) && ok; ++k)
                if (cudaIpcOpenMemHandle(&h->peer_base[side][k], *hs[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    h->peer_base[side][k] = nullptr;
                    if (k < 3) ok = false; else narrow_ok = false;
                }
        }
    } else {
        ok = false;
    }
    // all ranks must agree, or one side would wait for flags the other never writes
    int* vote_d = nullptr;
    CU(cudaMalloc(&vote_d, 2 * sizeof(int)));
    const int v[2] = {ok ? 0 : 1, narrow_ok ? 0 : 1};
    CU(cudaMemcpyAsync(vote_d, v, sizeof v, cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllReduce(vote_d, vote_d, 2, ncclInt32, ncclSum, h->comm, h->stream));
    int bad[2] = {0, 0};
    CU(cudaMemcpyAsync(bad, vote_d, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(vote_d);
    h->p2p = bad[0] == 0;
    if (!h->p2p || bad[1] != 0) {  // no narrow copy without the peer-memory halo
        cudaFree(h->narrow[0]);
        cudaFree(h->narrow[1]);
        h->narrow[0] = h->narrow[1] = nullptr;
    }
    if (getenv("LQFT_DEBUG")) fprintf(stderr, "[lqft] rank %d/%d halo: %s\n", r, ws, h->p2p ? "peer memory (IPC)" : "NCCL send/recv");
    return LQFT_OK;
}

static lqft_status set_halo_timeout() {
    if (const char* e = getenv("LQFT_HALO_TIMEOUT_S")) {
        const double sec = atof(e);
        if (sec > 0.0) {
            const unsigned long long ns = (unsigned long long)(sec * 1e9);
            CU(cudaMemcpyToSymbol(g_halo_timeout_ns, &ns, sizeof ns));
        }
    }
    return LQFT_OK;
}

// A halo wait that expired has set the error word and trapped (kernels_common.cuh, halo_wait): the context is
// gone, so the copy below fails with LQFT_ERR_CUDA; the word itself is reported when the trap has not landed yet.
static lqft_status check_halo_error(lqft_handle* h) {
    if (!h->halo_d) return LQFT_OK;
    HaloSync s{};
    CU(cudaMemcpyAsync(&s, h->halo_d, sizeof s, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (s.error) return fail(LQFT_ERR_STATE, "halo exchange timed out waiting for a neighbour rank (phase %llu)", h->phase);
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" lqft_status lqft_destroy(lqft_handle* h) {
    if (!h) return LQFT_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int side = 0; side < 2 && !h->self_halo; ++side) {
        if (side == 1 && h->peer_same) break;
        for (int k = 0; k < 5; ++k)
            if (h->peer_base[side][k]) cudaIpcCloseMemHandle(h->peer_base[side][k]);
    }
    cudaFree(h->halo_d);
    cudaFree(h->narrow[0]); cudaFree(h->narrow[1]); cudaFree(h->ncentre_d); cudaFree(h->nworst_d); cudaFree(h->phase_d);
    cudaFree(h->field[0]); cudaFree(h->field[1]);
    cudaFree(h->field64[0]); cudaFree(h->field64[1]); cudaFree(h->beta_d); cudaFree(h->offset64_d);
    cudaFree(h->obs64_d); cudaFree(h->mom64_d); cudaFree(h->part_d); cudaFree(h->elog64_d); cudaFree(h->gathered64_d);
    cudaFree(h->chains_d); cudaFree(h->thr_d); cudaFree(h->acc_d); cudaFree(h->obs_d);
    cudaFree(h->mom_d); cudaFree(h->shift_d); cudaFree(h->offset_d); cudaFree(h->staging_d); cudaFree(h->sweep_ctr_d);
    cudaFree(h->n_meas_d); cudaFree(h->elog_d); cudaFree(h->sites_d); cudaFree(h->gathered_d);
    cudaFree(h->alog_d); cudaFree(h->clog_d); cudaFree(h->ticket_d); cudaFreeHost(h->result_h); cudaFreeHost(h->cres_h); cudaFree(h->run_d);
    for (auto p : h->diff_sum) cudaFree(p);
    for (auto p : h->diff_comp) cudaFree(p);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    if (h->ev_bnd) cudaEventDestroy(h->ev_bnd);
    if (h->ev_comm) cudaEventDestroy(h->ev_comm);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    delete h;
    return LQFT_OK;
}

// A field of zeros in the 16-bit arrays, centre 0, and these arrays the live copy.
static lqft_status narrow_zero(lqft_handle* h) {
    const uint32_t nc = h->cfg.n_chains;
    for (int c = 0; c < 2; ++c)
        k_n16_zero<<<1024, 256, 0, h->stream>>>((uint32_t*)h->narrow[c], h->gn.chain_stride * nc / 2);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(h->ncentre_d, 0, sizeof(int32_t) * nc, h->stream));
    h->narrow_active = true;
    h->narrow_budget = (uint32_t)kN16Bias - 1u;
    h->nvalid = 0;
    h->next_ready = false;
    return LQFT_OK;
}

// The i32 arrays of the field (zero heights), allocated on first need.
static lqft_status need_i32(lqft_handle* h) {
    if (h->field[0]) return LQFT_OK;
    const size_t fbytes = sizeof(int32_t) * h->g.chain_stride * h->cfg.n_chains;
    for (int c = 0; c < 2; ++c) {
        if (cudaMalloc(&h->field[c], fbytes) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(h->field[0]);
            h->field[0] = h->field[1] = nullptr;
            return fail(LQFT_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the field", fbytes);
        }
        CU(cudaMemsetAsync(h->field[c], 0, fbytes, h->stream));
    }
    return LQFT_OK;
}

static lqft_status create_impl(const lqft_config* cfg, lqft_handle* h) {
    h->cfg = *cfg;
    const uint32_t X = cfg->x, Y = cfg->y, T = cfg->t, ws = cfg->world_size ? cfg->world_size : 1;
    h->cfg.world_size = ws;
    if (X < 2 || Y < 2 || T < 2 || (X & 1) || (Y & 1) || (T & 1))
        return fail(LQFT_ERR_INVALID, "extents %ux%ux%u must be even and >= 2 (red-black order)", X, Y, T);
    if (cfg->n_chains == 0 || cfg->n_chains > 65535)
        return fail(LQFT_ERR_INVALID, "n_chains=%u outside [1,65535]", cfg->n_chains);
    if (cfg->dtype != LQFT_I32 && cfg->dtype != LQFT_F64) return fail(LQFT_ERR_INVALID, "unknown dtype %u", cfg->dtype);
    if (cfg->dtype == LQFT_F64 && (ws > 1 || (cfg->flags & LQFT_FLAG_SELF_HALO)))
        return fail(LQFT_ERR_INVALID, "LQFT_F64 fields run on one GPU (no slab decomposition)");
    h->is64 = cfg->dtype == LQFT_F64;
    if ((uint64_t)X * Y / 2 >= (1ull << 31))
        return fail(LQFT_ERR_INVALID, "a t-plane of %ux%u exceeds 2^31 elements per colour", X, Y);
    if ((uint64_t)X * Y * T >= (1ull << 40))
        return fail(LQFT_ERR_INVALID, "lattice of %llu sites is too large",
                    (unsigned long long)X * Y * T);
    uint32_t t0 = 0, nt = T;
    ST(lqft_host_slab(T, ws, cfg->rank, &t0, &nt));
    if (ws > 1 && nt > 65535) return fail(LQFT_ERR_INVALID, "slab of %u planes exceeds grid limits", nt);
    if (T > 65535) return fail(LQFT_ERR_INVALID, "T=%u exceeds grid limits", T);

    int ndev = lqft_device_count();
    if (ndev <= 0)
        return fail(LQFT_ERR_CUDA, "no CUDA device is visible; liblqft has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(LQFT_ERR_INVALID, "device %d outside [0,%d)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10)
        return fail(LQFT_ERR_CUDA, "device %d is sm_%d%d; liblqft is built for sm_100a only",
                    cfg->device, prop.major, prop.minor);
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;

    Geom& g = h->g;
    g.X = X; g.Y = Y; g.T = T; g.Xh = X / 2;
    g.Tl = nt; g.t0 = t0;
    const bool self_halo = ws == 1 && (cfg->flags & LQFT_FLAG_SELF_HALO);
    g.ghost = (ws > 1 || self_halo) ? 1 : 0;
    g.plane = g.Xh * Y;
    g.entry_bytes = (uint32_t)sizeof(uint2);
    g.chain_stride = (uint64_t)g.plane * (g.Tl + 2 * g.ghost);
    h->n_slab = (uint64_t)X * Y * nt;
    h->n_global = (uint64_t)X * Y * T;

    const uint32_t nc = cfg->n_chains;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        // the second stream carries halo traffic underneath update kernels that fill every SM: highest priority
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, greatest));
    }
    CU(cudaEventCreate(&h->ev_start));
    CU(cudaEventCreate(&h->ev_stop));
    CU(cudaEventCreateWithFlags(&h->ev_bnd, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
    if (h->is64) {
        const size_t fbytes = sizeof(double) * g.chain_stride * nc;
        for (int c = 0; c < 2; ++c) {
            if (cudaMalloc(&h->field64[c], fbytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(LQFT_ERR_NOMEM, "cannot allocate %zu bytes of device memory for the field", fbytes);
            }
            CU(cudaMemsetAsync(h->field64[c], 0, fbytes, h->stream));
        }
    }
    if (h->is64) {
        CU(cudaMalloc(&h->beta_d, sizeof(double) * nc));
        CU(cudaMalloc(&h->offset64_d, sizeof(double) * nc));
        CU(cudaMemsetAsync(h->offset64_d, 0, sizeof(double) * nc, h->stream));
        CU(cudaMalloc(&h->obs64_d, sizeof(double) * nc * 2));
        CU(cudaMalloc(&h->mom64_d, sizeof(double) * (size_t)nc * T * 2));
        const uint32_t bd64 = 256, bx64 = (g.plane + bd64 - 1) / bd64;
        CU(cudaMalloc(&h->part_d, sizeof(double2) * (size_t)nc * g.Tl * bx64 * 4));
    }
    CU(cudaMalloc(&h->chains_d, sizeof(ChainDev) * nc));
    CU(cudaMalloc(&h->thr_d, sizeof(uint32_t) * (size_t)nc * kThrStride));
    CU(cudaMalloc(&h->acc_d, sizeof(unsigned long long) * (size_t)nc * kAccSlots));
    CU(cudaMalloc(&h->obs_d, sizeof(unsigned long long) * (size_t)nc * 2));
    CU(cudaMalloc(&h->mom_d, sizeof(unsigned long long) * (size_t)nc * T * 2));
    CU(cudaMalloc(&h->shift_d, sizeof(int32_t) * nc));
    CU(cudaMalloc(&h->offset_d, sizeof(int32_t) * nc));
    CU(cudaMemsetAsync(h->offset_d, 0, sizeof(int32_t) * nc, h->stream));
    CU(cudaMalloc(&h->sweep_ctr_d, sizeof(uint64_t)));
    CU(cudaMalloc(&h->n_meas_d, sizeof(uint32_t)));
    CU(cudaMemsetAsync(h->acc_d, 0, sizeof(unsigned long long) * (size_t)nc * kAccSlots, h->stream));
    CU(cudaMemsetAsync(h->obs_d, 0, sizeof(unsigned long long) * (size_t)nc * 2, h->stream));
    CU(cudaMemsetAsync(h->sweep_ctr_d, 0, sizeof(uint64_t), h->stream));
    CU(cudaMemsetAsync(h->n_meas_d, 0, sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->mom_d, 0, sizeof(unsigned long long) * (size_t)nc * T * 2, h->stream));
    CU(cudaMalloc(&h->ticket_d, sizeof(unsigned int)));
    CU(cudaMemsetAsync(h->ticket_d, 0, sizeof(unsigned int), h->stream));
    CU(cudaHostAlloc(&h->result_h, sizeof(long long) * (size_t)nc * 2, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&h->result_d, h->result_h, 0));
    CU(cudaHostAlloc(&h->cres_h, sizeof(double) * T, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&h->cres_d, h->cres_h, 0));
    CU(cudaMalloc(&h->run_d, sizeof(uint64_t) * 2));
    CU(cudaMemsetAsync(h->run_d, 0, sizeof(uint64_t) * 2, h->stream));
    h->chains_h.assign(nc, ChainDev{});
    h->beta.assign(nc, 0.0);
    h->configured.assign(nc, 0);
    h->thr_h.assign((size_t)nc * kThrStride, 0u);
    h->diff_sum.assign(nc, nullptr);
    h->diff_comp.assign(nc, nullptr);
    h->diff_count.assign(nc, 0);

    const uint32_t no_narrow = LQFT_FLAG_NO_NARROW | LQFT_FLAG_FORCE_GENERIC | LQFT_FLAG_HALO_NCCL;
    h->gn = g;
    if (g.ghost) {
        // deep halo of the narrow copy: as many planes as the thinnest slab allows, at most kN16GhostDepth, even
        // The exchange volume per half-sweep does not depend on the depth; depth trades the exchange latency
        // (~20 us, amortised over `depth` half-sweeps) against redundant planes ((depth - 1) / Tl of the work).
        // Small planes are latency dominated, large ones are not (measured on 2 GPUs: 256^2 planes 8 > 4 > 2,
        // 1024^2 planes 2 > 4 > 8).
        const uint64_t plane_bytes = (uint64_t)g.plane * sizeof(uint16_t);
        uint32_t depth = plane_bytes <= (128u << 10) ? kN16GhostDepth : (plane_bytes <= (512u << 10) ? 4u : 2u);
        if (const char* e = getenv("LQFT_N16_DEPTH")) depth = (uint32_t)atoi(e);  // tuning knob
        for (uint32_t r = 0; r < ws; ++r) {
            uint32_t rt0 = 0, rnt = 0;
            ST(lqft_host_slab(T, ws, r, &rt0, &rnt));
            depth = std::min(depth, rnt);
            if (r == (cfg->rank + ws - 1) % ws) h->tl_lo = rnt;
            if (r == (cfg->rank + 1) % ws) h->tl_hi = rnt;
        }
        h->gn.ghost = std::max(2u, depth & ~1u);  // even: a period of `depth` half-sweeps starts with colour 0
        h->gn.chain_stride = (uint64_t)g.plane * (g.Tl + 4 * h->gn.ghost);  // two sets of ghost planes (kernels_narrow.cuh)
    }
    if (!h->is64 && !(cfg->flags & no_narrow) && narrow_geometry_ok(g, 2, nc)) {
        // 16-bit storage of the field; without it the i32 kernels do all the work
        const size_t nbytes = sizeof(uint16_t) * h->gn.chain_stride * nc;
        if (cudaMalloc(&h->narrow[0], nbytes) != cudaSuccess || cudaMalloc(&h->narrow[1], nbytes) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(h->narrow[0]);
            h->narrow[0] = h->narrow[1] = nullptr;
        } else {
            CU(cudaMalloc(&h->ncentre_d, sizeof(int32_t) * nc));
            CU(cudaMalloc(&h->nworst_d, sizeof(uint32_t)));
            CU(cudaMalloc(&h->phase_d, sizeof(unsigned long long)));
        }
    }
    // i32 arrays: always on slabs (the fallback of a slab whose heights leave the band, halos by ncclSend/Recv); on one
    // GPU only where the 16-bit copy cannot be the storage — otherwise need_i32 allocates them if a field ever
    // leaves the band
    if (!h->is64 && (g.ghost != 0 || !h->narrow[0])) ST(need_i32(h));
    if (ws > 1) {
        ST(nccl_load());
        ncclUniqueId id;
        memcpy(&id, cfg->nccl_id, 128);
        NC(g_nccl.CommInitRank(&h->comm, (int)ws, id, (int)cfg->rank));
        CU(cudaMalloc(&h->halo_d, sizeof(HaloSync)));
        CU(cudaMemsetAsync(h->halo_d, 0, sizeof(HaloSync), h->stream));
        ST(set_halo_timeout());
        ST(setup_p2p(h));
    } else if (self_halo) {
        // one rank that is its own lower and upper neighbour: the whole peer-memory halo path (pushes,
        // flags, phases) on a single device
        CU(cudaMalloc(&h->halo_d, sizeof(HaloSync)));
        CU(cudaMemsetAsync(h->halo_d, 0, sizeof(HaloSync), h->stream));
        for (int side = 0; side < 2; ++side) {
            h->peer_base[side][0] = h->field[0];
            h->peer_base[side][1] = h->field[1];
            h->peer_base[side][2] = h->halo_d;
            h->peer_base[side][3] = h->narrow[0];
            h->peer_base[side][4] = h->narrow[1];
        }
        h->self_halo = true;
        h->p2p = true;
    }
    if (!h->is64 && !h->field[0]) ST(narrow_zero(h));  // the field of zeros lives in the 16-bit arrays from the start
    CU(cudaStreamSynchronize(h->stream));
    return LQFT_OK;
}

extern "C" lqft_status lqft_create(const lqft_config* cfg, lqft_handle** out) {
    if (!cfg || !out) return fail(LQFT_ERR_INVALID, "null argument");
    *out = nullptr;
    lqft_handle* h = new (std::nothrow) lqft_handle();
    if (!h) return fail(LQFT_ERR_NOMEM, "out of host memory");
    lqft_status s = create_impl(cfg, h);
    if (s != LQFT_OK) {
        std::string keep = g_err;
        lqft_destroy(h);
        g_err = keep;
        return s;
    }
    *out = h;
    return LQFT_OK;
}

extern "C" lqft_status lqft_set_chain(lqft_handle* h, uint32_t chain, double beta, uint64_t seed,
                                      uint32_t wil_w, uint32_t wil_h) {
    if (!h) return fail(LQFT_ERR_INVALID, "null handle");
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    if (wil_w > h->cfg.x || wil_h > h->cfg.t)
        return fail(LQFT_ERR_INVALID, "wilson loop %ux%u does not fit the %ux%u x-t plane", wil_w, wil_h,
                    h->cfg.x, h->cfg.t);
    uint32_t* thr = h->thr_h.data() + (size_t)chain * kThrStride;
    std::fill(thr, thr + kThrStride, 0u);  // kernels may read up to the handle's longest table: the tail must be 0
    const uint32_t n = lqft_host_threshold_table(beta, thr, kThrStride);
    if (n == 0)
        return fail(LQFT_ERR_INVALID,
                    "beta=%g is not representable: need beta > 0 and at most %u threshold entries", beta,
                    kThrStride);
    ChainDev& cd = h->chains_h[chain];
    cd.k0 = (uint32_t)seed;
    cd.k1 = (uint32_t)(seed >> 32);
    cd.n_thr = n;
    cd.wil_w = (wil_w && wil_h) ? wil_w : 0;
    cd.wil_h = (wil_w && wil_h) ? wil_h : 0;
    h->beta[chain] = beta;
    h->configured[chain] = 1;
    h->params_dirty = true;
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// halo exchange (world_size > 1): colour-c boundary planes -> neighbours' ghost planes
// ---------------------------------------------------------------------------------------------
static lqft_status exchange_halo_nccl(lqft_handle* h, int colour, cudaStream_t s) {
    const Geom& g = h->g;
    if (h->cfg.world_size <= 1) {  // self halo: periodic wrap by copy
        for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
            int32_t* f = h->field[colour] + (size_t)c * g.chain_stride;
            CU(cudaMemcpyAsync(f + (size_t)(g.Tl + 1) * g.plane, f + (size_t)1 * g.plane, sizeof(int32_t) * g.plane, cudaMemcpyDeviceToDevice, s));
            CU(cudaMemcpyAsync(f, f + (size_t)g.Tl * g.plane, sizeof(int32_t) * g.plane, cudaMemcpyDeviceToDevice, s));
        }
        return LQFT_OK;
    }
    const int ws = (int)h->cfg.world_size, r = (int)h->cfg.rank;
    const int lo = (r + ws - 1) % ws, hi = (r + 1) % ws;
    NC(g_nccl.GroupStart());
    for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
        int32_t* f = h->field[colour] + (size_t)c * g.chain_stride;
        // my first owned plane is the lower neighbour's upper ghost; my last owned plane is the
        // upper neighbour's lower ghost.
        NC(g_nccl.Send(f + (size_t)1 * g.plane, g.plane, ncclInt32, lo, h->comm, s));
        NC(g_nccl.Send(f + (size_t)g.Tl * g.plane, g.plane, ncclInt32, hi, h->comm, s));
        NC(g_nccl.Recv(f + (size_t)(g.Tl + 1) * g.plane, g.plane, ncclInt32, hi, h->comm, s));
        NC(g_nccl.Recv(f, g.plane, ncclInt32, lo, h->comm, s));
    }
    NC(g_nccl.GroupEnd());
    return LQFT_OK;
}

static lqft_status refresh_ghosts(lqft_handle* h) {
    if (h->g.ghost == 0 || h->ghosts_valid) return LQFT_OK;
    ST(exchange_halo_nccl(h, 0, h->stream));
    ST(exchange_halo_nccl(h, 1, h->stream));
    h->ghosts_valid = true;
    return LQFT_OK;
}

static lqft_status allreduce_i64(lqft_handle* h, void* buf, size_t n) {
    if (h->cfg.world_size <= 1) return LQFT_OK;
    NC(g_nccl.AllReduce(buf, buf, n, ncclInt64, ncclSum, h->comm, h->stream));
    return LQFT_OK;
}
static lqft_status allreduce_i32(lqft_handle* h, void* buf, size_t n) {
    if (h->cfg.world_size <= 1) return LQFT_OK;
    NC(g_nccl.AllReduce(buf, buf, n, ncclInt32, ncclSum, h->comm, h->stream));
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
static lqft_status narrow_need_ghost(lqft_handle* h);

static inline uint32_t block_for(uint32_t units) {
    uint32_t b = (units + 31u) & ~31u;
    return std::min(256u, std::max(32u, b));
}

// Storage plane, in a neighbour's narrow array, of the first plane this rank sends into ghost set `set` there:
// the lower neighbour's upper ghosts / the upper neighbour's lower ghosts (layout: kernels_v1.cuh, plane_self).
static uint32_t peer_ghost_plane(const lqft_handle* h, int side, uint32_t set) {
    const uint32_t d = h->gn.ghost;
    if (side == 0) return set ? h->tl_lo + 3 * d : d + h->tl_lo;
    return set ? h->tl_hi + 2 * d : 0u;
}

// Flag words and number of the flip that makes the other ghost set current (eager: the host's count; inside a
// captured period: relative to the device copy the replay starts from).
static void flip_args(lqft_handle* h, SlabArgs& sl) {
    sl.sync = 1;
    sl.mine = h->halo_d;
    sl.lo_sync = (HaloSync*)h->peer_base[0][2];
    sl.hi_sync = (HaloSync*)h->peer_base[1][2];
    sl.seq_base = h->capturing ? h->phase_d : nullptr;
    sl.seq_off = h->capturing ? h->phase - h->phase_capture0 : h->phase;
}

// Narrow copy: one half-sweep of colour `colour` over local planes [first, first + n_planes).  Slabs: `flip` opens a
// period (the kernel's prologue swaps the ghost sets with the neighbours), `push` sends the boundary planes of this
// colour into the neighbours' other set.
static lqft_status launch_half_narrow(lqft_handle* h, int colour, uint64_t sweep, const uint64_t* ctr, int32_t first,
                                      uint32_t n_planes, bool flip = false, bool push = false, bool measure = false,
                                      uint32_t extra = 0) {
    NarrowArgs a{h->gn, colour, sweep, ctr, first, n_planes, 0, n_planes, h->cfg.n_chains, h->narrow[colour],
                 h->narrow[colour ^ 1], h->chains_d, h->chains_h[0], h->thr_d, h->acc_d, h->max_thr, h->any_wilson, h->count_accepts,
                 h->stream, !(h->cfg.flags & LQFT_FLAG_NO_PDL)};
    if (flip || push || extra) {
        a.slab = true;
        a.sl.extra = extra;  // ghost planes per side updated along with the owned ones (ghost tasks of k_sweep_n16)
        if (flip) flip_args(h, a.sl);
        if (push) {
            const uint32_t d = h->gn.ghost, next = h->gn.gbuf ^ 1u;
            a.sl.depth = d;
            a.sl.lo = (uint16_t*)h->peer_base[0][3 + colour];
            a.sl.hi = (uint16_t*)h->peer_base[1][3 + colour];
            a.sl.lo_stride8 = (uint32_t)((uint64_t)h->gn.plane * (h->tl_lo + 4 * d) / 8);
            a.sl.hi_stride8 = (uint32_t)((uint64_t)h->gn.plane * (h->tl_hi + 4 * d) / 8);
            a.sl.lo_plane = peer_ghost_plane(h, 0, next);
            a.sl.hi_plane = peer_ghost_plane(h, 1, next);
        }
    }
    if (measure) {
        a.obs_e = h->obs_d;
        a.fin = h->run_fin;  // where the running schedule logs its energies; the launch's last block does it
        a.fin.ticket = h->ticket_d;
    }
    CU(narrow_half(h->narrowp, a));
    h->launches += 1;
    return LQFT_OK;
}

// One half-sweep of colour `colour` over local planes tl_first .. tl_first + n_planes - 1.
static lqft_status launch_half(lqft_handle* h, int colour, uint64_t sweep, const uint64_t* ctr,
                               uint32_t tl_first, uint32_t n_planes, cudaStream_t s) {
    if (n_planes == 0) return LQFT_OK;
    const Geom& g = h->g;
    if (h->is64) {
        const uint32_t bd = block_for(g.plane);
        dim3 grid((g.plane + bd - 1) / bd, g.Tl, h->cfg.n_chains);
        if (h->any_wilson)
            k_sweep_generic_f64<true><<<grid, bd, 0, s>>>(g, colour, sweep, ctr, h->field64[colour], h->field64[colour ^ 1],
                                                          h->chains_d, h->beta_d, h->acc_d);
        else
            k_sweep_generic_f64<false><<<grid, bd, 0, s>>>(g, colour, sweep, ctr, h->field64[colour], h->field64[colour ^ 1],
                                                           h->chains_d, h->beta_d, h->acc_d);
        CU(cudaGetLastError());
        h->launches += 1;
        return LQFT_OK;
    }
    int32_t* own = h->field[colour];
    const int32_t* oth = h->field[colour ^ 1];
    const bool vec = (g.X % 8 == 0) && !(h->cfg.flags & LQFT_FLAG_FORCE_GENERIC);
    const uint32_t units = vec ? g.plane / 4 : g.plane;
    const uint32_t bd = block_for(units);
    dim3 grid((units + bd - 1) / bd, n_planes, h->cfg.n_chains);
    const size_t smem = sizeof(uint32_t) * h->max_thr;
#define LAUNCH(K)                                                                                    \
    K<<<grid, bd, smem, s>>>(g, colour, sweep, ctr, tl_first, own, oth, h->chains_d, h->thr_d, h->acc_d)
    if (vec) {
        if (h->any_wilson) LAUNCH(k_sweep_vec4<true>); else LAUNCH(k_sweep_vec4<false>);
    } else {
        if (h->any_wilson) LAUNCH(k_sweep_generic<true>); else LAUNCH(k_sweep_generic<false>);
    }
#undef LAUNCH
    CU(cudaGetLastError());
    h->launches += 1;
    return LQFT_OK;
}

// One full red-black sweep.  `sweep` + (*ctr if ctr) is the RNG sweep counter.

static lqft_status narrow_fill(lqft_handle* h);

// The energy of a step can ride on its colour-1 half-sweep (kernels_narrow.cuh, n16_obs_pair): one GPU, 16-bit storage.
static bool can_fuse_energy(const lqft_handle* h) {
    return h->narrow_active && h->gn.ghost == 0 && !h->count_accepts && !(h->cfg.flags & LQFT_FLAG_NO_FUSED_ENERGY);
}

static lqft_status launch_sweep(lqft_handle* h, uint64_t sweep, const uint64_t* ctr, bool fuse_energy = false) {
    const Geom& g = h->g;
    if (h->narrow_active) {
        // narrow copy.  Slabs: a period of `depth` half-sweeps per ghost set; each half-sweep also updates the ghost
        // planes that still have valid neighbours and the valid depth shrinks by one; the last two push this rank's
        // boundary planes into the neighbours' other set, which the next period flips to (kernels_narrow.cuh)
        const uint32_t depth = h->gn.ghost;
        for (int colour = 0; colour < 2; ++colour) {
            if (depth == 0) {
                ST(launch_half_narrow(h, colour, sweep, ctr, 0, g.Tl, false, false, fuse_energy && colour == 1));
                continue;
            }
            bool flip = false;
            if (h->nvalid == 0) {
                if (!h->next_ready) ST(narrow_fill(h));
                flip = true;
                h->phase += 1;
                h->gn.gbuf ^= 1u;
                h->nvalid = depth;
                h->next_ready = false;
            }
            const uint32_t extra = h->nvalid - 1;  // ghost planes per side this half-sweep may still update
            const bool push = h->nvalid <= 2;       // nvalid 2: colour 0 for the last time in this period, 1: colour 1
            h->nvalid -= 1;
            ST(launch_half_narrow(h, colour, sweep, ctr, 0, g.Tl, flip, push, false, extra));
            if (h->nvalid == 0) h->next_ready = true;
        }
        return LQFT_OK;
    }
    if (g.ghost == 0) {
        ST(launch_half(h, 0, sweep, ctr, 0, g.Tl, h->stream));
        ST(launch_half(h, 1, sweep, ctr, 0, g.Tl, h->stream));
        return LQFT_OK;
    }
    // slab, NCCL: boundary planes first, exchange on the comm stream while the interior updates
    for (int colour = 0; colour < 2; ++colour) {
        ST(launch_half(h, colour, sweep, ctr, 0, 1, h->stream));
        ST(launch_half(h, colour, sweep, ctr, g.Tl - 1, 1, h->stream));
        CU(cudaEventRecord(h->ev_bnd, h->stream));
        CU(cudaStreamWaitEvent(h->comm_stream, h->ev_bnd, 0));
        ST(exchange_halo_nccl(h, colour, h->comm_stream));
        CU(cudaEventRecord(h->ev_comm, h->comm_stream));
        ST(launch_half(h, colour, sweep, ctr, 1, g.Tl - 2, h->stream));
        CU(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
    }
    return LQFT_OK;
}

// f64 handles: compensated two-pass reduction, deterministic order
static lqft_status launch_observe_f64(lqft_handle* h, bool moments) {
    const Geom& g = h->g;
    const uint32_t bd = 256, bx = (g.plane + bd - 1) / bd;
    dim3 grid(bx, g.Tl, h->cfg.n_chains);
    k_observe_f64<<<grid, bd, 0, h->stream>>>(g, h->field64[0], h->field64[1], h->part_d);
    k_observe_fold_f64<<<h->cfg.n_chains, 128, 0, h->stream>>>(g, h->part_d, bx, h->obs64_d, moments ? h->mom64_d : nullptr);
    CU(cudaGetLastError());
    h->launches += 2;
    return LQFT_OK;
}

// One measurement of chains [chain0, chain0 + n) of an i32 handle (kernels_observe.cuh): energy and action sums,
// slice moments if asked, logged / folded on the device as `fin` and `due` say.
struct Measure {
    uint32_t due = 0;        // kObs* bits the finaliser logs
    bool moments = false;    // accumulate the slice moments
    bool fold = false;       // correlator epilogue (consumes the moments)
    uint64_t sweep = 0;      // RNG counter of the step (relative to *ctr when ctr is given)
    const uint64_t* ctr = nullptr;
    uint32_t chain0 = 0, n = 0;
    ObsFin fin{};
};

static lqft_status need_gathered(lqft_handle* h, uint32_t reps) {
    const size_t want = (size_t)reps * h->cfg.n_chains;
    if (want <= h->gathered_cap) return LQFT_OK;
    cudaFree(h->gathered_d);
    h->gathered_d = nullptr;
    h->gathered_cap = 0;
    CU(cudaMalloc(&h->gathered_d, sizeof(long long) * want));
    h->gathered_cap = want;
    return LQFT_OK;
}

template <typename H>
static lqft_status launch_measure_t(lqft_handle* h, const Geom& g, const H* f0, const H* f1, Measure m) {
    const uint32_t ws = h->cfg.world_size;
    const bool vec = g.X % 8 == 0 && !(h->cfg.flags & LQFT_FLAG_FORCE_GENERIC);
    m.fin.ticket = h->ticket_d;
    m.fin.fused = (vec && ws == 1) ? 1u : 0u;
    if (vec) {
        // a few blocks per SM, each a run of rows of one plane: at least one quad per thread
        const uint32_t Qx = g.Xh / 4;
        const uint32_t per = g.Tl * m.n;
        uint32_t ny = (8u * (uint32_t)h->sm_count + per - 1) / per;
        ny = std::min(ny, std::max(1u, g.Y * Qx / 256u));
        ny = std::max(1u, std::min(ny, g.Y));
        dim3 grid(ny, g.Tl, m.n);
        if (m.moments)
            k_observe_rows<H, true><<<grid, 256, 0, h->stream>>>(g, f0, f1, m.chain0, h->obs_d, h->mom_d, m.fin, m.due, m.sweep, m.ctr,
                                                                h->chains_d);
        else
            k_observe_rows<H, false><<<grid, 256, 0, h->stream>>>(g, f0, f1, m.chain0, h->obs_d, h->mom_d, m.fin, m.due, m.sweep, m.ctr,
                                                                 h->chains_d);
    } else if constexpr (std::is_same<H, int32_t>::value) {
        const uint32_t bd = block_for(g.plane);
        dim3 grid((g.plane + bd - 1) / bd, g.Tl, m.n);
        k_observe<<<grid, bd, 0, h->stream>>>(g, f0 + (size_t)m.chain0 * g.chain_stride, f1 + (size_t)m.chain0 * g.chain_stride,
                                              h->obs_d + (size_t)m.chain0 * 2,
                                              m.moments ? h->mom_d + (size_t)m.chain0 * g.T * 2 : nullptr);
    } else {
        return fail(LQFT_ERR_STATE, "the 16-bit copy has no scalar observe kernel");
    }
    CU(cudaGetLastError());
    h->launches += 1;
    if (m.fin.fused) return LQFT_OK;
    ST(allreduce_i64(h, h->obs_d + (size_t)m.chain0 * 2, (size_t)m.n * 2));
    if (m.moments) ST(allreduce_i64(h, h->mom_d + (size_t)m.chain0 * g.T * 2, (size_t)m.n * g.T * 2));
    k_obs_finalize<<<1, 256, 0, h->stream>>>(m.fin, m.due, m.sweep, m.ctr, m.chain0, m.n, h->obs_d);
    CU(cudaGetLastError());
    h->launches += 1;
    if (m.fold) {
        const long long* gathered = nullptr;
        if (ws > 1 && m.fin.reps) {
            if ((size_t)m.fin.reps * h->cfg.n_chains > h->gathered_cap) return fail(LQFT_ERR_STATE, "correlator scratch not allocated");
            k_corr_gather<H><<<dim3((m.fin.reps + 127) / 128, m.n), 128, 0, h->stream>>>(g, f0, f1, m.fin, m.sweep, m.ctr, m.chain0,
                                                                                        h->chains_d, h->gathered_d);
            CU(cudaGetLastError());
            h->launches += 1;
            ST(allreduce_i64(h, h->gathered_d + (size_t)m.chain0 * m.fin.reps, (size_t)m.n * m.fin.reps));
            gathered = h->gathered_d;
        }
        k_corr_fold<H><<<m.n, 128, 0, h->stream>>>(g, f0, f1, m.fin, m.sweep, m.ctr, m.chain0, h->chains_d, gathered, h->mom_d);
        CU(cudaGetLastError());
        h->launches += 1;
    }
    return LQFT_OK;
}

static lqft_status launch_measure(lqft_handle* h, const Measure& m) {
    if (h->narrow_active) {
        ST(narrow_need_ghost(h));  // the last owned plane looks one plane up
        return launch_measure_t<uint16_t>(h, h->gn, h->narrow[0], h->narrow[1], m);
    }
    ST(launch_measure_t<int32_t>(h, h->g, h->field[0], h->field[1], m));
    return LQFT_OK;
}

// DifferencePlotOutputData::update of one chain on whichever copy is live (bond differences need no offset).
static lqft_status launch_difference(lqft_handle* h, uint32_t chain) {
    const Geom& g = h->g;
    const uint32_t bd = block_for(g.plane);
    dim3 grid((g.plane + bd - 1) / bd, g.Tl, 1);
    if (h->narrow_active) {
        ST(narrow_need_ghost(h));
        k_difference_update_h<uint16_t><<<grid, bd, 0, h->stream>>>(h->gn, h->narrow[0] + (size_t)chain * h->gn.chain_stride,
                                                                   h->narrow[1] + (size_t)chain * h->gn.chain_stride,
                                                                   h->diff_sum[chain], h->diff_comp[chain]);
        CU(cudaGetLastError());
        h->launches += 1;
        return LQFT_OK;
    }
    if (h->is64)
        k_difference_update_f64<<<grid, bd, 0, h->stream>>>(g, h->field64[0] + (size_t)chain * g.chain_stride,
                                                            h->field64[1] + (size_t)chain * g.chain_stride,
                                                            h->diff_sum[chain], h->diff_comp[chain]);
    else
        k_difference_update_h<int32_t><<<grid, bd, 0, h->stream>>>(g, h->field[0] + (size_t)chain * g.chain_stride,
                                                                  h->field[1] + (size_t)chain * g.chain_stride,
                                                                  h->diff_sum[chain], h->diff_comp[chain]);
    CU(cudaGetLastError());
    h->launches += 1;
    return LQFT_OK;
}

static lqft_status need_difference(lqft_handle* h, uint32_t chain) {
    if (h->diff_sum[chain]) return LQFT_OK;
    const size_t bytes = sizeof(double) * 3 * h->n_slab;
    if (cudaMalloc(&h->diff_sum[chain], bytes) != cudaSuccess || cudaMalloc(&h->diff_comp[chain], bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(LQFT_ERR_NOMEM, "cannot allocate %zu bytes for the difference plot", 2 * bytes);
    }
    CU(cudaMemsetAsync(h->diff_sum[chain], 0, bytes, h->stream));
    CU(cudaMemsetAsync(h->diff_comp[chain], 0, bytes, h->stream));
    return LQFT_OK;
}

static lqft_status launch_normalize(lqft_handle* h, uint64_t counter, const uint64_t* ctr,
                                    int64_t explicit_site, int32_t only_chain) {
    // normalize_random without a pass over the field: offset[chain] := stored height of the picked site
    // (kernels_v1.cuh, k_pick_shift).  Reads owned planes only, so it is not a ghost operation.
    const Geom& g = h->g;
    const uint32_t nc = h->cfg.n_chains;
    if (h->is64) {
        k_pick_offset_f64<<<(nc + 127) / 128, 128, 0, h->stream>>>(g, h->field64[0], h->field64[1], h->chains_d, nc, counter,
                                                                  ctr, explicit_site, only_chain, h->offset64_d);
        CU(cudaGetLastError());
        h->launches += 1;
        return LQFT_OK;
    }
    if (h->narrow_active)
        k_pick_shift<<<(nc + 127) / 128, 128, 0, h->stream>>>(h->gn, (const uint16_t*)h->narrow[0], (const uint16_t*)h->narrow[1],
                                                             h->chains_d, nc, counter, ctr, explicit_site, only_chain, h->shift_d,
                                                             (const int32_t*)h->ncentre_d, (int32_t)kN16Bias);
    else
        k_pick_shift<<<(nc + 127) / 128, 128, 0, h->stream>>>(g, (const int32_t*)h->field[0], (const int32_t*)h->field[1],
                                                             h->chains_d, nc, counter, ctr, explicit_site, only_chain, h->shift_d);
    CU(cudaGetLastError());
    ST(allreduce_i32(h, h->shift_d, nc));
    k_set_offset<<<(nc + 127) / 128, 128, 0, h->stream>>>(h->shift_d, h->offset_d, nc, only_chain);
    CU(cudaGetLastError());
    h->launches += 2;
    return LQFT_OK;
}

static lqft_status ready(lqft_handle* h) {
    if (!h) return fail(LQFT_ERR_INVALID, "null handle");
    ST(use_device(h));
    ST(upload_params(h));
    if (!h->narrow_active && !h->is64) {
        ST(need_i32(h));
        ST(refresh_ghosts(h));
    }
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// field I/O
// ---------------------------------------------------------------------------------------------
extern "C" lqft_status lqft_init_hot(lqft_handle* h) {
    if (!h) return fail(LQFT_ERR_INVALID, "null handle");
    ST(use_device(h));
    ST(upload_params(h));
    const Geom& g = h->g;
    const uint32_t bd = block_for(g.plane);
    dim3 grid((g.plane + bd - 1) / bd, g.Tl, h->cfg.n_chains);
    CU(cudaMemsetAsync(h->offset_d, 0, sizeof(int32_t) * h->cfg.n_chains, h->stream));
    if (h->is64) {
        CU(cudaMemsetAsync(h->offset64_d, 0, sizeof(double) * h->cfg.n_chains, h->stream));
        k_init_hot_f64<<<grid, bd, 0, h->stream>>>(g, h->field64[0], h->field64[1], h->chains_d);
    } else if (h->narrowp.enabled) {
        // straight into the 16-bit arrays: hot-start heights lie in [-128, 127], centre 0, no measurement needed
        CU(cudaMemsetAsync(h->ncentre_d, 0, sizeof(int32_t) * h->cfg.n_chains, h->stream));
        k_init_hot<uint16_t><<<grid, bd, 0, h->stream>>>(h->gn, h->narrow[0], h->narrow[1], h->chains_d, (int32_t)kN16Bias);
        h->narrow_active = true;
        h->narrow_budget = (uint32_t)kN16Bias - 1u - 128u;
        h->nvalid = 0;
        h->next_ready = false;
        h->narrow_segments += 1;
    } else {
        ST(need_i32(h));
        k_init_hot<int32_t><<<grid, bd, 0, h->stream>>>(g, h->field[0], h->field[1], h->chains_d, 0);
        h->narrow_active = false;
    }
    CU(cudaGetLastError());
    h->launches += 1;
    h->ghosts_valid = false;
    return LQFT_OK;
}

static lqft_status need_staging(lqft_handle* h) {
    if (!h->staging_d) {
        if (cudaMalloc(&h->staging_d, (h->is64 ? sizeof(double) : sizeof(int32_t)) * h->n_slab) != cudaSuccess) {
            cudaGetLastError();
            return fail(LQFT_ERR_NOMEM, "cannot allocate the %llu-byte staging buffer",
                        (unsigned long long)(sizeof(int32_t) * h->n_slab));
        }
    }
    return LQFT_OK;
}

extern "C" lqft_status lqft_upload(lqft_handle* h, uint32_t chain, const void* values) {
    if (!h || !values) return fail(LQFT_ERR_INVALID, "null argument");
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    ST(use_device(h));
    ST(need_staging(h));
    const Geom& g = h->g;
    const size_t esz = h->is64 ? sizeof(double) : sizeof(int32_t);
    CU(cudaMemcpyAsync(h->staging_d, values, esz * h->n_slab, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->offset_d + chain, 0, sizeof(int32_t), h->stream));
    const uint32_t bd = block_for(g.plane);
    dim3 grid((g.plane + bd - 1) / bd, g.Tl, 1);
    if (h->is64) {
        CU(cudaMemsetAsync(h->offset64_d + chain, 0, sizeof(double), h->stream));
        k_split_f64<<<grid, bd, 0, h->stream>>>(g, (const double*)h->staging_d, h->field64[0] + (size_t)chain * g.chain_stride,
                                                h->field64[1] + (size_t)chain * g.chain_stride);
    } else {
        bool to_i32 = !h->narrow_active;
        if (h->narrow_active && g.ghost != 0) {
            // slabs: measuring the band is a collective; an upload is not one.  Back to i32; the next run re-enters.
            ST(narrow_exit(h));
            to_i32 = true;
        }
        if (!to_i32) {
            // straight into the 16-bit arrays around centre 0, if the chain fits the band
            CU(cudaMemsetAsync(h->ncentre_d + chain, 0, sizeof(int32_t), h->stream));
            CU(cudaMemsetAsync(h->nworst_d, 0, sizeof(uint32_t), h->stream));
            k_split<uint16_t><<<grid, bd, 0, h->stream>>>(h->gn, h->staging_d, h->narrow[0] + (size_t)chain * h->gn.chain_stride,
                                                         h->narrow[1] + (size_t)chain * h->gn.chain_stride, h->ncentre_d + chain,
                                                         (int32_t)kN16Bias, h->nworst_d);
            CU(cudaGetLastError());
            h->launches += 1;
            uint32_t worst = 0;
            CU(cudaMemcpyAsync(&worst, h->nworst_d, sizeof worst, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (worst > (uint32_t)kN16Guard) {
                ST(narrow_exit(h));  // every chain back to i32 (this one's 16-bit image is meaningless and is replaced below)
                to_i32 = true;
            } else {
                h->narrow_budget = std::min(h->narrow_budget, (uint32_t)kN16Bias - 1u - worst);
                h->narrow_segments += 1;
            }
        }
        if (to_i32) {
            ST(need_i32(h));
            k_split<int32_t><<<grid, bd, 0, h->stream>>>(g, h->staging_d, h->field[0] + (size_t)chain * g.chain_stride,
                                                        h->field[1] + (size_t)chain * g.chain_stride, nullptr, 0, nullptr);
            CU(cudaGetLastError());
            h->launches += 1;
            h->ghosts_valid = false;
        }
    }
    if (h->is64) {
        CU(cudaGetLastError());
        h->launches += 1;
    }
    CU(cudaStreamSynchronize(h->stream));  // the host buffer may be reused on return
    return LQFT_OK;
}

extern "C" lqft_status lqft_download(lqft_handle* h, uint32_t chain, void* values) {
    if (!h || !values) return fail(LQFT_ERR_INVALID, "null argument");
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    ST(use_device(h));
    ST(need_staging(h));
    const Geom& g = h->g;
    const uint32_t bd = block_for(g.plane);
    dim3 grid((g.plane + bd - 1) / bd, g.Tl, 1);
    if (h->is64)
        k_merge_f64<<<grid, bd, 0, h->stream>>>(g, (double*)h->staging_d, h->field64[0] + (size_t)chain * g.chain_stride,
                                                h->field64[1] + (size_t)chain * g.chain_stride, h->offset64_d + chain);
    else if (h->narrow_active)
        k_merge<uint16_t><<<grid, bd, 0, h->stream>>>(h->gn, h->staging_d, h->narrow[0] + (size_t)chain * h->gn.chain_stride,
                                                     h->narrow[1] + (size_t)chain * h->gn.chain_stride, h->offset_d + chain,
                                                     h->ncentre_d + chain, (int32_t)kN16Bias);
    else {
        ST(need_i32(h));
        k_merge<int32_t><<<grid, bd, 0, h->stream>>>(g, h->staging_d, h->field[0] + (size_t)chain * g.chain_stride,
                                                    h->field[1] + (size_t)chain * g.chain_stride, h->offset_d + chain, nullptr, 0);
    }
    CU(cudaGetLastError());
    h->launches += 1;
    CU(cudaMemcpyAsync(values, h->staging_d, (h->is64 ? sizeof(double) : sizeof(int32_t)) * h->n_slab, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return check_halo_error(h);  // a lost neighbour must not pass for a valid field
}

// ---------------------------------------------------------------------------------------------
// update
// ---------------------------------------------------------------------------------------------
static lqft_status read_accepted(lqft_handle* h, uint64_t* accepted) {
    const uint32_t nc = h->cfg.n_chains;
    std::vector<unsigned long long> slots((size_t)nc * kAccSlots);
    CU(cudaMemcpyAsync(slots.data(), h->acc_d, sizeof(unsigned long long) * slots.size(),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemsetAsync(h->acc_d, 0, sizeof(unsigned long long) * slots.size(), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (uint32_t c = 0; c < nc; ++c) {
        uint64_t tot = 0;
        for (int s = 0; s < kAccSlots; ++s) tot += slots[(size_t)c * kAccSlots + s];
        accepted[c] = tot;
    }
    return LQFT_OK;
}

static lqft_status run_impl(lqft_handle* h, const lqft_schedule* s, const lqft_run_outputs* out);

extern "C" lqft_status lqft_sweep(lqft_handle* h, uint64_t first_sweep, uint32_t n_sweeps, uint64_t* accepted) {
    lqft_schedule s{};
    s.sweep_base = first_sweep;
    s.first_step = 0;
    s.n_steps = n_sweeps;
    lqft_run_outputs o{};
    o.accepted = accepted;
    return run_impl(h, &s, &o);
}

extern "C" lqft_status lqft_normalize(lqft_handle* h, uint64_t sweep_index) {
    ST(ready(h));
    return launch_normalize(h, sweep_index, nullptr, -1, -1);
}

extern "C" lqft_status lqft_normalize_site(lqft_handle* h, uint32_t chain, uint64_t site) {
    ST(ready(h));
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    if (site >= h->n_global) return fail(LQFT_ERR_INVALID, "site %llu outside the lattice", (unsigned long long)site);
    return launch_normalize(h, 0, nullptr, (int64_t)site, (int32_t)chain);
}

extern "C" lqft_status lqft_shift_values(lqft_handle* h, uint32_t chain, double shift) {
    ST(ready(h));
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    if (h->is64) {
        k_add_offset_f64<<<1, 1, 0, h->stream>>>(h->offset64_d, chain, shift);
        CU(cudaGetLastError());
        h->launches += 1;
        return LQFT_OK;
    }
    const int32_t s = (int32_t)shift;
    if ((double)s != shift) return fail(LQFT_ERR_INVALID, "shift %g is not an i32 value", shift);
    k_add_offset<<<1, 1, 0, h->stream>>>(h->offset_d, chain, s);  // true = stored - offset
    CU(cudaGetLastError());
    h->launches += 1;
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// observables
// ---------------------------------------------------------------------------------------------
// Energy and action sums of every chain: one launch, one copy, one synchronisation.
static lqft_status measure_sums(lqft_handle* h, std::vector<long long>& obs) {
    const uint32_t nc = h->cfg.n_chains;
    Measure m;
    m.n = nc;
    m.fin.result = h->result_d;
    ST(launch_measure(h, m));
    CU(cudaStreamSynchronize(h->stream));
    obs.assign(h->result_h, h->result_h + (size_t)nc * 2);
    return h->g.ghost ? check_halo_error(h) : LQFT_OK;
}

// f64 handles: energy / action (and moments) as doubles
static lqft_status measure64(lqft_handle* h, bool moments, std::vector<double>& obs) {
    const uint32_t nc = h->cfg.n_chains;
    ST(launch_observe_f64(h, moments));
    obs.resize((size_t)nc * 2);
    CU(cudaMemcpyAsync(obs.data(), h->obs64_d, sizeof(double) * obs.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LQFT_OK;
}

extern "C" lqft_status lqft_energy(lqft_handle* h, int64_t* e_int, double* e_obs) {
    ST(ready(h));
    if (h->is64) {  // float_energy of a Field<f64>; e_int is its truncation
        std::vector<double> o;
        ST(measure64(h, false, o));
        for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
            if (e_int) e_int[c] = (int64_t)o[(size_t)c * 2];
            if (e_obs) e_obs[c] = h->beta[c] * o[(size_t)c * 2];
        }
        return LQFT_OK;
    }
    std::vector<long long> obs;
    ST(measure_sums(h, obs));
    for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
        if (e_int) e_int[c] = obs[(size_t)c * 2];
        if (e_obs) e_obs[c] = h->beta[c] * (double)obs[(size_t)c * 2];  // fields.rs:669-671
    }
    return LQFT_OK;
}

extern "C" lqft_status lqft_action(lqft_handle* h, int64_t* s_int, double* s_obs) {
    ST(ready(h));
    if (h->is64) {
        std::vector<double> o;
        ST(measure64(h, false, o));
        for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
            if (s_int) s_int[c] = (int64_t)o[(size_t)c * 2 + 1];
            if (s_obs) s_obs[c] = h->beta[c] * o[(size_t)c * 2 + 1];
        }
        return LQFT_OK;
    }
    std::vector<long long> obs;
    ST(measure_sums(h, obs));
    for (uint32_t c = 0; c < h->cfg.n_chains; ++c) {
        if (s_int) s_int[c] = obs[(size_t)c * 2 + 1];
        if (s_obs) s_obs[c] = h->beta[c] * (double)obs[(size_t)c * 2 + 1];  // fields.rs:665-667
    }
    return LQFT_OK;
}

extern "C" lqft_status lqft_slice_moments(lqft_handle* h, uint32_t chain, int64_t* s1, int64_t* s2) {
    ST(ready(h));
    if (h->is64) return fail(LQFT_ERR_INVALID, "integer slice moments are defined for LQFT_I32 fields");
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    const uint32_t T = h->g.T;
    Measure m;
    m.moments = true;
    m.chain0 = chain;
    m.n = 1;
    m.fin.result = h->result_d;
    ST(launch_measure(h, m));
    std::vector<long long> mom((size_t)T * 2);
    int32_t off = 0;
    CU(cudaMemcpyAsync(mom.data(), h->mom_d + (size_t)chain * T * 2, sizeof(long long) * T * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemsetAsync(h->mom_d + (size_t)chain * T * 2, 0, sizeof(long long) * T * 2, h->stream));  // moments are zero between uses
    CU(cudaMemcpyAsync(&off, h->offset_d + chain, sizeof off, cudaMemcpyDeviceToHost, h->stream));
    int32_t centre = 0;
    if (h->narrow_active) CU(cudaMemcpyAsync(&centre, h->ncentre_d + chain, sizeof centre, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->narrow_active) off -= centre - (int32_t)kN16Bias;  // 16-bit copy: stored = height - centre + bias
    if (h->g.ghost) ST(check_halo_error(h));
    // moments of the TRUE heights h = stored - off:  sum h = S1 - A off,  sum h^2 = S2 - 2 off S1 + A off^2
    const long long A = (long long)h->g.X * h->g.Y, o = off;
    for (uint32_t t = 0; t < T; ++t) {
        const long long m1 = mom[(size_t)t * 2], m2 = mom[(size_t)t * 2 + 1];
        if (s1) s1[t] = m1 - A * o;
        if (s2) s2[t] = m2 - 2 * o * m1 + A * o * o;
    }
    return LQFT_OK;
}

static lqft_status need_sites(lqft_handle* h, uint32_t reps) {
    if (reps <= h->sites_cap) return LQFT_OK;
    cudaFree(h->sites_d);
    cudaFree(h->gathered64_d);
    h->sites_d = nullptr;
    h->gathered64_d = nullptr;
    h->sites_cap = 0;
    CU(cudaMalloc(&h->sites_d, sizeof(uint64_t) * reps));
    if (h->is64) CU(cudaMalloc(&h->gathered64_d, sizeof(double) * reps));
    h->sites_cap = reps;
    return LQFT_OK;
}

extern "C" lqft_status lqft_correlator(lqft_handle* h, uint32_t chain, const uint64_t* ref_sites,
                                       uint32_t reps, uint64_t sweep_index, double* c_of_tau) {
    ST(ready(h));
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    if (!c_of_tau) return fail(LQFT_ERR_INVALID, "null output");
    const Geom& g = h->g;
    const uint32_t T = g.T;
    if (ref_sites)
        for (uint32_t r = 0; r < reps; ++r)
            if (ref_sites[r] >= h->n_global)
                return fail(LQFT_ERR_INVALID, "reference site %llu outside the lattice", (unsigned long long)ref_sites[r]);
    if (h->is64) {
        std::vector<uint64_t> sites(reps);
        const ChainDev& cd = h->chains_h[chain];
        for (uint32_t r = 0; r < reps; ++r)
            sites[r] = ref_sites ? ref_sites[r] : pick_site(cd.k0, cd.k1, LQFT_STREAM_CORRELATOR, sweep_index, r, h->n_global);
        std::vector<double> o;
        ST(measure64(h, true, o));
        std::vector<double> m((size_t)T * 2), hv(reps, 0.0);
        CU(cudaMemcpyAsync(m.data(), h->mom64_d + (size_t)chain * T * 2, sizeof(double) * T * 2, cudaMemcpyDeviceToHost, h->stream));
        if (reps) {
            ST(need_sites(h, reps));
            CU(cudaMemcpyAsync(h->sites_d, sites.data(), sizeof(uint64_t) * reps, cudaMemcpyHostToDevice, h->stream));
            k_gather_sites_f64<<<(reps + 127) / 128, 128, 0, h->stream>>>(g, h->field64[0] + (size_t)chain * g.chain_stride,
                                                                         h->field64[1] + (size_t)chain * g.chain_stride,
                                                                         h->sites_d, reps, h->gathered64_d);
            CU(cudaGetLastError());
            h->launches += 1;
            CU(cudaMemcpyAsync(hv.data(), h->gathered64_d, sizeof(double) * reps, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        const long double A = (long double)g.X * g.Y;
        std::vector<long double> c(T, 0.0L);
        for (uint32_t r = 0; r < reps; ++r) {
            const uint32_t tx = (uint32_t)(sites[r] / ((uint64_t)g.X * g.Y));
            const long double v = hv[r];
            for (uint32_t tau = 0; tau < T; ++tau) {
                const uint32_t t = (tx + tau) % T;
                c[tau] += A * v * v - 2.0L * v * (long double)m[(size_t)t * 2] + (long double)m[(size_t)t * 2 + 1];
            }
        }
        for (uint32_t tau = 0; tau < T; ++tau) c_of_tau[tau] = (double)c[tau];
        return LQFT_OK;
    }
    // i32: moments of this chain only, reference heights and the O(reps * T) epilogue on the device (kernels_observe.cuh)
    Measure m;
    m.due = kObsCorr;
    m.moments = m.fold = true;
    m.sweep = sweep_index;
    m.chain0 = chain;
    m.n = 1;
    m.fin.clog = h->cres_d;
    m.fin.reps = reps;
    if (ref_sites && reps) {
        ST(need_sites(h, reps));
        CU(cudaMemcpyAsync(h->sites_d, ref_sites, sizeof(uint64_t) * reps, cudaMemcpyHostToDevice, h->stream));
        m.fin.sites = h->sites_d;
    }
    if (h->cfg.world_size > 1) ST(need_gathered(h, reps));
    ST(launch_measure(h, m));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(c_of_tau, h->cres_h, sizeof(double) * T);
    return h->g.ghost ? check_halo_error(h) : LQFT_OK;
}

extern "C" lqft_status lqft_difference_update(lqft_handle* h, uint32_t chain) {
    ST(ready(h));
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    ST(need_difference(h, chain));
    ST(launch_difference(h, chain));
    h->diff_count[chain] += 1;
    return LQFT_OK;
}

extern "C" lqft_status lqft_difference_read(lqft_handle* h, uint32_t chain, double* mean, uint64_t* count) {
    if (!h) return fail(LQFT_ERR_INVALID, "null handle");
    ST(use_device(h));
    if (chain >= h->cfg.n_chains) return fail(LQFT_ERR_INVALID, "chain %u outside [0,%u)", chain, h->cfg.n_chains);
    const uint64_t n = h->diff_count[chain];
    if (count) *count = n;
    if (!mean) return LQFT_OK;
    const size_t elems = 3 * h->n_slab;
    if (!h->diff_sum[chain]) {
        for (size_t i = 0; i < elems; ++i) mean[i] = 0.0;  // WelfordsAlgorithm64::get_mean with 0 entries
        return LQFT_OK;
    }
    CU(cudaMemcpyAsync(mean, h->diff_sum[chain], sizeof(double) * elems, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < elems; ++i) mean[i] = mean[i] / (double)n;  // kahan.rs:90-95
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// 16-bit storage (kernels_narrow.cuh)
// ---------------------------------------------------------------------------------------------
// The 16-bit arrays are the storage of the field wherever they can be (also between calls); the i32 arrays take over
// only while some height lies outside the band.  Bookkeeping: a sweep moves a height by at most one, so after a
// conversion that measured `worst` = max |height - centre| the copy is good for kN16Bias - 1 - worst sweeps
// (narrow_budget) without anyone looking at it; when they are used up the band is measured again (narrow_recheck) —
// the only host synchronisation of the update path, once per > 1000 sweeps.
static lqft_status allreduce_max_u32(lqft_handle* h, void* buf, size_t n) {
    if (h->cfg.world_size <= 1) return LQFT_OK;
    NC(g_nccl.AllReduce(buf, buf, n, ncclUint32, ncclMax, h->comm, h->stream));
    return LQFT_OK;
}

static lqft_status read_worst(lqft_handle* h, uint32_t* worst) {
    ST(allreduce_max_u32(h, h->nworst_d, 1));
    CU(cudaMemcpyAsync(worst, h->nworst_d, sizeof *worst, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LQFT_OK;
}

// i32 field -> 16-bit copy biased around each chain's normalisation offset; *ok = every height of every rank lies
// within kN16Guard of its centre.
static lqft_status narrow_enter(lqft_handle* h, bool* ok) {
    const Geom& g = h->g;
    const uint32_t nc = h->cfg.n_chains;
    *ok = false;
    k_narrow_begin<<<(nc + 127) / 128, 128, 0, h->stream>>>(h->offset_d, h->ncentre_d, nc, h->nworst_d);
    const uint64_t count = (uint64_t)g.Tl * g.plane;  // owned planes; ghost planes are filled by the first period
    dim3 grid((uint32_t)((count / 4 + 255) / 256), nc);
    for (int c = 0; c < 2; ++c)
        k_narrow<<<grid, 256, 0, h->stream>>>(h->field[c] + (size_t)g.ghost * g.plane, g.chain_stride,
                                              h->narrow[c] + (size_t)h->gn.ghost * g.plane, h->gn.chain_stride, count,
                                              h->ncentre_d, h->nworst_d);
    CU(cudaGetLastError());
    h->launches += 3;
    uint32_t worst = 0;
    ST(read_worst(h, &worst));
    *ok = worst <= (uint32_t)kN16Guard;
    h->narrow_active = *ok;
    h->nvalid = 0;            // neither ghost set holds anything yet: the first half-sweep fills and flips
    h->next_ready = false;
    if (*ok) {
        h->narrow_budget = (uint32_t)kN16Bias - 1u - worst;
        h->narrow_segments += 1;
    } else {
        h->narrow_retry = kN16Segment;
    }
    return LQFT_OK;
}

// 16-bit copy -> i32 arrays, which become the live copy.
static lqft_status narrow_exit(lqft_handle* h) {
    const Geom& g = h->g;
    ST(need_i32(h));
    const uint64_t count = (uint64_t)g.Tl * g.plane;
    dim3 grid((uint32_t)((count / 4 + 255) / 256), h->cfg.n_chains);
    for (int c = 0; c < 2; ++c)
        k_widen<<<grid, 256, 0, h->stream>>>(h->narrow[c] + (size_t)h->gn.ghost * g.plane, h->gn.chain_stride,
                                             h->field[c] + (size_t)g.ghost * g.plane, g.chain_stride, count, h->ncentre_d);
    CU(cudaGetLastError());
    h->launches += 2;
    h->narrow_active = false;
    if (g.ghost) h->ghosts_valid = false;  // the i32 ghost planes were not maintained meanwhile
    return LQFT_OK;
}

// The budget is used up: measure the heights against the chains' current offsets; within the guard band they are
// re-centred there in place and the copy lives on, otherwise the field moves to the i32 arrays.
static lqft_status narrow_recheck(lqft_handle* h) {
    const Geom& g = h->g;
    const uint32_t nc = h->cfg.n_chains;
    const uint64_t count = (uint64_t)g.Tl * g.plane;
    dim3 grid((uint32_t)((count / 4 + 255) / 256), nc);
    CU(cudaMemsetAsync(h->nworst_d, 0, sizeof(uint32_t), h->stream));
    for (int c = 0; c < 2; ++c)
        k_n16_worst<<<grid, 256, 0, h->stream>>>(h->narrow[c] + (size_t)h->gn.ghost * g.plane, h->gn.chain_stride, count,
                                                 h->ncentre_d, h->offset_d, h->nworst_d);
    CU(cudaGetLastError());
    h->launches += 2;
    uint32_t worst = 0;
    ST(read_worst(h, &worst));
    if (worst > (uint32_t)kN16Guard) {
        h->narrow_retry = kN16Segment;
        return narrow_exit(h);
    }
    for (int c = 0; c < 2; ++c)
        k_n16_rebias<<<grid, 256, 0, h->stream>>>(h->narrow[c] + (size_t)h->gn.ghost * g.plane, h->gn.chain_stride, count,
                                                  h->ncentre_d, h->offset_d);
    k_narrow_begin<<<(nc + 127) / 128, 128, 0, h->stream>>>(h->offset_d, h->ncentre_d, nc, h->nworst_d);
    CU(cudaGetLastError());
    h->launches += 3;
    h->narrow_budget = (uint32_t)kN16Bias - 1u - worst;
    h->nvalid = 0;  // the ghost planes were not moved along: fill and flip
    h->next_ready = false;
    h->narrow_segments += 1;
    return LQFT_OK;
}

// First fill of the neighbours' other ghost set with this rank's boundary planes (k_n16_fill): after the copy was
// built, or when a period is cut short.  The flip that follows tells the neighbours.
static lqft_status narrow_fill(lqft_handle* h) {
    const Geom& gn = h->gn;
    const uint32_t next = gn.gbuf ^ 1u;
    FillArgs a{};
    for (int c = 0; c < 2; ++c) {
        a.mine[c] = h->narrow[c];
        a.lo[c] = (uint16_t*)h->peer_base[0][3 + c];
        a.hi[c] = (uint16_t*)h->peer_base[1][3 + c];
    }
    a.stride_mine = gn.chain_stride;
    a.stride_lo = (uint64_t)gn.plane * (h->tl_lo + 4 * gn.ghost);
    a.stride_hi = (uint64_t)gn.plane * (h->tl_hi + 4 * gn.ghost);
    a.plane = gn.plane;
    a.depth = gn.ghost;
    a.tl_mine = gn.Tl;
    a.lo_plane = peer_ghost_plane(h, 0, next);
    a.hi_plane = peer_ghost_plane(h, 1, next);
    k_n16_fill<<<dim3(kN16FillBlocks, 4, h->cfg.n_chains), 256, 0, h->stream>>>(a);
    CU(cudaGetLastError());
    h->launches += 1;
    h->next_ready = true;
    return LQFT_OK;
}

// Operations other than a half-sweep that look at a ghost plane of the narrow copy (the observables: the last owned
// plane looks one plane up): if the current set is used up, flip to the other one.
static lqft_status narrow_need_ghost(lqft_handle* h) {
    if (!h->narrow_active || h->gn.ghost == 0 || h->nvalid != 0) return LQFT_OK;
    if (!h->next_ready) ST(narrow_fill(h));
    h->phase += 1;
    SlabArgs sl{};
    flip_args(h, sl);
    k_n16_flip<<<1, 32, 0, h->stream>>>(sl);
    CU(cudaGetLastError());
    h->launches += 1;
    h->gn.gbuf ^= 1u;
    h->nvalid = h->gn.ghost;
    h->next_ready = false;
    return LQFT_OK;
}

// ---------------------------------------------------------------------------------------------
// fused schedule (computation.rs:229-270, 313-340)
// ---------------------------------------------------------------------------------------------
static lqft_status launch_energy_log_f64(lqft_handle* h) {
    ST(launch_observe_f64(h, false));
    k_energy_log_f64<<<1, 128, 0, h->stream>>>(h->obs64_d, h->elog64_d, h->elog_cap, h->cfg.n_chains, h->n_meas_d);
    CU(cudaGetLastError());
    h->launches += 1;
    return LQFT_OK;
}

// One step of the schedule: sweep, then every observable that is due (computation.rs:262-265), then normalize if
// due (computation.rs:267-269).
static lqft_status launch_step(lqft_handle* h, const lqft_schedule* s, uint64_t step_arg,
                               uint64_t step_for_gating, const uint64_t* ctr) {
    uint32_t due = 0;
    if (s->energy_every && step_for_gating % s->energy_every == 0) due |= kObsEnergy;
    if (s->action_every && step_for_gating % s->action_every == 0) due |= kObsAction;
    if (s->correlator_every && step_for_gating % s->correlator_every == 0) due |= kObsCorr;
    const bool fused = due == kObsEnergy && can_fuse_energy(h);
    ST(launch_sweep(h, s->sweep_base + step_arg, ctr, fused));
    if (fused) {
        // the colour-1 half-sweep has accumulated the energies and its last block has logged them
    } else if (h->is64) {
        if (due & kObsEnergy) ST(launch_energy_log_f64(h));
    } else if (due) {
        Measure m;
        m.due = due;
        m.moments = m.fold = (due & kObsCorr) != 0;
        m.sweep = s->sweep_base + step_arg;
        m.ctr = ctr;
        m.n = h->cfg.n_chains;
        m.fin = h->run_fin;
        ST(launch_measure(h, m));
    }
    if (s->difference_every && step_for_gating % s->difference_every == 0)
        for (uint32_t c = 0; c < h->cfg.n_chains; ++c) ST(launch_difference(h, c));
    if (s->normalize_every && step_for_gating % s->normalize_every == 0)
        ST(launch_normalize(h, s->sweep_base + step_arg, ctr, -1, -1));
    return LQFT_OK;
}

static uint64_t gcd64(uint64_t a, uint64_t b) { return b ? gcd64(b, a % b) : a; }
static uint64_t lcm64(uint64_t a, uint64_t b) { return a / gcd64(a, b) * b; }

extern "C" lqft_status lqft_run(lqft_handle* h, const lqft_schedule* s, double* e_obs, int64_t* e_int,
                                uint64_t* accepted) {
    lqft_run_outputs o{};
    o.e_obs = e_obs;
    o.e_int = e_int;
    o.accepted = accepted;
    return run_impl(h, s, &o);
}

extern "C" lqft_status lqft_run_observables(lqft_handle* h, const lqft_schedule* s, const lqft_run_outputs* out) {
    lqft_run_outputs none{};
    return run_impl(h, s, out ? out : &none);
}

template <typename T>
static lqft_status grow_log(lqft_handle* h, T*& log, uint32_t& cap, uint64_t n_meas, size_t per_meas, const char* what) {
    if (n_meas <= cap) return LQFT_OK;
    cudaFree(log);
    log = nullptr;
    cap = 0;
    if (cudaMalloc(&log, sizeof(T) * (size_t)h->cfg.n_chains * n_meas * per_meas) != cudaSuccess) {
        cudaGetLastError();
        return fail(LQFT_ERR_NOMEM, "cannot allocate the %s log (%llu measurements)", what, (unsigned long long)n_meas);
    }
    cap = (uint32_t)n_meas;
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();  // graphs bake the log pointers and capacities
    return LQFT_OK;
}

static lqft_status run_impl(lqft_handle* h, const lqft_schedule* s, const lqft_run_outputs* out) {
    if (!s) return fail(LQFT_ERR_INVALID, "null schedule");
    ST(ready(h));
    uint64_t* accepted = out->accepted;
    h->count_accepts = accepted != nullptr;
    const uint32_t nc = h->cfg.n_chains;
    const uint64_t n_meas = lqft_run_measurements(s->first_step, s->n_steps, s->energy_every);
    const uint64_t n_act = lqft_run_measurements(s->first_step, s->n_steps, s->action_every);
    const uint64_t n_corr = lqft_run_measurements(s->first_step, s->n_steps, s->correlator_every);
    const uint64_t n_diff = lqft_run_measurements(s->first_step, s->n_steps, s->difference_every);
    if (n_meas > 0xffffffffull || n_act > 0xffffffffull || n_corr > 0xffffffffull)
        return fail(LQFT_ERR_INVALID, "too many measurements in one run");
    if (h->is64 && (s->action_every || s->correlator_every || s->difference_every))
        return fail(LQFT_ERR_INVALID, "LQFT_F64 handles log the energy only inside lqft_run; use lqft_action / lqft_correlator / "
                                      "lqft_difference_update between calls");
    if (n_meas > h->elog_cap && h->is64) {
        cudaFree(h->elog64_d);
        h->elog64_d = nullptr;
        CU(cudaMalloc(&h->elog64_d, sizeof(double) * (size_t)nc * n_meas));
    }
    ST(grow_log(h, h->elog_d, h->elog_cap, n_meas, 1, "energy"));
    ST(grow_log(h, h->alog_d, h->alog_cap, n_act, 1, "action"));
    ST(grow_log(h, h->clog_d, h->clog_cap, n_corr, h->g.T, "correlator"));
    if (n_corr && h->cfg.world_size > 1) ST(need_gathered(h, s->correlator_reps));
    if (n_diff)
        for (uint32_t c = 0; c < nc; ++c) ST(need_difference(h, c));
    {
        ObsFin& f = h->run_fin;
        f = ObsFin{};
        f.elog = h->elog_d;
        f.alog = h->alog_d;
        f.clog = h->clog_d;
        f.cap_e = h->elog_cap;
        f.cap_a = h->alog_cap;
        f.cap_c = h->clog_cap;
        f.every_e = s->energy_every;
        f.every_a = s->action_every;
        f.every_c = s->correlator_every;
        f.reps = s->correlator_reps;
        f.run = h->run_d;
        const uint64_t run[2] = {s->sweep_base, s->first_step};
        CU(cudaMemcpyAsync(h->run_d, run, sizeof run, cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaMemsetAsync(h->n_meas_d, 0, sizeof(uint32_t), h->stream));
    if (accepted) CU(cudaMemsetAsync(h->acc_d, 0, sizeof(unsigned long long) * (size_t)nc * kAccSlots, h->stream));
    // the cluster-resident kernel adds every CTA's share into the log slots
    const bool cluster_ok = h->clusterp.enabled && !s->correlator_every && !s->difference_every;
    if (cluster_ok && n_meas) CU(cudaMemsetAsync(h->elog_d, 0, sizeof(long long) * (size_t)nc * h->elog_cap, h->stream));
    if (cluster_ok && n_act) CU(cudaMemsetAsync(h->alog_d, 0, sizeof(long long) * (size_t)nc * h->alog_cap, h->stream));

    uint64_t cycle = 1;
    for (uint32_t f : {s->energy_every, s->normalize_every, s->action_every, s->correlator_every, s->difference_every})
        if (f) cycle = std::min<uint64_t>(lcm64(cycle, f), 1u << 20);
    // one captured graph = `reps` cycles (about kGraphSteps sweeps); it is replayed while whole periods remain
    constexpr uint64_t kGraphSteps = 32;
    const bool graph_ok = !(h->cfg.flags & LQFT_FLAG_NO_GRAPH) && cycle <= 64 && s->n_steps >= 2 * cycle && s->n_steps >= 8;
    const bool extra_obs = s->action_every || s->correlator_every || s->difference_every;

    CU(cudaEventRecord(h->ev_start, h->stream));
    uint64_t step = s->first_step;
    const uint64_t end = s->first_step + s->n_steps;
    if (h->resident_ok && !extra_obs && s->n_steps > 0) {
        // small lattice: the whole schedule in one launch, one CTA per chain, field in shared memory
        ResidentArgs a{h->g, s->sweep_base, s->first_step, s->n_steps, s->energy_every, s->normalize_every,
                       h->elog_cap, h->count_accepts ? 1u : 0u};
        const size_t smem = resident_smem_bytes(h->g, h->max_thr);
        const int vec = h->g.Xh % 4 == 0 ? 4 : (h->g.Xh % 2 == 0 ? 2 : 1);
#define LQFT_RES(W, V)                                                                                                 \
    k_run_resident<W, V><<<nc, kResidentThreads, smem, h->stream>>>(a, h->field[0], h->field[1], h->chains_d, h->thr_d, \
                                                                    h->offset_d, h->elog_d, h->acc_d)
        if (h->any_wilson) {
            if (vec == 4) LQFT_RES(true, 4); else if (vec == 2) LQFT_RES(true, 2); else LQFT_RES(true, 1);
        } else {
            if (vec == 4) LQFT_RES(false, 4); else if (vec == 2) LQFT_RES(false, 2); else LQFT_RES(false, 1);
        }
#undef LQFT_RES
        CU(cudaGetLastError());
        h->launches += 1;
        step = end;
    }
    // Steps left for the update kernels.  Where the 16-bit arrays are admitted they are (or become) the live copy;
    // a segment lasts as long as their budget does, then the band is measured again.  Outside the band the call
    // goes on with the i32 kernels.
    while (step < end) {
        uint64_t seg_end = end;
        if (h->narrowp.enabled && !h->narrow_active && h->narrow_retry == 0) {
            bool ok = false;
            ST(narrow_enter(h, &ok));
        }
        if (h->narrow_active && h->narrow_budget == 0) ST(narrow_recheck(h));
        if (h->narrow_active) {
            seg_end = std::min<uint64_t>(end, step + h->narrow_budget);
        } else {
            if (h->narrowp.enabled) seg_end = std::min<uint64_t>(end, step + h->narrow_retry);
            ST(refresh_ghosts(h));  // leaving the 16-bit copy leaves the i32 ghost planes stale
        }
        const uint64_t seg_begin = step;
        if (h->narrow_active && cluster_ok && seg_end - step >= 2) {
            // the whole segment in one launch, the chains resident in shared memory (kernels_cluster.cuh)
            ClusterArgs a{};
            a.g = h->gn;
            a.sweep_base = s->sweep_base;
            a.first_step = step;
            a.n_steps = seg_end - step;
            a.energy_every = s->energy_every;
            a.action_every = s->action_every;
            a.normalize_every = s->normalize_every;
            a.elog_cap = h->elog_cap;
            a.alog_cap = h->alog_cap;
            a.e_slot0 = (uint32_t)lqft_run_measurements(s->first_step, step - s->first_step, s->energy_every);
            a.a_slot0 = (uint32_t)lqft_run_measurements(s->first_step, step - s->first_step, s->action_every);
            a.n_cta = h->clusterp.n_cta;
            a.tp = h->clusterp.tp;
            a.max_thr = h->max_thr;
            CU(cluster_launch(h->clusterp, a, nc, h->any_wilson, h->count_accepts, h->narrow[0], h->narrow[1], h->chains_d, h->thr_d,
                              h->offset_d, h->ncentre_d, h->elog_d, h->alog_d, h->acc_d, h->stream));
            h->launches += 1;
            step = seg_end;
        }
        // slabs replay graphs only on the narrow copy: the numbers of its flips live in device memory
        const bool use_graph = graph_ok && (h->g.ghost == 0 || h->narrow_active);
        const bool deep = h->narrow_active && h->gn.ghost != 0;
        if (use_graph && step < seg_end) {
            while (step < seg_end && step % cycle != 0) {
                ST(launch_step(h, s, step, step, nullptr));
                ++step;
            }
            const uint64_t avail = (seg_end - step) / cycle;
            uint64_t reps = std::max<uint64_t>(1, std::min<uint64_t>(avail, std::max<uint64_t>(1, kGraphSteps / cycle)));
            if (deep) {
                // a replay must leave the ghost sets as it found them: whole periods of `depth` half-sweeps, and an
                // even number of them (every period flips the set): the graph period is a multiple of `depth` sweeps
                const uint64_t m = h->gn.ghost / gcd64(cycle, h->gn.ghost);
                reps = std::max<uint64_t>(m, (reps + m - 1) / m * m);
                if (reps > avail) reps = avail / m * m;
            }
            const uint64_t period = reps * cycle;
            const uint64_t n_periods = period ? (seg_end - step) / period : 0;
            if (n_periods) {
                if (deep && !(h->nvalid == 0 && h->next_ready)) {
                    // entry state of a replay: the current set used up, the other one filled
                    ST(narrow_fill(h));
                    h->nvalid = 0;
                }
                GraphKey key{s->energy_every, s->normalize_every, s->action_every, s->correlator_every, s->correlator_reps,
                             s->difference_every, (uint32_t)period, h->count_accepts, h->narrow_active, deep ? h->gn.gbuf : 0u};
                auto it = h->graphs.find(key);
                if (it == h->graphs.end()) {
                    // capture one period; every kernel reads the sweep counter (and, on slabs, the number of the
                    // flip it performs) from device memory
                    const unsigned long long phase_before = h->phase;
                    const uint64_t before = h->launches;
                    lqft_schedule rel = *s;
                    rel.sweep_base = 0;
                    cudaGraph_t graph = nullptr;
                    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                    h->capturing = true;
                    h->phase_capture0 = phase_before;
                    lqft_status st = LQFT_OK;
                    for (uint64_t i = 0; i < period && st == LQFT_OK; ++i) st = launch_step(h, &rel, i, i, h->sweep_ctr_d);
                    if (st == LQFT_OK) {
                        k_tick<<<1, 1, 0, h->stream>>>(h->sweep_ctr_d, period, h->phase_d, h->phase - phase_before);
                        h->launches += 1;
                    }
                    h->capturing = false;
                    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                    if (st != LQFT_OK) {
                        if (graph) cudaGraphDestroy(graph);
                        return st;
                    }
                    if (ce != cudaSuccess) return fail(LQFT_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
                    if (deep && !(h->nvalid == 0 && h->next_ready && h->gn.gbuf == key.gbuf))
                        return fail(LQFT_ERR_STATE, "internal: a captured period must end where it began (ghost sets)");
                    GraphVal val;
                    val.kernels = h->launches - before;
                    val.phases = h->phase - phase_before;
                    h->launches = before;
                    h->phase = phase_before;
                    ce = cudaGraphInstantiate(&val.exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (ce != cudaSuccess) return fail(LQFT_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
                    it = h->graphs.emplace(key, val).first;
                }
                const uint64_t base = s->sweep_base + step;
                CU(cudaMemcpyAsync(h->sweep_ctr_d, &base, sizeof base, cudaMemcpyHostToDevice, h->stream));
                if (deep) CU(cudaMemcpyAsync(h->phase_d, &h->phase, sizeof h->phase, cudaMemcpyHostToDevice, h->stream));
                for (uint64_t c = 0; c < n_periods; ++c) CU(cudaGraphLaunch(it->second.exec, h->stream));
                h->launches += it->second.kernels * n_periods;
                h->phase += it->second.phases * n_periods;
                step += n_periods * period;
            }
        }
        while (step < seg_end) {
            ST(launch_step(h, s, step, step, nullptr));
            ++step;
        }
        if (h->narrow_active) h->narrow_budget -= (uint32_t)(step - seg_begin);
        else if (h->narrowp.enabled) h->narrow_retry -= (uint32_t)std::min<uint64_t>(h->narrow_retry, step - seg_begin);
    }
    CU(cudaEventRecord(h->ev_stop, h->stream));
    h->timed = true;
    for (uint32_t c = 0; c < nc && n_diff; ++c) h->diff_count[c] += n_diff;

    // results: one copy per log at the end of the call
    bool synced = false;
    if (n_meas && (out->e_obs || out->e_int) && h->is64) {
        std::vector<double> log((size_t)nc * h->elog_cap);
        CU(cudaMemcpyAsync(log.data(), h->elog64_d, sizeof(double) * log.size(), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        synced = true;
        for (uint32_t c = 0; c < nc; ++c)
            for (uint64_t m = 0; m < n_meas; ++m) {
                const double e = log[(size_t)c * h->elog_cap + m];
                if (out->e_int) out->e_int[(size_t)c * n_meas + m] = (int64_t)e;
                if (out->e_obs) out->e_obs[(size_t)c * n_meas + m] = h->beta[c] * e;
            }
    } else {
        std::vector<long long> elog, alog;
        if (n_meas && (out->e_obs || out->e_int)) {
            elog.resize((size_t)nc * h->elog_cap);
            CU(cudaMemcpyAsync(elog.data(), h->elog_d, sizeof(long long) * elog.size(), cudaMemcpyDeviceToHost, h->stream));
        }
        if (n_act && (out->s_obs || out->s_int)) {
            alog.resize((size_t)nc * h->alog_cap);
            CU(cudaMemcpyAsync(alog.data(), h->alog_d, sizeof(long long) * alog.size(), cudaMemcpyDeviceToHost, h->stream));
        }
        if (n_corr && out->corr)  // rows of T doubles, chain-major like the other logs
            for (uint32_t c = 0; c < nc; ++c)
                CU(cudaMemcpyAsync(out->corr + (size_t)c * n_corr * h->g.T, h->clog_d + (size_t)c * h->clog_cap * h->g.T,
                                   sizeof(double) * n_corr * h->g.T, cudaMemcpyDeviceToHost, h->stream));
        if (!elog.empty() || !alog.empty() || (n_corr && out->corr)) {
            CU(cudaStreamSynchronize(h->stream));
            synced = true;
        }
        for (uint32_t c = 0; c < nc && !elog.empty(); ++c)
            for (uint64_t m = 0; m < n_meas; ++m) {
                const long long e = elog[(size_t)c * h->elog_cap + m];
                if (out->e_int) out->e_int[(size_t)c * n_meas + m] = e;
                if (out->e_obs) out->e_obs[(size_t)c * n_meas + m] = h->beta[c] * (double)e;  // fields.rs:669-671
            }
        for (uint32_t c = 0; c < nc && !alog.empty(); ++c)
            for (uint64_t m = 0; m < n_act; ++m) {
                const long long a = alog[(size_t)c * h->alog_cap + m];
                if (out->s_int) out->s_int[(size_t)c * n_act + m] = a;
                if (out->s_obs) out->s_obs[(size_t)c * n_act + m] = h->beta[c] * (double)a;  // fields.rs:665-667
            }
    }
    if (accepted) {
        ST(read_accepted(h, accepted));
        synced = true;
    }
    // slabs: a call that handed results back has synchronised; say so if a neighbour was lost on the way
    if (h->g.ghost && synced) ST(check_halo_error(h));
    return LQFT_OK;
}

extern "C" lqft_status lqft_synchronize(lqft_handle* h) {
    if (!h) return fail(LQFT_ERR_INVALID, "null handle");
    ST(use_device(h));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->comm_stream));
    return check_halo_error(h);
}

extern "C" lqft_status lqft_last_device_ms(lqft_handle* h, float* ms) {
    if (!h || !ms) return fail(LQFT_ERR_INVALID, "null argument");
    if (!h->timed) return fail(LQFT_ERR_STATE, "no timed call yet");
    ST(use_device(h));
    CU(cudaEventSynchronize(h->ev_stop));
    CU(cudaEventElapsedTime(ms, h->ev_start, h->ev_stop));
    return LQFT_OK;
}

extern "C" lqft_status lqft_kernel_launches(lqft_handle* h, uint64_t* n) {
    if (!h || !n) return fail(LQFT_ERR_INVALID, "null argument");
    *n = h->launches;
    return LQFT_OK;
}

extern "C" lqft_status lqft_narrow_segments(lqft_handle* h, uint64_t* n) {
    if (!h || !n) return fail(LQFT_ERR_INVALID, "null argument");
    *n = h->narrow_segments;
    return LQFT_OK;
}

extern "C" lqft_status lqft_plan_info(lqft_handle* h, char* buf, size_t cap) {
    if (!h || !buf || cap == 0) return fail(LQFT_ERR_INVALID, "null argument");
    ST(use_device(h));
    ST(upload_params(h));
    const char* path = h->is64 ? "f64" : h->resident_ok ? "resident" : h->clusterp.enabled ? "cluster" : h->narrowp.enabled ? "narrow"
                       : ((h->g.X % 8 == 0 && !(h->cfg.flags & LQFT_FLAG_FORCE_GENERIC)) ? "vec4" : "generic");
    uint32_t chunks = 0, rows = 0, groups = 0;
    int occ = 0;
    if (h->narrowp.enabled) {
        chunks = narrow_chunks(h->narrowp, h->gn, h->g.Tl, h->cfg.n_chains, h->any_wilson, false);
        rows = (h->g.Y + chunks - 1) / chunks;
        groups = narrow_groups(h->narrowp, h->g.Tl);
        occ = narrow_occupancy(h->narrowp, h->any_wilson, false, false, narrow_warps(h->narrowp, h->g.Tl));
    }
    snprintf(buf, cap, "path=%s chunks=%u rows_per_chunk=%u plane_blocks=%u blocks_per_sm=%d ghost_depth=%u p2p=%d storage=%s "
             "i32_allocated=%d budget=%u", path, chunks, rows, groups, occ, h->narrowp.enabled ? h->gn.ghost : h->g.ghost,
             h->p2p ? 1 : 0, h->is64 ? "f64" : h->narrow_active ? "u16" : "i32", h->field[0] ? 1 : 0,
             h->narrow_active ? h->narrow_budget : 0u);
    return LQFT_OK;
}

extern "C" lqft_status lqft_stream(lqft_handle* h, void** stream) {
    if (!h || !stream) return fail(LQFT_ERR_INVALID, "null argument");
    *stream = (void*)h->stream;
    return LQFT_OK;
}

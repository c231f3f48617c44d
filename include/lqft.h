/*
 * lqft.h — C ABI of liblqft, the B200 (sm_100a) Monte-Carlo engine for the 3-D
 * discrete-Gaussian height model of TimeTeleporter/lattice_qft.
 *
 * This header is the drop-in boundary for the reference's Metropolis + observables
 * hot path.  The reference has no FFI of its own; every entry point below names the
 * Rust item (file:line under the reference checkout) whose work it takes over.  The
 * Rust-side `extern "C"` block a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every call returns lqft_status (0 = ok); the message of the last failure on the
 *     calling thread is lqft_last_error().
 *   - handles are opaque; one handle = one device, one CUDA stream, `n_chains`
 *     independent Markov chains on lattices of identical extents (sim.rs:90-93 runs
 *     one chain per rayon worker; here one chain per batch slot).
 *   - all state is device resident.  Host buffers cross the boundary only in
 *     lqft_upload / lqft_download and in the small observable outputs.
 *   - host buffers holding a field are in the REFERENCE's lexicographic order
 *     idx = x + X*(y + Y*t) (lattice.rs:93-104); the colour-split device layout is
 *     private to the library.
 *   - extents must be even (red-black ordering) and >= 2.
 *   - there is NO CPU fallback: without a usable CUDA device lqft_create fails with
 *     LQFT_ERR_CUDA.  The `lqft_host_*` functions are pure host arithmetic (RNG,
 *     threshold table, geometry) and need no device; they define the decision spec
 *     that the device kernels and the test oracle share.
 */
#ifndef LQFT_H
#define LQFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQFT_VERSION_MAJOR 0
#define LQFT_VERSION_MINOR 2

typedef enum lqft_status {
    LQFT_OK = 0,
    LQFT_ERR_INVALID = 1, /* bad argument (odd extent, beta <= 0, chain out of range ...) */
    LQFT_ERR_CUDA = 2,    /* CUDA runtime error, or no device: there is no CPU fallback    */
    LQFT_ERR_NCCL = 3,    /* NCCL missing or failed (only when world_size > 1)            */
    LQFT_ERR_STATE = 4,   /* call sequence error (chain not configured ...)               */
    LQFT_ERR_NOMEM = 5
} lqft_status;

typedef enum lqft_dtype {
    LQFT_I32 = 0, /* Field<i32,...>: the type of every driver path (computation.rs:226,308) */
    LQFT_F64 = 1  /* Field<f64,...>: admitted by HeightVariable (fields.rs:573-599).  One GPU, scalar
                     kernels; upload/download take double; integer outputs (e_int, s_int) hold the
                     truncated f64 value; lqft_slice_moments is not defined; hot start = i8 values */
} lqft_dtype;

/* RNG stream tags: second counter word of Philox4x32-10 (see lqft_host_philox4x32). */
#define LQFT_STREAM_COLOUR0 0u   /* Metropolis draws, even sites (x+y+t even) */
#define LQFT_STREAM_COLOUR1 1u   /* Metropolis draws, odd sites               */
#define LQFT_STREAM_NORMALIZE 2u /* site pick of normalize_random             */
#define LQFT_STREAM_HOTSTART 3u  /* i8-uniform hot start                      */
#define LQFT_STREAM_CORRELATOR 4u/* reference sites of the correlator         */

#define LQFT_MAX_THRESHOLDS 2048u

typedef struct lqft_config {
    uint32_t x, y, t;     /* GLOBAL lattice extents; Lattice3d<X,Y,T> (lattice.rs:127-141) */
    uint32_t dtype;       /* lqft_dtype */
    uint32_t n_chains;    /* independent chains in this handle (>= 1) */
    int32_t device;       /* CUDA device ordinal */
    uint32_t world_size;  /* ranks sharing ONE lattice as t-slabs; 1 = whole lattice here  */
    uint32_t rank;        /* this rank's slab index: owns t in [rank*T/ws, (rank+1)*T/ws)  */
    uint32_t flags;       /* LQFT_FLAG_* */
    uint32_t reserved;
    uint8_t nccl_id[128]; /* ncclUniqueId from lqft_comm_unique_id (rank 0), world_size > 1 */
} lqft_config;

#define LQFT_FLAG_NONE 0u
#define LQFT_FLAG_NO_GRAPH 1u     /* launch sweeps eagerly instead of through CUDA graphs   */
#define LQFT_FLAG_FORCE_GENERIC 2u/* always use the scalar kernels (test cross-check)       */
#define LQFT_FLAG_HALO_NCCL 4u    /* world_size>1: i32 field with ncclSend/Recv halos (default: 16-bit storage
                                     whose boundary planes travel over peer memory, CUDA IPC) */
/* 8u and 16u selected the round-1 i32 marching / streaming kernels; those are gone, the bits are ignored */
#define LQFT_FLAG_NO_PDL 32u      /* no programmatic dependent launch between half-sweeps     */
#define LQFT_FLAG_NO_RESIDENT 128u /* multi-launch path instead of the one-launch kernels that keep a chain in shared
                                     memory (one CTA up to 36^3, a thread-block cluster up to ~1.5 MB per chain) */
#define LQFT_FLAG_NO_NARROW 256u   /* powers of two 16 <= X <= 1024: keep the field in the i32 arrays instead of the 16-bit
                                     storage (same results, used by the tests to cross-check) */
#define LQFT_FLAG_NO_FUSED_ENERGY 512u /* measure a due energy with the observe kernel instead of inside the colour-1
                                     half-sweep of that step (same results, used by the tests to cross-check) */
#define LQFT_FLAG_SELF_HALO 64u   /* world_size==1: ghost planes + the peer-memory halo path with
                                     this rank as its own neighbour (tests the multi-GPU kernel) */

typedef struct lqft_handle lqft_handle;

/* ---- library ------------------------------------------------------------------- */
uint32_t lqft_version(void);
const char* lqft_last_error(void);
/* Number of CUDA devices visible (0 when there is none; never fails). */
int32_t lqft_device_count(void);
/* Fill id[128] with a fresh ncclUniqueId (rank 0 calls it, the caller broadcasts it). */
lqft_status lqft_comm_unique_id(uint8_t id[128]);

/* ---- lifetime ------------------------------------------------------------------ */
/* Replaces Lattice::new + Field::new (lattice.rs:65-80, fields.rs:32-41): allocates the
 * device fields (zero-initialised).  No neighbour table is built: neighbours are computed. */
lqft_status lqft_create(const lqft_config* cfg, lqft_handle** out);
lqft_status lqft_destroy(lqft_handle* h);

/* Per-chain parameters.  `beta` is the reference's `temp` (algorithm.rs:127-132); `seed`
 * keys the chain's Philox streams (the reference is unseeded, computation.rs:222).
 * wilson_width/height > 0 selects WilsonField::from_field(field, width, height)
 * (fields.rs:550-567): +y bonds at y==0 && x<width && t<height are shifted by +1. */
lqft_status lqft_set_chain(lqft_handle* h, uint32_t chain, double beta, uint64_t seed,
                           uint32_t wilson_width, uint32_t wilson_height);

/* Hot start of ALL chains: Field::<i8>::random -> Field::<i32>::from_field
 * (computation.rs:225-226, fields.rs:43-61): heights uniform in [-128,127], drawn from
 * stream LQFT_STREAM_HOTSTART of each chain's seed. */
lqft_status lqft_init_hot(lqft_handle* h);

/* Field::from_values (fields.rs:63-67) / read-back of Field.values.  `values` holds this
 * rank's slab (X*Y*T/world_size elements of the handle's dtype) in lexicographic order. */
lqft_status lqft_upload(lqft_handle* h, uint32_t chain, const void* values);
lqft_status lqft_download(lqft_handle* h, uint32_t chain, void* values);

/* ---- update -------------------------------------------------------------------- */
/* n_sweeps red-black Metropolis sweeps of every chain; sweep s uses RNG counter
 * first_sweep + s.  Replaces Metropolis::field_sweep / wilson_sweep (algorithm.rs:115-123,
 * 165-174) — per-site logic of metropolis_single (algorithm.rs:127-158) unchanged, site
 * order red-black instead of lexicographic.  accepted[n_chains] (nullable) receives the
 * number of accepted proposals summed over the n_sweeps (the `usize` field_sweep returns),
 * for this rank's slab.  Asynchronous unless `accepted` is given — with one exception: where the 16-bit arrays store the
 * field (powers of two 16 <= X <= 1024), the heights are measured against the band once the sweep budget of the last
 * measurement is used up (at most once per 1024 sweeps), and that measurement is read by the host. */
lqft_status lqft_sweep(lqft_handle* h, uint64_t first_sweep, uint32_t n_sweeps,
                       uint64_t* accepted);

/* HeightField::normalize_random (fields.rs:621-625) for all chains: subtract the height of
 * one uniformly picked site; the pick comes from stream LQFT_STREAM_NORMALIZE at counter
 * `sweep_index` (lqft_host_pick_site).  lqft_normalize_site takes the site explicitly
 * (global lexicographic index) = shift_values(values[site]) (fields.rs:628-632). */
lqft_status lqft_normalize(lqft_handle* h, uint64_t sweep_index);
lqft_status lqft_normalize_site(lqft_handle* h, uint32_t chain, uint64_t site);
/* HeightField::shift_values (fields.rs:628-632) with an explicit shift (i32 handles). */
lqft_status lqft_shift_values(lqft_handle* h, uint32_t chain, double shift);

/* ---- observables (all chains at once; outputs are host arrays of n_chains) -------- */
/* Action::integer_energy / energy_observable (fields.rs:728-737, 669-671):
 * e_int = sum_sites[(d+x)^2 + (d+y)^2 - (d+t)^2] exactly in int64 (the reference wraps in
 * i32, quirk Q5: its value is the low 32 bits of e_int); e_obs = beta * e_int.
 * world_size > 1: summed over ranks (every rank receives the total). */
lqft_status lqft_energy(lqft_handle* h, int64_t* e_int, double* e_obs);
/* Action::float_action / action_observable (fields.rs:719-726, 665-667). */
lqft_status lqft_action(lqft_handle* h, int64_t* s_int, double* s_obs);
/* Per-time-slice moments S1[t] = sum_{x,y} h, S2[t] = sum h^2 (global T entries each). */
lqft_status lqft_slice_moments(lqft_handle* h, uint32_t chain, int64_t* s1, int64_t* s2);
/* CorrelationPlotOutputData::update (outputdata.rs:313-344): c[tau], tau < T, summed over
 * `reps` reference sites.  ref_sites == NULL draws them from stream LQFT_STREAM_CORRELATOR
 * at counter `sweep_index` (lqft_host_pick_site), else uses the given global indices. */
lqft_status lqft_correlator(lqft_handle* h, uint32_t chain, const uint64_t* ref_sites,
                            uint32_t reps, uint64_t sweep_index, double* c_of_tau);
/* DifferencePlotOutputData::update / plot (outputdata.rs:241-276): running per-bond mean of
 * bond_difference for the 3 positive directions.  _update accumulates the current
 * configuration of `chain`; _read returns mean[3][X*Y*T/world_size] (direction-major,
 * lexicographic) and the number of accumulated configurations. */
lqft_status lqft_difference_update(lqft_handle* h, uint32_t chain);
lqft_status lqft_difference_read(lqft_handle* h, uint32_t chain, double* mean, uint64_t* count);

/* ---- fused measurement schedule -------------------------------------------------- */
/* The loop bodies of Simulation::run / WilsonSim::run (computation.rs:229-270, 313-340) for
 * steps [first_step, first_step + n_steps): per step one sweep (RNG counter sweep_base +
 * step), then every observable whose frequency divides the step (the OutputData filter of
 * computation.rs:262-265, set_frequency outputdata.rs:99-107), then if step % normalize_every
 * == 0 normalize_random (computation.rs:267-269).  A frequency of 0 switches that item off.
 *   energy_every      EnergyObservable::update        (outputdata.rs:168-178)
 *   action_every      ActionObservableNew::update     (outputdata.rs:202-212)
 *   correlator_every  CorrelationPlotOutputData::update with `correlator_reps` reference sites
 *                     (outputdata.rs:313-344); the sites of step s come from stream
 *                     LQFT_STREAM_CORRELATOR at counter sweep_base + s, rep r (lqft_host_pick_site) —
 *                     the same sites lqft_correlator(h, chain, NULL, reps, sweep_base + s, ..) draws
 *   difference_every  DifferencePlotOutputData::update of every chain (outputdata.rs:241-256);
 *                     read with lqft_difference_read
 * Measurements are taken of every chain.  Everything runs on the device without host round
 * trips (one CUDA graph per period of the schedule); the logs are copied back once at the end. */
typedef struct lqft_schedule {
    uint64_t sweep_base;    /* RNG sweep counter of step 0 (burn-in offset)              */
    uint64_t first_step;
    uint64_t n_steps;
    uint32_t energy_every;
    uint32_t normalize_every; /* 10 in the reference (computation.rs:242,267)            */
    uint32_t action_every;
    uint32_t correlator_every;
    uint32_t correlator_reps; /* 100 in src/bin/sim.rs:78                                */
    uint32_t difference_every;
} lqft_schedule;
/* Where lqft_run_observables puts its logs; every pointer is nullable.  n_x =
 * lqft_run_measurements(first_step, n_steps, x_every); all arrays are chain-major:
 * e_obs[chain * n_energy + m], corr[(chain * n_corr + m) * T + tau]. */
typedef struct lqft_run_outputs {
    double* e_obs;   int64_t* e_int;   /* energy_observable = beta * integer_energy (fields.rs:669-671) */
    double* s_obs;   int64_t* s_int;   /* action_observable = beta * action (fields.rs:665-667)         */
    double* corr;                      /* T values per measurement (one row of correlation_{i}.csv)     */
    uint64_t* accepted;                /* [n_chains] accepted proposals over the call (this rank's slab) */
} lqft_run_outputs;
uint64_t lqft_run_measurements(uint64_t first_step, uint64_t n_steps, uint32_t every);
lqft_status lqft_run_observables(lqft_handle* h, const lqft_schedule* s, const lqft_run_outputs* out);
/* lqft_run_observables with the energy log and the acceptance counts only. */
lqft_status lqft_run(lqft_handle* h, const lqft_schedule* s, double* e_obs /*nullable*/,
                     int64_t* e_int /*nullable*/, uint64_t* accepted /*nullable*/);

lqft_status lqft_synchronize(lqft_handle* h);
/* CUDA-event time in milliseconds spent by the device in the last lqft_sweep / lqft_run
 * call on this handle (measured on the handle's stream). */
lqft_status lqft_last_device_ms(lqft_handle* h, float* ms);
/* Number of kernels this handle has launched so far (graph replays count their nodes). */
lqft_status lqft_kernel_launches(lqft_handle* h, uint64_t* n);
/* How often this handle has taken the heights into its 16-bit storage or confirmed them there: a hot start, an upload
 * that fits the band, a conversion from the i32 arrays, a re-centring when the sweep budget of the last measurement
 * was used up.  0 = every sweep went through the i32 kernels (X not a power of two in [16, 1024], LQFT_FLAG_NO_NARROW,
 * or heights outside the guard band around the chain's normalisation offset).  lqft_plan_info says which copy is
 * live right now (storage=u16|i32). */
lqft_status lqft_narrow_segments(lqft_handle* h, uint64_t* n);
/* One line describing the launch plan the handle would use for its update kernels right now, e.g.
 * "path=narrow chunks=37 rows_per_chunk=7 plane_blocks=16 blocks_per_sm=4 ghost_depth=0 p2p=0" (for logs and for
 * tests that must know which kernel and which chunking a parity check exercised). */
lqft_status lqft_plan_info(lqft_handle* h, char* buf, size_t cap);
/* The handle's cudaStream_t (for callers that time with their own events). */
lqft_status lqft_stream(lqft_handle* h, void** stream);

/* ---- decision spec: pure host arithmetic, no device needed ------------------------- */
/* Philox4x32-10 (Salmon et al. 2011; constants of Random123).  key = seed lo/hi. */
void lqft_host_philox4x32(uint64_t seed, const uint32_t ctr[4], uint32_t out[4]);
/* The 32-bit word that drives site `idx` (global lexicographic) of colour stream
 * (x+y+t)&1 in sweep `sweep`: with grp = idx>>3 (64 bits), counter =
 * (grp & 0xffffffff, colour | (grp>>32)<<8, sweep_lo, sweep_hi), word (idx>>1)&3; the second
 * word is just `colour` below 2^35 sites.  Bit 31 = coin (1 -> +1, 0 -> -1; algorithm.rs:138-143), low 31 bits u:
 * draw = u * 2^-31 (algorithm.rs:151). */
uint32_t lqft_host_site_word(uint64_t seed, uint64_t idx, uint32_t colour, uint64_t sweep);
/* Threshold table for `beta`: entry k serves m = k - 3 where the proposal changes the local
 * action by  new - old = 6 + 2m  (m = coin_sign * (6 v - sum_nbrs + wilson_offset)):
 * thr[k] = min(floor(exp(-(6+2m) * beta) * 2^31), 2^31 - 1); accept iff u <= thr[k]
 * (<=> draw <= exp((old-new)*beta), algorithm.rs:152-153).  m beyond the table maps to the
 * last entry, which is 0.  Returns the number of entries written, 0 if beta is not
 * representable (beta <= 0 or more than LQFT_MAX_THRESHOLDS entries needed). */
uint32_t lqft_host_threshold_table(double beta, uint32_t* thr, uint32_t capacity);
/* Uniform site in [0, n_sites) from (seed, stream, counter, rep): used by normalize and by
 * the correlator reference sites. */
uint64_t lqft_host_pick_site(uint64_t seed, uint32_t stream, uint64_t counter, uint32_t rep,
                             uint64_t n_sites);
/* Hot-start height of site idx (int8 range). */
int32_t lqft_host_hot_value(uint64_t seed, uint64_t idx);
/* Lattice::calc_index_from_coords / calc_coords_from_index (lattice.rs:93-118) and the
 * neighbour row [+x,+y,+t,-x,-y,-t] of Lattice::values (lattice.rs:49-51,65-80). */
uint64_t lqft_host_index(uint32_t X, uint32_t Y, uint32_t T, uint32_t x, uint32_t y, uint32_t t);
void lqft_host_coords(uint32_t X, uint32_t Y, uint32_t T, uint64_t idx, uint32_t xyz[3]);
void lqft_host_neighbours(uint32_t X, uint32_t Y, uint32_t T, uint64_t idx, uint64_t nbr[6]);
/* Wilson offset of a site: +1 on the y==0 side of a shifted bond, -1 on the y==1 side, else
 * 0 (fields.rs:556-561, 821-837). */
int32_t lqft_host_wilson_offset(uint32_t Y, uint32_t x, uint32_t y, uint32_t t, uint32_t width,
                                uint32_t height);
/* Slab plan of rank r of ws along t: first owned plane and plane count (T % (2*ws) == 0). */
lqft_status lqft_host_slab(uint32_t T, uint32_t world_size, uint32_t rank, uint32_t* t0,
                           uint32_t* nt);

#ifdef __cplusplus
}
#endif
#endif /* LQFT_H */

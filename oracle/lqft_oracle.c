/*
 * lqft_oracle.c — CPU restatement of the reference's Metropolis + observables path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load
 * or call it, and only as the checker or the timed CPU baseline.  liblqft never links it.
 *
 * What it restates (paths relative to the reference checkout, TimeTeleporter/lattice_qft):
 *   geometry        src/lattice.rs:22-34 (move_one), :65-80 (neighbour table), :93-118 (index<->coords)
 *   local action    src/fields.rs:743-750 (Field::assumed_local_action)
 *   Wilson bonds    src/fields.rs:282-291 (BondsFieldNew::get_value), :550-567 (from_field),
 *                   :809-837 (assumed_local_action / assumed_bond_difference)
 *   accept/reject   src/algorithm.rs:137-157 (metropolis_single) = :185-205 (wilson_single)
 *   sweep order     src/algorithm.rs:117-121 (lexicographic); red-black order added for the GPU
 *   energy/action   src/fields.rs:728-737, 719-726, 660-671
 *   correlator      src/outputdata.rs:313-344
 *   normalize       src/fields.rs:621-632
 *   schedule        src/computation.rs:225-270, 304-340
 *   kahan/welford   src/kahan.rs:6-49, 78-138
 *
 * Pinning.  The reference cannot be built here (no Rust toolchain; the crate needs a 2023
 * nightly), and it is unseeded (ThreadRng, computation.rs:222), so no reference trajectory
 * exists to compare with: the sweep itself is "parity unpinned" at the RNG boundary.  What IS
 * pinned, in tests/test_oracle_golden.py: every known-answer fact of the reference's own unit
 * tests that touches this path (lattice.rs:191-239, fields.rs:851-896, kahan.rs:57-75,140-168),
 * the Philox4x32-10 known-answer vectors of Random123, and the published ensemble energies of
 * data/comp_energy_stats (statistical, tests/test_ensemble.py).
 *
 * Randomness.  The reference draws `coin: bool` and `draw: f64 in [0,1]` from ThreadRng per
 * site (algorithm.rs:138,151).  Here both come from one 32-bit word of a counter-based
 * Philox4x32-10 stream keyed by (seed; site, colour, sweep) — see orc_site_word — so that a
 * GPU and this file can be fed identical per-site uniforms.  The accept test is evaluated
 * exactly as the reference writes it, in f64 with libm exp.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_MAX_D 8

/* ---------------------------------------------------------------------------------------- */
/* geometry: src/lattice.rs                                                                    */
/* ---------------------------------------------------------------------------------------- */
typedef struct orc_lattice {
    uint32_t d;
    uint64_t size[ORC_MAX_D];
    uint64_t n_sites;
    uint64_t* values; /* n_sites * 2d neighbour indices: first d = +dir, last d = -dir (:49-51) */
} orc_lattice;

/* lattice.rs:93-104 */
uint64_t orc_index_from_coords(const orc_lattice* l, const uint64_t* coords) {
    uint64_t sum = 0;
    for (uint32_t i = 0; i < l->d; ++i) {
        uint64_t product = coords[i] % l->size[i]; /* cleanup, :14-19 */
        for (uint32_t j = 0; j < i; ++j) product *= l->size[j];
        sum += product;
    }
    return sum;
}

/* lattice.rs:106-118 */
void orc_coords_from_index(const orc_lattice* l, uint64_t index, uint64_t* coords) {
    for (uint32_t i = 0; i < l->d; ++i) {
        uint64_t product = 1;
        for (uint32_t j = 0; j < i; ++j) product *= l->size[j];
        coords[i] = (index / product) % l->size[i];
    }
}

/* lattice.rs:22-34 */
static void move_one(const orc_lattice* l, uint64_t* coords, uint32_t direction) {
    int is_forward = direction < l->d;
    uint32_t dir = direction % l->d;
    coords[dir] = is_forward ? coords[dir] + 1 : coords[dir] + (l->size[dir] - 1);
    for (uint32_t i = 0; i < l->d; ++i) coords[i] %= l->size[i];
}

/* lattice.rs:65-80 */
orc_lattice* orc_lattice_new(uint32_t d, const uint64_t* size) {
    if (d == 0 || d > ORC_MAX_D) return NULL;
    orc_lattice* l = (orc_lattice*)calloc(1, sizeof *l);
    l->d = d;
    l->n_sites = 1;
    for (uint32_t i = 0; i < d; ++i) {
        l->size[i] = size[i];
        l->n_sites *= size[i];
    }
    l->values = (uint64_t*)malloc(sizeof(uint64_t) * l->n_sites * 2 * d);
    for (uint64_t index = 0; index < l->n_sites; ++index) {
        for (uint32_t direction = 0; direction < 2 * d; ++direction) {
            uint64_t c[ORC_MAX_D];
            orc_coords_from_index(l, index, c);
            move_one(l, c, direction);
            l->values[index * 2 * d + direction] = orc_index_from_coords(l, c);
        }
    }
    return l;
}

void orc_lattice_free(orc_lattice* l) {
    if (!l) return;
    free(l->values);
    free(l);
}

uint64_t orc_n_sites(const orc_lattice* l) { return l->n_sites; }
const uint64_t* orc_neighbours(const orc_lattice* l, uint64_t index) {
    return l->values + index * 2 * l->d;
}

/* ---------------------------------------------------------------------------------------- */
/* randomness: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), written from the paper      */
/* ---------------------------------------------------------------------------------------- */
void orc_philox4x32(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The word that drives site idx: counter (idx>>3, colour, sweep lo, sweep hi), word (idx>>1)&3. */
uint32_t orc_site_word(uint64_t seed, uint64_t idx, uint32_t colour, uint64_t sweep) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t grp = idx >> 3;
    uint32_t ctr[4] = {(uint32_t)grp, colour | ((uint32_t)(grp >> 32) << 8), (uint32_t)sweep,
                       (uint32_t)(sweep >> 32)};
    uint32_t out[4];
    orc_philox4x32(key, ctr, out);
    return out[(idx >> 1) & 3u];
}

uint64_t orc_pick_site(uint64_t seed, uint32_t stream, uint64_t counter, uint32_t rep, uint64_t n) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {rep, stream, (uint32_t)counter, (uint32_t)(counter >> 32)};
    uint32_t out[4];
    orc_philox4x32(key, ctr, out);
    uint64_t v = (uint64_t)out[0] | ((uint64_t)out[1] << 32);
    return (uint64_t)(((unsigned __int128)v * (unsigned __int128)n) >> 64);
}

/* Hot start: Field::<i8>::random -> i32 (computation.rs:225-226): uniform in [-128,127]. */
int32_t orc_hot_value(uint64_t seed, uint64_t idx) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t g = idx >> 2;
    uint32_t ctr[4] = {(uint32_t)g, 3u, (uint32_t)(g >> 32), 0u};
    uint32_t out[4];
    orc_philox4x32(key, ctr, out);
    return (int32_t)(int8_t)(out[idx & 3u] >> 24);
}

void orc_init_hot(const orc_lattice* l, int32_t* values, uint64_t seed) {
    for (uint64_t i = 0; i < l->n_sites; ++i) values[i] = orc_hot_value(seed, i);
}

/* ---------------------------------------------------------------------------------------- */
/* fields: src/fields.rs                                                                       */
/* ---------------------------------------------------------------------------------------- */
/* fields.rs:743-750 — wrapping i32 arithmetic like a release build */
static int32_t assumed_local_action(const orc_lattice* l, const int32_t* values, uint64_t index,
                                    int32_t value) {
    uint32_t sum = 0;
    const uint64_t* nb = orc_neighbours(l, index);
    for (uint32_t i = 0; i < 2 * l->d; ++i) {
        uint32_t diff = (uint32_t)value - (uint32_t)values[nb[i]];
        sum += diff * diff;
    }
    return (int32_t)sum;
}

int32_t orc_local_action(const orc_lattice* l, const int32_t* values, uint64_t index, int32_t value) {
    return assumed_local_action(l, values, index, value);
}

/* LinksField over BondsFieldNew<bool> (fields.rs:150-291, 350-398): links[index*3 + dir], dir < 3 */
typedef struct orc_wilson {
    uint8_t* links; /* n_sites * 3 */
    int spin;       /* always true, fields.rs:565 */
} orc_wilson;

/* BondsFieldNew::get_value (fields.rs:282-291) */
static int links_is_active(const orc_lattice* l, const orc_wilson* w, uint64_t index, uint32_t direction) {
    uint32_t d = l->d;
    if (direction % (2 * d) < d) return w->links[index * d + direction];
    uint64_t neighbour = orc_neighbours(l, index)[direction];
    return w->links[neighbour * d + direction % d];
}

/* WilsonField::from_field (fields.rs:550-567); 3-D only */
orc_wilson* orc_wilson_new(const orc_lattice* l, uint64_t width, uint64_t height) {
    if (l->d != 3) return NULL;
    orc_wilson* w = (orc_wilson*)calloc(1, sizeof *w);
    w->links = (uint8_t*)calloc(l->n_sites * 3, 1);
    w->spin = 1;
    for (uint64_t index = 0; index < l->n_sites; ++index) {
        uint64_t c[ORC_MAX_D];
        orc_coords_from_index(l, index, c);
        if (c[1] == 0 && c[0] < width && c[2] < height) w->links[index * 3 + 1] = 1; /* activate(index, 1) */
    }
    return w;
}
void orc_wilson_free(orc_wilson* w) {
    if (!w) return;
    free(w->links);
    free(w);
}
int orc_wilson_is_active(const orc_lattice* l, const orc_wilson* w, uint64_t index, uint32_t direction) {
    return links_is_active(l, w, index, direction);
}

/* fields.rs:821-837 */
static int32_t wilson_assumed_bond_difference(const orc_lattice* l, const int32_t* values,
                                              const orc_wilson* w, uint64_t index, uint32_t direction,
                                              int32_t value) {
    uint64_t neighbour = orc_neighbours(l, index)[direction];
    if (links_is_active(l, w, index, direction)) {
        int32_t direction_modifier = direction == 1 ? 1 : (direction == 4 ? -1 : 0);
        int32_t spin_modifier = w->spin ? 1 : -1;
        return (int32_t)((uint32_t)value - (uint32_t)values[neighbour] +
                         (uint32_t)(direction_modifier * spin_modifier));
    }
    return (int32_t)((uint32_t)value - (uint32_t)values[neighbour]); /* fields.rs:756-759 */
}

/* fields.rs:809-815 */
static int32_t wilson_assumed_local_action(const orc_lattice* l, const int32_t* values,
                                           const orc_wilson* w, uint64_t index, int32_t value) {
    uint32_t sum = 0;
    for (uint32_t direction = 0; direction < 6; ++direction) {
        uint32_t diff = (uint32_t)wilson_assumed_bond_difference(l, values, w, index, direction, value);
        sum += diff * diff;
    }
    return (int32_t)sum;
}

int32_t orc_wilson_bond_action(const orc_lattice* l, const int32_t* values, const orc_wilson* w,
                               uint64_t index, uint32_t direction) {
    int32_t d = wilson_assumed_bond_difference(l, values, w, index, direction, values[index]);
    return (int32_t)((uint32_t)d * (uint32_t)d);
}

/* ---------------------------------------------------------------------------------------- */
/* algorithm: src/algorithm.rs                                                                 */
/* ---------------------------------------------------------------------------------------- */
/* metropolis_single (algorithm.rs:127-158) / wilson_single (:178-206) with the two random
 * values passed in.  Returns 1 if accepted. */
int orc_metropolis_single(const orc_lattice* l, int32_t* values, const orc_wilson* w, uint64_t index,
                          double temp, int coin, double draw) {
    int32_t old_value = values[index];
    int32_t new_value = coin ? old_value + 1 : old_value - 1;
    int32_t old_action, new_action;
    if (w) {
        old_action = wilson_assumed_local_action(l, values, w, index, old_value);
        new_action = wilson_assumed_local_action(l, values, w, index, new_value);
    } else {
        old_action = assumed_local_action(l, values, index, old_value);
        new_action = assumed_local_action(l, values, index, new_value);
    }
    int32_t delta = (int32_t)((uint32_t)old_action - (uint32_t)new_action);
    double prob = exp((double)delta * temp);
    if (draw <= prob) {
        values[index] = new_value;
        return 1;
    }
    return 0;
}

static int site_colour(const orc_lattice* l, uint64_t index) {
    uint64_t c[ORC_MAX_D], s = 0;
    orc_coords_from_index(l, index, c);
    for (uint32_t i = 0; i < l->d; ++i) s += c[i];
    return (int)(s & 1u);
}

static int update_site(const orc_lattice* l, int32_t* values, const orc_wilson* w, uint64_t index,
                       int colour, double temp, uint64_t seed, uint64_t sweep) {
    uint32_t word = orc_site_word(seed, index, (uint32_t)colour, sweep);
    int coin = (int)(word >> 31);                                    /* 1 -> +1 (algorithm.rs:140-143) */
    double draw = (double)(word & 0x7fffffffu) * (1.0 / 2147483648.0); /* u * 2^-31 */
    return orc_metropolis_single(l, values, w, index, temp, coin, draw);
}

/* One sweep.  order 0: `for index in 0..SIZE` (algorithm.rs:117-121);  order 1: red-black — all
 * sites with even coordinate sum in index order, then all odd ones.  Returns #accepted. */
uint64_t orc_sweep(const orc_lattice* l, int32_t* values, const orc_wilson* w, double temp,
                   uint64_t seed, uint64_t sweep, int order) {
    uint64_t acceptance = 0;
    if (order == 0) {
        for (uint64_t index = 0; index < l->n_sites; ++index)
            acceptance += (uint64_t)update_site(l, values, w, index, site_colour(l, index), temp, seed, sweep);
    } else {
        for (int colour = 0; colour < 2; ++colour)
            for (uint64_t index = 0; index < l->n_sites; ++index)
                if (site_colour(l, index) == colour)
                    acceptance += (uint64_t)update_site(l, values, w, index, colour, temp, seed, sweep);
    }
    return acceptance;
}

/* ---------------------------------------------------------------------------------------- */
/* observables                                                                                 */
/* ---------------------------------------------------------------------------------------- */
/* Field::bond_action (fields.rs:683-686, 752-759) */
static int32_t bond_action(const orc_lattice* l, const int32_t* values, uint64_t index, uint32_t direction) {
    uint32_t diff = (uint32_t)values[index] - (uint32_t)values[orc_neighbours(l, index)[direction]];
    return (int32_t)(diff * diff);
}

/* integer_energy (fields.rs:728-737): accumulated in T = i32 (wrapping, quirk Q5).  Returns the
 * exact int64 sum; *wrapped receives what the reference's release build would hold. */
int64_t orc_integer_energy(const orc_lattice* l, const int32_t* values, int32_t* wrapped) {
    uint32_t acc32 = 0;
    int64_t acc64 = 0;
    uint32_t d = l->d;
    for (uint64_t index = 0; index < l->n_sites; ++index) {
        uint32_t site = 0;
        int64_t site64 = 0;
        for (uint32_t direction = 0; direction + 1 < d; ++direction) {
            int32_t b = bond_action(l, values, index, direction);
            site += (uint32_t)b;
            site64 += b;
        }
        int32_t bt = bond_action(l, values, index, d - 1);
        site -= (uint32_t)bt;
        site64 -= bt;
        acc32 += site;
        acc64 += site64;
    }
    if (wrapped) *wrapped = (int32_t)acc32;
    return acc64;
}

/* energy_observable (fields.rs:669-671) on the exact sum */
double orc_energy_observable(const orc_lattice* l, const int32_t* values, double temp) {
    return temp * (double)orc_integer_energy(l, values, NULL);
}

/* float_action (fields.rs:719-726): sequential f64 sum of the integer bond actions */
double orc_float_action(const orc_lattice* l, const int32_t* values) {
    double sum = 0.0;
    for (uint64_t index = 0; index < l->n_sites; ++index)
        for (uint32_t direction = 0; direction < l->d; ++direction)
            sum += (double)bond_action(l, values, index, direction);
    return sum;
}
double orc_action_observable(const orc_lattice* l, const int32_t* values, double temp) {
    return temp * orc_float_action(l, values);
}

/* CorrelationPlotOutputData::update (outputdata.rs:313-344) with the reference sites passed in */
void orc_correlator(const orc_lattice* l, const int32_t* values, const uint64_t* x_indices,
                    uint64_t repetitions, double* correlation_function) {
    const uint32_t T_DIM = 2;
    uint64_t max_t = l->size[T_DIM];
    for (uint64_t t = 0; t < max_t; ++t) correlation_function[t] = 0.0;
    for (uint64_t rep = 0; rep < repetitions; ++rep) {
        uint64_t x_index = x_indices[rep];
        uint64_t c[ORC_MAX_D];
        orc_coords_from_index(l, x_index, c);
        uint64_t x_t_coord = c[T_DIM];
        int32_t x_value = values[x_index];
        for (uint64_t index = 0; index < l->n_sites; ++index) {
            orc_coords_from_index(l, index, c);
            uint64_t t = (c[T_DIM] + max_t - x_t_coord) % max_t;
            uint32_t diff = (uint32_t)x_value - (uint32_t)values[index];
            int32_t squared = (int32_t)(diff * diff);
            correlation_function[t] = correlation_function[t] + (double)squared;
        }
    }
}

/* shift_values (fields.rs:628-632) */
void orc_shift_values(int32_t* values, uint64_t n, int32_t shift) {
    for (uint64_t i = 0; i < n; ++i) values[i] = (int32_t)((uint32_t)values[i] - (uint32_t)shift);
}

/* ---------------------------------------------------------------------------------------- */
/* kahan.rs                                                                                    */
/* ---------------------------------------------------------------------------------------- */
typedef struct orc_kahan {
    double value, kahan;
    uint64_t entries;
} orc_kahan;

void orc_kahan_add(orc_kahan* k, double value) { /* kahan.rs:21-31 */
    double y = value - k->kahan;
    double t = k->value + y;
    k->kahan = (t - k->value) - y;
    k->value = t;
    k->entries += 1;
}
void orc_kahan_add_f32(float* value, float* kahan, float x) { /* KahanSummation<f32>, kahan.rs:69-75 */
    volatile float y = x - *kahan;
    volatile float t = *value + y;
    *kahan = (t - *value) - y;
    *value = t;
}
double orc_kahan_mean(const orc_kahan* k) { return k->entries ? k->value / (double)k->entries : 0.0; }

typedef struct orc_welford {
    orc_kahan sum, diff_square_sum;
    uint64_t entries;
} orc_welford;

double orc_welford_mean(const orc_welford* w) { return w->entries ? w->sum.value / (double)w->entries : 0.0; }
void orc_welford_update(orc_welford* w, double value) { /* kahan.rs:100-107 */
    double old_mean = orc_welford_mean(w);
    orc_kahan_add(&w->sum, value);
    w->entries += 1;
    double new_mean = orc_welford_mean(w);
    orc_kahan_add(&w->diff_square_sum, (value - old_mean) * (value - new_mean));
}
double orc_welford_var(const orc_welford* w) { return w->entries ? w->diff_square_sum.value / (double)w->entries : 0.0; }
double orc_welford_var_unbiased(const orc_welford* w) {
    return w->entries > 1 ? w->diff_square_sum.value / (double)(w->entries - 1) : 0.0;
}
/* kahan.rs:135-137 (sic: var / sqrt(n)) */
double orc_welford_standard_error(const orc_welford* w) {
    return orc_welford_var_unbiased(w) / sqrt((double)w->entries);
}

/* DifferencePlotOutputData::update (outputdata.rs:241-257): wel is [3][n_sites] Welford cells */
void orc_difference_update(const orc_lattice* l, const int32_t* values, orc_welford* wel) {
    for (uint32_t direction = 0; direction < l->d; ++direction)
        for (uint64_t index = 0; index < l->n_sites; ++index) {
            int32_t diff = (int32_t)((uint32_t)values[index] -
                                     (uint32_t)values[orc_neighbours(l, index)[direction]]);
            orc_welford_update(&wel[(uint64_t)direction * l->n_sites + index], (double)diff);
        }
}
void orc_difference_means(const orc_lattice* l, const orc_welford* wel, double* mean) {
    for (uint64_t i = 0; i < (uint64_t)l->d * l->n_sites; ++i) mean[i] = orc_welford_mean(&wel[i]);
}
uint64_t orc_sizeof_welford(void) { return sizeof(orc_welford); }

/* ---------------------------------------------------------------------------------------- */
/* schedule: Simulation::run / WilsonSim::run loop bodies (computation.rs:229-270, 313-340)     */
/* ---------------------------------------------------------------------------------------- */
/* steps [first_step, first_step+n_steps): sweep (RNG counter sweep_base+step); energy if
 * step % energy_every == 0 (plain field, also for Wilson chains — computation.rs:335);
 * normalize_random if step % normalize_every == 0 with the site picked from stream 2.
 * Returns the number of energies written to e_obs. */
uint64_t orc_run(const orc_lattice* l, int32_t* values, const orc_wilson* w, double temp, uint64_t seed,
                 uint64_t sweep_base, uint64_t first_step, uint64_t n_steps, uint32_t energy_every,
                 uint32_t normalize_every, int order, double* e_obs, uint64_t* accepted) {
    uint64_t n = 0, acc = 0;
    for (uint64_t step = first_step; step < first_step + n_steps; ++step) {
        acc += orc_sweep(l, values, w, temp, seed, sweep_base + step, order);
        if (energy_every && step % energy_every == 0) {
            double e = orc_energy_observable(l, values, temp);
            if (e_obs) e_obs[n] = e;
            ++n;
        }
        if (normalize_every && step % normalize_every == 0) {
            uint64_t site = orc_pick_site(seed, 2u, sweep_base + step, 0u, l->n_sites);
            orc_shift_values(values, l->n_sites, values[site]);
        }
    }
    if (accepted) *accepted = acc;
    return n;
}

/* ---------------------------------------------------------------------------------------- */
/* timed CPU baseline: the reference's data path — lexicographic in-place sweep over a           */
/* precomputed neighbour table, i32 heights, two local-action evaluations + exp per site —      */
/* run as independent chains, one per thread (rayon usage of sim.rs:42-45, 90-93).              */
/* ---------------------------------------------------------------------------------------- */
typedef struct bench_arg {
    const orc_lattice* l;
    double temp;
    uint64_t seed;
    uint64_t sweeps;
    uint32_t energy_every;
    double e_last;
    uint64_t accepted;
    pthread_barrier_t* start;   /* every chain has built its field and colour table */
    struct timespec t0, t1;     /* this chain's sweep loop */
} bench_arg;

static void* bench_thread(void* p) {
    bench_arg* a = (bench_arg*)p;
    int32_t* values = (int32_t*)malloc(sizeof(int32_t) * a->l->n_sites);
    orc_init_hot(a->l, values, a->seed);
    uint8_t* colour = (uint8_t*)malloc(a->l->n_sites);
    for (uint64_t i = 0; i < a->l->n_sites; ++i) colour[i] = (uint8_t)site_colour(a->l, i);
    uint64_t acc = 0;
    const uint32_t key[2] = {(uint32_t)a->seed, (uint32_t)(a->seed >> 32)};
    pthread_barrier_wait(a->start);
    clock_gettime(CLOCK_MONOTONIC, &a->t0);
    for (uint64_t s = 0; s < a->sweeps; ++s) {
        /* same words as orc_site_word, but each Philox block (8 consecutive indices = 4 sites per
         * colour) is generated once instead of once per site */
        uint32_t blk[2][4];
        for (uint64_t index = 0; index < a->l->n_sites; ++index) {
            if ((index & 7u) == 0) {
                uint64_t grp = index >> 3;
                for (uint32_t c = 0; c < 2; ++c) {
                    uint32_t ctr[4] = {(uint32_t)grp, c | ((uint32_t)(grp >> 32) << 8), (uint32_t)s,
                                       (uint32_t)(s >> 32)};
                    orc_philox4x32(key, ctr, blk[c]);
                }
            }
            uint32_t word = blk[colour[index]][(index >> 1) & 3u];
            int coin = (int)(word >> 31);
            double draw = (double)(word & 0x7fffffffu) * (1.0 / 2147483648.0);
            acc += (uint64_t)orc_metropolis_single(a->l, values, NULL, index, a->temp, coin, draw);
        }
        if (a->energy_every && s % a->energy_every == 0) a->e_last = orc_energy_observable(a->l, values, a->temp);
        if (s % 10 == 0) {
            uint64_t site = orc_pick_site(a->seed, 2u, s, 0u, a->l->n_sites);
            orc_shift_values(values, a->l->n_sites, values[site]);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &a->t1);
    a->accepted = acc;
    free(colour);
    free(values);
    return NULL;
}

/* Runs `threads` independent chains of `sweeps` lexicographic sweeps each.  Returns the wall seconds of the sweep
 * loops alone: thread start, hot start and colour table are built before a barrier, the clock runs from the first
 * chain entering its loop to the last one leaving it. */
double orc_bench_chains(const orc_lattice* l, double temp, uint64_t seed, uint64_t sweeps,
                        uint32_t energy_every, uint32_t threads, double* e_last) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    bench_arg* args = (bench_arg*)calloc(threads, sizeof(bench_arg));
    pthread_barrier_t start;
    pthread_barrier_init(&start, NULL, threads);
    for (uint32_t i = 0; i < threads; ++i) {
        args[i].l = l;
        args[i].temp = temp;
        args[i].seed = seed + i;
        args[i].sweeps = sweeps;
        args[i].energy_every = energy_every;
        args[i].start = &start;
        pthread_create(&th[i], NULL, bench_thread, &args[i]);
    }
    for (uint32_t i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    pthread_barrier_destroy(&start);
    double first = 0.0, last = 0.0;
    for (uint32_t i = 0; i < threads; ++i) {
        const double a = (double)args[i].t0.tv_sec + 1e-9 * (double)args[i].t0.tv_nsec;
        const double b = (double)args[i].t1.tv_sec + 1e-9 * (double)args[i].t1.tv_nsec;
        if (i == 0 || a < first) first = a;
        if (i == 0 || b > last) last = b;
    }
    if (e_last) *e_last = args[0].e_last;
    free(args);
    free(th);
    return last - first;
}

/* ---------------------------------------------------------------------------------------- */
/* slab form (multi-GPU decomposition check): the same per-site rule on one rank's t-slab      */
/* ---------------------------------------------------------------------------------------- */
/* `values` holds (tl + 2) planes of X*Y sites in lexicographic order: plane 0 = ghost (global
 * plane t0-1), planes 1..tl = owned (global t0 .. t0+tl-1), plane tl+1 = ghost.  One half-sweep of
 * `colour` over the owned planes, RNG keyed by the GLOBAL site index, Wilson sheet by global
 * coordinates (fields.rs:556-561).  Returns #accepted. */
uint64_t orc_half_sweep_slab(uint32_t X, uint32_t Y, uint32_t T, uint32_t t0, uint32_t tl, int32_t* values,
                             int colour, double temp, uint64_t seed, uint64_t sweep, uint32_t wil_w,
                             uint32_t wil_h) {
    (void)T;
    uint64_t acc = 0;
    const uint64_t plane = (uint64_t)X * Y;
    for (uint32_t p = 1; p <= tl; ++p) {
        const uint32_t t = t0 + p - 1;
        for (uint32_t y = 0; y < Y; ++y)
            for (uint32_t x = 0; x < X; ++x) {
                if (((x + y + t) & 1u) != (uint32_t)colour) continue;
                const uint64_t here = p * plane + (uint64_t)y * X + x;
                const uint64_t nb[6] = {
                    p * plane + (uint64_t)y * X + (x + 1) % X,       p * plane + (uint64_t)((y + 1) % Y) * X + x,
                    (p + 1) * plane + (uint64_t)y * X + x,           p * plane + (uint64_t)y * X + (x + X - 1) % X,
                    p * plane + (uint64_t)((y + Y - 1) % Y) * X + x, (p - 1) * plane + (uint64_t)y * X + x};
                /* Wilson: +1 seen from (x,0,t) along +y, -1 seen from (x,1,t) along -y */
                int32_t mod[6] = {0, 0, 0, 0, 0, 0};
                if (wil_w && wil_h && x < wil_w && t < wil_h) {
                    if (y == 0) mod[1] = 1;
                    if (y == 1 % Y) mod[4] = -1;
                }
                const uint64_t gidx = (uint64_t)t * plane + (uint64_t)y * X + x;
                const uint32_t word = orc_site_word(seed, gidx, (uint32_t)colour, sweep);
                const int coin = (int)(word >> 31);
                const double draw = (double)(word & 0x7fffffffu) * (1.0 / 2147483648.0);
                const int32_t old_value = values[here];
                const int32_t new_value = coin ? old_value + 1 : old_value - 1;
                uint32_t old_action = 0, new_action = 0;
                for (int d = 0; d < 6; ++d) {
                    const uint32_t d0 = (uint32_t)old_value - (uint32_t)values[nb[d]] + (uint32_t)mod[d];
                    const uint32_t d1 = (uint32_t)new_value - (uint32_t)values[nb[d]] + (uint32_t)mod[d];
                    old_action += d0 * d0;
                    new_action += d1 * d1;
                }
                const double prob = exp((double)(int32_t)(old_action - new_action) * temp);
                if (draw <= prob) {
                    values[here] = new_value;
                    ++acc;
                }
            }
    }
    return acc;
}

/* ---------------------------------------------------------------------------------------- */
/* Field<f64,...>: the same generic code paths instantiated with T = f64 (fields.rs:573-599)   */
/* ---------------------------------------------------------------------------------------- */
static double assumed_local_action_f64(const orc_lattice* l, const double* values, const orc_wilson* w,
                                       uint64_t index, double value) {
    double sum = 0.0;
    const uint64_t* nb = orc_neighbours(l, index);
    for (uint32_t direction = 0; direction < 2 * l->d; ++direction) {
        double diff = value - values[nb[direction]];
        if (w && links_is_active(l, w, index, direction)) { /* fields.rs:821-837 */
            double direction_modifier = direction == 1 ? 1.0 : (direction == 4 ? -1.0 : 0.0);
            diff = diff + direction_modifier * (w->spin ? 1.0 : -1.0);
        }
        sum = sum + diff * diff;
    }
    return sum;
}

uint64_t orc_sweep_f64(const orc_lattice* l, double* values, const orc_wilson* w, double temp, uint64_t seed,
                       uint64_t sweep, int order) {
    uint64_t acceptance = 0;
    for (int pass = 0; pass < (order ? 2 : 1); ++pass)
        for (uint64_t index = 0; index < l->n_sites; ++index) {
            const int colour = site_colour(l, index);
            if (order && colour != pass) continue;
            const uint32_t word = orc_site_word(seed, index, (uint32_t)colour, sweep);
            const int coin = (int)(word >> 31);
            const double draw = (double)(word & 0x7fffffffu) * (1.0 / 2147483648.0);
            const double old_value = values[index];
            const double new_value = coin ? old_value + 1.0 : old_value - 1.0; /* T::from(1_i8) */
            const double old_action = assumed_local_action_f64(l, values, w, index, old_value);
            const double new_action = assumed_local_action_f64(l, values, w, index, new_value);
            const double prob = exp((old_action - new_action) * temp);
            if (draw <= prob) {
                values[index] = new_value;
                ++acceptance;
            }
        }
    return acceptance;
}

/* integer_energy with T = f64 (fields.rs:728-737): the reference's left-to-right folds */
double orc_energy_f64(const orc_lattice* l, const double* values) {
    double acc = 0.0;
    for (uint64_t index = 0; index < l->n_sites; ++index) {
        const uint64_t* nb = orc_neighbours(l, index);
        double site = 0.0;
        for (uint32_t direction = 0; direction + 1 < l->d; ++direction) {
            double d = values[index] - values[nb[direction]];
            site = site + d * d;
        }
        double dt = values[index] - values[nb[l->d - 1]];
        site = site - dt * dt;
        acc = acc + site;
    }
    return acc;
}

double orc_action_f64(const orc_lattice* l, const double* values) { /* fields.rs:719-726 */
    double sum = 0.0;
    for (uint64_t index = 0; index < l->n_sites; ++index)
        for (uint32_t direction = 0; direction < l->d; ++direction) {
            double d = values[index] - values[orc_neighbours(l, index)[direction]];
            sum += d * d;
        }
    return sum;
}

void orc_correlator_f64(const orc_lattice* l, const double* values, const uint64_t* x_indices, uint64_t repetitions,
                        double* correlation_function) { /* outputdata.rs:313-344 */
    const uint32_t T_DIM = 2;
    uint64_t max_t = l->size[T_DIM];
    for (uint64_t t = 0; t < max_t; ++t) correlation_function[t] = 0.0;
    for (uint64_t rep = 0; rep < repetitions; ++rep) {
        uint64_t c[ORC_MAX_D];
        orc_coords_from_index(l, x_indices[rep], c);
        const uint64_t x_t_coord = c[T_DIM];
        const double x_value = values[x_indices[rep]];
        for (uint64_t index = 0; index < l->n_sites; ++index) {
            orc_coords_from_index(l, index, c);
            const uint64_t t = (c[T_DIM] + max_t - x_t_coord) % max_t;
            const double diff = x_value - values[index];
            correlation_function[t] = correlation_function[t] + diff * diff;
        }
    }
}

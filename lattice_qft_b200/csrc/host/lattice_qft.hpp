// lattice_qft.hpp — C++17 host mirror of the reference crate's modules for the accelerated path, over
// the C ABI of include/lqft.h.  Same names and argument meaning as the Rust items they stand for
// (file:line citations are relative to the reference checkout); extents are run-time values where the
// crate uses const generics.  Header only; link with -llqft.
#pragma once
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../../include/lqft.h"

namespace lattice_qft {

// Box<dyn Error> of the reference: every failed ABI call throws with lqft_last_error().
struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(int status) {
    if (status != LQFT_OK) throw Error(status, lqft_last_error());
}

// ---- src/lattice.rs ---------------------------------------------------------------------------
namespace lattice {
struct Lattice3d {  // lattice.rs:127-161
    std::array<uint32_t, 3> size;
    Lattice3d(uint32_t max_x, uint32_t max_y, uint32_t max_t) : size{max_x, max_y, max_t} {}
    uint64_t n_sites() const { return (uint64_t)size[0] * size[1] * size[2]; }
    uint64_t calc_index_from_coords(uint32_t x, uint32_t y, uint32_t t) const {  // lattice.rs:93-104
        return lqft_host_index(size[0], size[1], size[2], x, y, t);
    }
    std::array<uint32_t, 3> calc_coords_from_index(uint64_t index) const {  // lattice.rs:106-118
        std::array<uint32_t, 3> c{};
        lqft_host_coords(size[0], size[1], size[2], index, c.data());
        return c;
    }
    std::array<uint64_t, 6> get_neighbours_array(uint64_t index) const {  // lattice.rs:83-85, order :49-51
        std::array<uint64_t, 6> n{};
        lqft_host_neighbours(size[0], size[1], size[2], index, n.data());
        return n;
    }
};
}  // namespace lattice

// ---- the device handle (what Simulation::run owns for a whole run) ------------------------------
class Gpu {
    lqft_handle* h_ = nullptr;
    uint32_t n_chains_ = 0;

  public:
    Gpu(const lattice::Lattice3d& l, uint32_t n_chains = 1, int device = 0, uint32_t flags = 0) : n_chains_(n_chains) {
        lqft_config cfg{};
        cfg.x = l.size[0]; cfg.y = l.size[1]; cfg.t = l.size[2];
        cfg.dtype = LQFT_I32; cfg.n_chains = n_chains; cfg.device = device;
        cfg.world_size = 1; cfg.rank = 0; cfg.flags = flags;
        check(lqft_create(&cfg, &h_));
    }
    Gpu(const Gpu&) = delete;
    Gpu& operator=(const Gpu&) = delete;
    ~Gpu() { lqft_destroy(h_); }
    lqft_handle* raw() const { return h_; }
    uint32_t n_chains() const { return n_chains_; }
};

// ---- src/fields.rs ------------------------------------------------------------------------------
namespace fields {
struct Field {  // fields.rs:17-68 with T = i32 (computation.rs:226)
    const lattice::Lattice3d* lattice;
    std::shared_ptr<Gpu> gpu;
    uint64_t seed = 0;
    uint64_t sweep = 0;  // running sweep counter: the state a ThreadRng would carry
    uint32_t wilson_width = 0, wilson_height = 0;

    static Field new_(const lattice::Lattice3d& l, int device = 0) {  // Field::new
        Field f;
        f.lattice = &l;
        f.gpu = std::make_shared<Gpu>(l, 1, device);
        return f;
    }
    static Field random(const lattice::Lattice3d& l, uint64_t seed, int device = 0) {  // Field::random + from_field
        Field f = new_(l, device);
        f.seed = seed;
        check(lqft_set_chain(f.gpu->raw(), 0, 1.0, seed, 0, 0));
        check(lqft_init_hot(f.gpu->raw()));
        return f;
    }
    static Field from_values(const lattice::Lattice3d& l, const std::vector<int32_t>& values, int device = 0) {
        if (values.size() != l.n_sites()) throw Error(LQFT_ERR_INVALID, "values.len() != SIZE");  // fields.rs:64
        Field f = new_(l, device);
        check(lqft_set_chain(f.gpu->raw(), 0, 1.0, 0, 0, 0));
        check(lqft_upload(f.gpu->raw(), 0, values.data()));
        return f;
    }
    std::vector<int32_t> values() const {  // Field::values
        std::vector<int32_t> v(lattice->n_sites());
        check(lqft_download(gpu->raw(), 0, v.data()));
        return v;
    }
    // Action trait (fields.rs:651-768)
    int64_t integer_energy() const { int64_t e; check(lqft_energy(gpu->raw(), &e, nullptr)); return e; }
    int64_t integer_action() const { int64_t s; check(lqft_action(gpu->raw(), &s, nullptr)); return s; }
    double float_energy() const { return (double)integer_energy(); }
    double float_action() const { return (double)integer_action(); }
    double energy_observable(double temp) const { return temp * float_energy(); }
    double action_observable(double temp) const { return temp * float_action(); }
    // HeightField trait (fields.rs:604-633)
    void shift_values(int32_t shift) { check(lqft_shift_values(gpu->raw(), 0, (double)shift)); }
    void normalize_random() { check(lqft_normalize(gpu->raw(), sweep)); }
};

struct WilsonField {  // fields.rs:528-568
    Field field;
    static WilsonField from_field(Field f, uint32_t width, uint32_t height) {
        f.wilson_width = width;
        f.wilson_height = height;
        return WilsonField{std::move(f)};
    }
    void normalize_random() { field.normalize_random(); }
    void shift_values(int32_t s) { field.shift_values(s); }
};
}  // namespace fields

// ---- src/algorithm.rs ---------------------------------------------------------------------------
namespace algorithm {
struct Metropolis {  // algorithm.rs:105-207
    // Algorithm::field_sweep: returns the number of accepted proposals
    uint64_t field_sweep(fields::Field& field, double temp) const {
        check(lqft_set_chain(field.gpu->raw(), 0, temp, field.seed, 0, 0));
        uint64_t accepted = 0;
        check(lqft_sweep(field.gpu->raw(), field.sweep++, 1, &accepted));
        return accepted;
    }
    // WilsonAlgorithm::wilson_sweep
    uint64_t wilson_sweep(fields::WilsonField& w, double temp) const {
        fields::Field& f = w.field;
        check(lqft_set_chain(f.gpu->raw(), 0, temp, f.seed, f.wilson_width, f.wilson_height));
        uint64_t accepted = 0;
        check(lqft_sweep(f.gpu->raw(), f.sweep++, 1, &accepted));
        return accepted;
    }
};
struct AlgorithmType {  // algorithm.rs:11-47; the cluster variants are out of scope
    Metropolis algo;
    static AlgorithmType new_metropolis() { return AlgorithmType{}; }
    std::string to_string() const { return "Metropolis"; }
};
}  // namespace algorithm

// ---- src/outputdata.rs --------------------------------------------------------------------------
namespace outputdata {
enum class Kind { ActionObservable, DifferencePlot, CorrelationData, EnergyObservable };
struct OutputData {  // outputdata.rs:42-119
    Kind kind;
    double temp = 0.0;
    uint64_t repetitions = 0;
    uint64_t frequency = 1;
    std::vector<double> values;                // Observe::results
    std::vector<std::vector<double>> corr_fn;  // Plotting::plot of the correlator
    std::vector<double> difference_mean;       // [3][SIZE]

    static OutputData new_action_observable(double temp) { return {Kind::ActionObservable, temp}; }
    static OutputData new_energy_observable(double temp) { return {Kind::EnergyObservable, temp}; }
    static OutputData new_difference_plot(const lattice::Lattice3d&) { return {Kind::DifferencePlot}; }
    static OutputData new_correlation_plot(const lattice::Lattice3d&, uint64_t repetitions) {
        OutputData o{Kind::CorrelationData};
        o.repetitions = repetitions;
        return o;
    }
    OutputData set_frequency(uint64_t f) && { frequency = f; return std::move(*this); }
    bool is_step_for_next_update(uint64_t step) const { return step % frequency == 0; }

    // UpdateOutputData::update (outputdata.rs:114-149)
    void update(const fields::Field& field, uint64_t sweep_index) {
        lqft_handle* h = field.gpu->raw();
        switch (kind) {
            case Kind::EnergyObservable: { double e; check(lqft_energy(h, nullptr, &e)); values.push_back(e); break; }
            case Kind::ActionObservable: { double s; check(lqft_action(h, nullptr, &s)); values.push_back(s); break; }
            case Kind::CorrelationData: {
                std::vector<double> c(field.lattice->size[2]);
                check(lqft_correlator(h, 0, nullptr, (uint32_t)repetitions, sweep_index, c.data()));
                corr_fn.push_back(std::move(c));
                break;
            }
            case Kind::DifferencePlot: check(lqft_difference_update(h, 0)); break;
        }
    }
    void finish(const fields::Field& field) {
        if (kind != Kind::DifferencePlot) return;
        difference_mean.assign(3 * field.lattice->n_sites(), 0.0);
        uint64_t count = 0;
        check(lqft_difference_read(field.gpu->raw(), 0, difference_mean.data(), &count));
    }
};
}  // namespace outputdata

// ---- src/export.rs: the csv crate's number formatting ---------------------------------------------
namespace export_ {
inline std::string fmt_f64(double v) {  // ryu: shortest round trip, ".0" on integers
    char buf[64];
    for (int prec = 1; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    std::string s(buf);
    if (s.find('e') != std::string::npos) {  // fall back to positional inside ryu's positional range
        double a = std::fabs(v);
        if (a >= 1e-5 && a < 1e16) {
            snprintf(buf, sizeof buf, "%.17f", v);
            s = buf;
            while (s.back() == '0') s.pop_back();
            if (s.back() == '.') s += "0";
            // shortest digits that still round trip
            for (size_t n = s.find('.') + 2; n < s.size(); ++n)
                if (std::strtod(s.substr(0, n).c_str(), nullptr) == v) { s = s.substr(0, n); break; }
            return s;
        }
        size_t e = s.find('e');
        std::string mant = s.substr(0, e);
        int ex = std::atoi(s.c_str() + e + 1);
        return mant + "e" + std::to_string(ex);
    }
    if (s.find('.') == std::string::npos && s.find("inf") == std::string::npos && s.find("nan") == std::string::npos) s += ".0";
    return s;
}
}  // namespace export_

// ---- src/computation.rs -------------------------------------------------------------------------
namespace computation {
using outputdata::Kind;
using outputdata::OutputData;

struct Computation {  // enum Computation::{Simulation, WilsonSim} (computation.rs:18-103)
    const lattice::Lattice3d* lattice;
    double temp;
    algorithm::AlgorithmType algorithm;
    uint64_t burnin, iterations;
    std::vector<OutputData> output;
    bool wilson = false;
    uint32_t width = 0, height = 0;
    uint64_t seed = 0;
    int device = 0;
    int64_t duration = -1;  // Option<u64>

    static Computation new_simulation(const lattice::Lattice3d& l, double temp, algorithm::AlgorithmType a,
                                      uint64_t burnin, uint64_t iterations, std::vector<OutputData> output) {
        return Computation{&l, temp, a, burnin, iterations, std::move(output)};
    }
    static Computation new_wilson_sim(const lattice::Lattice3d& l, double temp, algorithm::AlgorithmType a,
                                      uint64_t burnin, uint64_t iterations, uint32_t width, uint32_t height,
                                      std::vector<OutputData> output) {
        Computation c{&l, temp, a, burnin, iterations, std::move(output)};
        c.wilson = true;
        c.width = width;
        c.height = height;
        return c;
    }
    std::string to_string() const {  // computation.rs:32-52
        std::ostringstream s;
        if (wilson) s << width << "x" << height << " Wilson ";
        s << algorithm.to_string() << " Simulation";
        return s.str();
    }

    // Compute::run (computation.rs:221-281, 304-350): the device handle lives for the whole run.  The burn-in is one
    // lqft_run call; the measurement phase is ONE lqft_run_observables call in which the first output of each kind
    // is measured and logged on the device at its own frequency (computation.rs:262-265).  Only a second output of
    // the same kind (never used by the reference's drivers) falls back to a host loop between calls.
    Computation run() && {
        auto t0 = std::chrono::steady_clock::now();
        fields::Field field = fields::Field::random(*lattice, seed, device);
        lqft_handle* h = field.gpu->raw();
        check(lqft_set_chain(h, 0, temp, seed, wilson ? width : 0, wilson ? height : 0));
        lqft_schedule s{};
        s.sweep_base = 0; s.first_step = 0; s.n_steps = burnin; s.normalize_every = 10;
        if (burnin) check(lqft_run(h, &s, nullptr, nullptr, nullptr));
        auto freq = [&](const OutputData& o) { return wilson ? (uint64_t)1 : o.frequency; };  // quirk Q4
        OutputData* fused[4] = {nullptr, nullptr, nullptr, nullptr};
        std::vector<OutputData*> others;
        for (auto& o : output) {
            OutputData*& slot = fused[(int)o.kind];
            if (!slot) slot = &o; else others.push_back(&o);
        }
        const uint32_t T = lattice->size[2];
        uint64_t step = 0;
        while (step < iterations) {
            uint64_t nxt = iterations - 1;
            for (auto* o : others) nxt = std::min(nxt, (step + freq(*o) - 1) / freq(*o) * freq(*o));
            const uint64_t n = nxt - step + 1;
            s = lqft_schedule{};
            s.sweep_base = burnin; s.first_step = step; s.n_steps = n; s.normalize_every = 10;
            OutputData* fe = fused[(int)Kind::EnergyObservable];
            OutputData* fa = fused[(int)Kind::ActionObservable];
            OutputData* fc = fused[(int)Kind::CorrelationData];
            OutputData* fd = fused[(int)Kind::DifferencePlot];
            if (fe) s.energy_every = (uint32_t)freq(*fe);
            if (fa) s.action_every = (uint32_t)freq(*fa);
            if (fc) { s.correlator_every = (uint32_t)freq(*fc); s.correlator_reps = (uint32_t)fc->repetitions; }
            if (fd) s.difference_every = (uint32_t)freq(*fd);
            const uint64_t ne = lqft_run_measurements(step, n, s.energy_every), na = lqft_run_measurements(step, n, s.action_every),
                           ncr = lqft_run_measurements(step, n, s.correlator_every);
            std::vector<double> e(ne), a(na), c(ncr * T);
            lqft_run_outputs out{};
            out.e_obs = ne ? e.data() : nullptr;
            out.s_obs = na ? a.data() : nullptr;
            out.corr = ncr ? c.data() : nullptr;
            check(lqft_run_observables(h, &s, &out));
            if (fe) fe->values.insert(fe->values.end(), e.begin(), e.end());
            if (fa) fa->values.insert(fa->values.end(), a.begin(), a.end());
            for (uint64_t m = 0; m < ncr; ++m) fc->corr_fn.emplace_back(c.begin() + m * T, c.begin() + (m + 1) * T);
            for (auto* o : others)
                if (nxt % freq(*o) == 0) o->update(field, burnin + nxt);
            step = nxt + 1;
        }
        for (auto& o : output) o.finish(field);
        duration = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
        return std::move(*this);
    }
};

// parse_simulation_results (computation.rs:129-199): same files, same columns, same number format.
inline void parse_simulation_results(std::vector<Computation>& data, const std::string& root = ".") {
    const std::string results = root + "/data/results.csv";
    const char* header = "index,d,size,x,y,t,temp,comptype,burnin,iterations,duration,algorithm_statistic,action_data,"
                         "energy_data,difference_data,correlation_data";
    for (auto& comp : data) {
        uint64_t index = 0;
        {
            std::ifstream in(results);
            std::string line, last;
            bool first = true;
            while (std::getline(in, line)) {
                if (first) { first = false; continue; }
                if (!line.empty()) last = line;
            }
            if (!last.empty()) index = std::stoull(last.substr(0, last.find(','))) + 1;
        }
        bool action = false, energy = false, diff = false, corr = false;
        for (auto& o : comp.output) {
            const std::string idx = std::to_string(index);
            if (o.kind == Kind::EnergyObservable || o.kind == Kind::ActionObservable) {
                const bool is_e = o.kind == Kind::EnergyObservable;
                std::ofstream f(root + (is_e ? "/data/energy_data/energy_" : "/data/action_data/action_") + idx + ".csv",
                                std::ios::app);
                for (double v : o.values) f << export_::fmt_f64(v) << "\n";
                (is_e ? energy : action) = true;
            } else if (o.kind == Kind::CorrelationData) {
                std::ofstream f(root + "/data/correlation_data/correlation_" + idx + ".csv");
                for (auto& row : o.corr_fn) {
                    for (size_t i = 0; i < row.size(); ++i) f << (i ? "," : "") << export_::fmt_f64(row[i]);
                    f << "\n";
                }
                corr = true;
            } else {
                const uint64_t n = comp.lattice->n_sites();
                for (int dir = 0; dir < 3; ++dir) {
                    std::ofstream f(root + "/data/plot_data/difference_" + idx + "_" + std::to_string(dir) + ".csv");
                    f << "x,y,t,value\n";
                    for (uint64_t i = 0; i < n; ++i) {
                        auto c = comp.lattice->calc_coords_from_index(i);
                        f << c[0] << "," << c[1] << "," << c[2] << "," << export_::fmt_f64(o.difference_mean[dir * n + i]) << "\n";
                    }
                }
                diff = true;
            }
        }
        std::ifstream probe(results);
        const bool fresh = !probe.good() || probe.peek() == std::ifstream::traits_type::eof();
        probe.close();
        std::ofstream out(results, std::ios::app);
        if (fresh) out << header << "\n";
        auto b = [](bool v) { return v ? "true" : "false"; };
        out << index << ",3," << comp.lattice->n_sites() << "," << comp.lattice->size[0] << "," << comp.lattice->size[1] << ","
            << comp.lattice->size[2] << "," << export_::fmt_f64(comp.temp) << "," << comp.to_string() << "," << comp.burnin << ","
            << comp.iterations << "," << (comp.wilson || comp.duration < 0 ? std::string() : std::to_string(comp.duration))
            << ",," << b(action) << "," << b(energy) << "," << b(diff) << "," << b(corr) << "\n";
    }
}
}  // namespace computation
}  // namespace lattice_qft

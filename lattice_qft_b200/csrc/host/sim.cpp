// sim.cpp — src/bin/sim.rs written against the C++ host mirror.  The reference's driver is configured
// by compile-time constants (sim.rs:16-38); the same quantities are command-line arguments here:
//   lqft_sim CUBE BURNIN ITERATIONS [--wilson] [--energy FREQ] [--corr REPS FREQ] [--temps a,b,c] [--root DIR]
// One chain per coupling of TEMP_ARY (default INVESTIGATE_ARY9, lib.rs:50), then parse_simulation_results.
#include <array>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>

#include "lattice_qft.hpp"

using namespace lattice_qft;

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cerr << "usage: lqft_sim CUBE BURNIN ITERATIONS [--wilson] [--energy FREQ] [--corr REPS FREQ] [--temps a,b] [--root DIR]\n";
        return 2;
    }
    const uint32_t cube = std::atoi(argv[1]);
    const uint64_t burnin = std::strtoull(argv[2], nullptr, 10), iterations = std::strtoull(argv[3], nullptr, 10);
    std::vector<double> temps = {0.2, 0.21, 0.22, 0.23, 0.24, 0.25, 0.26, 0.27, 0.28, 0.29};  // INVESTIGATE_ARY9
    bool wilson = false;
    uint64_t energy_freq = 10, corr_reps = 0, corr_freq = 10;
    std::string root = ".";
    for (int i = 4; i < argc; ++i) {
        if (!strcmp(argv[i], "--wilson")) wilson = true;
        else if (!strcmp(argv[i], "--energy") && i + 1 < argc) energy_freq = std::strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--corr") && i + 2 < argc) { corr_reps = std::strtoull(argv[++i], nullptr, 10); corr_freq = std::strtoull(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "--root") && i + 1 < argc) root = argv[++i];
        else if (!strcmp(argv[i], "--temps") && i + 1 < argc) {
            temps.clear();
            std::stringstream ss(argv[++i]);
            std::string tok;
            while (std::getline(ss, tok, ',')) temps.push_back(std::atof(tok.c_str()));
        }
    }
    for (const char* d : {"/data/energy_data", "/data/action_data", "/data/correlation_data", "/data/plot_data"})
        std::filesystem::create_directories(root + d);
    try {
        lattice::Lattice3d lattice(cube, cube, cube);  // sim.rs:48
        std::vector<computation::Computation> comps;
        uint64_t seed = 0xC0FFEE;
        for (double temp : temps) {  // sim.rs:56-87
            std::vector<outputdata::OutputData> observables;
            if (energy_freq) observables.push_back(outputdata::OutputData::new_energy_observable(temp).set_frequency(energy_freq));
            if (corr_reps) observables.push_back(outputdata::OutputData::new_correlation_plot(lattice, corr_reps).set_frequency(corr_freq));
            if (wilson) {
                for (uint32_t width = 1; width <= 5; ++width) {  // the commented block of sim.rs:62-76 (all five widths)
                    auto c = computation::Computation::new_wilson_sim(lattice, temp, algorithm::AlgorithmType::new_metropolis(),
                                                                      burnin, iterations, cube * width / 6, cube, observables);
                    c.seed = seed++;
                    comps.push_back(std::move(c));
                }
            }
            auto c = computation::Computation::new_simulation(lattice, temp, algorithm::AlgorithmType::new_metropolis(), burnin,
                                                              iterations, observables);
            c.seed = seed++;
            comps.push_back(std::move(c));
        }
        std::vector<computation::Computation> data;
        for (auto& c : comps) {  // sim.rs:90-93 (rayon there; the chains are independent)
            try {
                data.push_back(std::move(c).run());
                std::cout << export_::fmt_f64(data.back().temp) << " " << data.back().to_string() << " took " << data.back().duration
                          << " secs" << std::endl;
            } catch (const Error& e) {
                std::cerr << "dropped: " << e.what() << std::endl;  // .filter_map(|c| c.run().ok())
            }
        }
        computation::parse_simulation_results(data, root);
    } catch (const Error& e) {
        std::cerr << "lqft error " << e.status << ": " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

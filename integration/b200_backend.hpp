// Shared by the two replacement translation units (NaiveAlgorithm_b200.cpp, BarnesHutAlgorithm_b200.cpp): the time
// loop of the reference's startSimulation (NaiveAlgorithm.cpp:82-259 == BarnesHutAlgorithm.cpp:102-276), expressed
// with the C ABI of include/nbody_b200.h and filling the reference's OWN nBodyAlgorithm members (snapshot maps,
// energies, TimeMeasurement), so that the reference's unmodified main.cpp, InputParser, TimeMeasurement and
// generateParaViewOutput keep working on top of it.  This is the code a maintainer of the reference would add.
#pragma once

#include "nBodyAlgorithm.hpp"      // the reference's header, unmodified
#include "SimulationData.hpp"
#include "Configuration.hpp"

#include <nbody_b200.h>

#include <cmath>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <string>

namespace b200 {

inline void check(nb_ctx *ctx, int status, const char *what) {
    if (status != NB_OK)
        throw std::runtime_error(std::string(what) + ": " + (ctx ? nb_last_error(ctx) : nb_status_string(status)));
}

// the configuration:: globals the reference's main.cpp has filled -> nb_config (Configuration.hpp:12-90)
inline nb_ctx *open_context(const nBodyAlgorithm &alg) {
    nb_config cfg;
    nb_config_default(&cfg);
    cfg.G = alg.G;                                                        // nBodyAlgorithm.hpp:55-61
    cfg.epsilon2 = configuration::epsilon2;
    cfg.block_size = configuration::naive_algorithm::blockSize;
    cfg.opt_stage = configuration::naive_algorithm::optimization_stage;
    cfg.theta = configuration::barnes_hut_algorithm::theta;
    cfg.sort_bodies = configuration::barnes_hut_algorithm::sortBodies ? 1 : 0;
    cfg.wg_size_barnes_hut = configuration::barnes_hut_algorithm::workGroupSize;
    const d_type::int_t n = configuration::numberOfBodies;
    if (n) cfg.storage_size_param = (int) (configuration::barnes_hut_algorithm::storageSizeParameter / n);
    nb_ctx *ctx = nullptr;
    check(nullptr, nb_create(&cfg, &ctx), "nb_create");
    return ctx;
}

// forces(): one force evaluation on the device (nb_naive_accel, or nb_bh_build + nb_bh_accel) + its timer entries
inline void run_time_loop(nBodyAlgorithm &a, nb_ctx *ctx, const SimulationData &d, const std::function<void()> &forces) {
    const std::size_t n = d.mass.size();
    check(ctx, nb_set_bodies(ctx, n, d.mass.data(), d.positions_x.data(), d.positions_y.data(), d.positions_z.data(),
                             d.velocities_x.data(), d.velocities_y.data(), d.velocities_z.data()), "nb_set_bodies");
    check(ctx, nb_enable_timers(ctx, 1), "nb_enable_timers");
    char name[256] = "";
    check(ctx, nb_device_name(ctx, name, sizeof name), "nb_device_name");
    std::string device = name;
    a.timer.setProperties(a.description, configuration::numberOfBodies, device);
    a.timer.addTimingSequence("Leapfrog Part 1");
    a.timer.addTimingSequence("Leapfrog Part 2");

    auto store_accelerations = [&](d_type::int_t step) {                 // nBodyAlgorithm::storeAccelerations
        a.acceleration[step].resize(n);
        check(ctx, nb_get_acceleration_norms(ctx, a.acceleration[step].data()), "nb_get_acceleration_norms");
    };
    auto compute_energy = [&](d_type::int_t step) {                      // nBodyAlgorithm::computeEnergy
        double e[4];
        check(ctx, nb_energy(ctx, e), "nb_energy");
        a.kineticEnergy[step] = e[0];
        a.potentialEnergy[step] = e[1];
        a.totalEnergy[step] = e[2];
        a.virialEquilibrium[step] = e[3];
    };

    // step 0 of the output: the input state, velocities shifted by the reference's own adjustVelocities (output only)
    a.positions_x[0] = d.positions_x; a.positions_y[0] = d.positions_y; a.positions_z[0] = d.positions_z;
    a.velocities_x[0] = d.velocities_x; a.velocities_y[0] = d.velocities_y; a.velocities_z[0] = d.velocities_z;
    a.adjustVelocities(d);

    double time = 0.0, sinceLastVisualization = 0.0;
    d_type::int_t step = 0;
    forces();
    if (configuration::compute_energy) compute_energy(step);
    store_accelerations(step);
    std::cout << "Finished initial step " << step << std::endl << std::endl;
    time += a.dt;
    sinceLastVisualization += a.dt;
    step += 1;

    double ms[NB_T_COUNT];
    while (time <= a.t_end + 0.000001) {
        const bool visualize = std::abs(sinceLastVisualization - a.visualizationStepWidth) < 0.000001;
        check(ctx, nb_leapfrog_part1(ctx, a.dt), "nb_leapfrog_part1");
        if (visualize) {
            a.positions_x[step].resize(n); a.positions_y[step].resize(n); a.positions_z[step].resize(n);
            check(ctx, nb_get_positions(ctx, a.positions_x[step].data(), a.positions_y[step].data(),
                                        a.positions_z[step].data()), "nb_get_positions");
        }
        forces();
        check(ctx, nb_leapfrog_part2(ctx, a.dt), "nb_leapfrog_part2");
        check(ctx, nb_get_timers(ctx, ms), "nb_get_timers");
        a.timer.addTimeToSequence("Leapfrog Part 1", ms[NB_T_LEAPFROG1]);
        a.timer.addTimeToSequence("Leapfrog Part 2", ms[NB_T_LEAPFROG2]);
        if (visualize) {
            std::cout << "Finished step " << step << std::endl << std::endl;
            store_accelerations(step);
            a.velocities_x[step].resize(n); a.velocities_y[step].resize(n); a.velocities_z[step].resize(n);
            check(ctx, nb_get_velocities(ctx, a.velocities_x[step].data(), a.velocities_y[step].data(),
                                         a.velocities_z[step].data()), "nb_get_velocities");
            if (configuration::compute_energy) compute_energy(step);
            step += 1;
            sinceLastVisualization = 0.0;
        }
        time += a.dt;
        sinceLastVisualization += a.dt;
    }
    check(ctx, nb_synchronize(ctx), "nb_synchronize");
}

}  // namespace b200

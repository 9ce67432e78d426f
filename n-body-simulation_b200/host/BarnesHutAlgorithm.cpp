#include "BarnesHutAlgorithm.hpp"

#include <stdexcept>

namespace {
void checked(nb_ctx *ctx, int status, const char *what) {
    if (status == NB_OK) return;
    throw std::runtime_error(std::string(what) + ": " + nb_status_string(status) + " (" + (ctx ? nb_last_error(ctx) : "") + ")");
}
}  // namespace

void BarnesHutOctree::computeMinMaxValuesAABB(double out[7]) {
    checked(ctx, nb_bh_aabb(ctx, out), "nb_bh_aabb");
    min_x = out[0]; min_y = out[1]; min_z = out[2];
    max_x = out[3]; max_y = out[4]; max_z = out[5];
    AABB_EdgeLength = out[6];
}

void BarnesHutOctree::buildOctree(TimeMeasurement &) { checked(ctx, nb_bh_build(ctx), "nb_bh_build"); }

BarnesHutAlgorithm::BarnesHutAlgorithm(double dt, double tEnd, double visualizationStepWidth, std::string &outputDirectory)
    : nBodyAlgorithm(dt, tEnd, visualizationStepWidth, outputDirectory), octree(ctx) {
    description = "Barnes-Hut Algorithm";
}

void BarnesHutAlgorithm::computeAccelerations() { check(nb_bh_accel(ctx), "nb_bh_accel"); }

void BarnesHutAlgorithm::startSimulation(const SimulationData &simulationData) {
    openDevice(simulationData);
    // same sequence names as the reference's times.json (BarnesHutAlgorithm.cpp:79-100); the phases of the lock-based
    // builder map onto the sort-based one: "Sort bodies for subtrees" = keys + radix sort, "Build subtrees" = node
    // emission.  "Build octree to level" / "Prepare subtrees" / "Sort bodies" have no counterpart (the in-order
    // permutation falls out of the radix sort): they are recorded as 0 ms per step so that times.json carries the same
    // keys and the same number of entries as the reference's file (ParallelOctreeTopDownSubtrees.cpp:76-92).
    for (const char *name : {"Total Time", "Octree creation", "Acceleration Kernel Time", "AABB creation",
                             "Compute center of mass", "Build octree to level", "Prepare subtrees",
                             "Sort bodies for subtrees", "Build subtrees"})
        timer.addTimingSequence(name);
    const bool sorted = configuration::barnes_hut_algorithm::sortBodies;
    if (sorted) timer.addTimingSequence("Sort bodies");
    batchAlgorithm = 1;
    recordForceTimers = [this, sorted](const double *ms) {
        timer.addTimeToSequence("Octree creation", ms[NB_T_TREE_TOTAL]);
        timer.addTimeToSequence("Acceleration Kernel Time", ms[NB_T_ACCEL]);
        timer.addTimeToSequence("Total Time", ms[NB_T_TREE_TOTAL] + ms[NB_T_ACCEL]);
        timer.addTimeToSequence("AABB creation", ms[NB_T_AABB]);
        timer.addTimeToSequence("Sort bodies for subtrees", ms[NB_T_KEYS_SORT]);
        timer.addTimeToSequence("Build subtrees", ms[NB_T_BUILD]);
        timer.addTimeToSequence("Compute center of mass", ms[NB_T_COM]);
        timer.addTimeToSequence("Build octree to level", 0.0);
        timer.addTimeToSequence("Prepare subtrees", 0.0);
        if (sorted) timer.addTimeToSequence("Sort bodies", 0.0);
    };
    runTimeLoop(simulationData, [this]() {
        octree.buildOctree(timer);
        computeAccelerations();
        double ms[NB_T_COUNT];
        check(nb_get_timers(ctx, ms), "nb_get_timers");
        recordForceTimers(ms);
    });
}

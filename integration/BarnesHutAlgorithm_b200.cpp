// Drop-in replacement for the reference's src/simulationBackend/BarnesHutAlgorithm.cpp.
//
// Compiled against the reference's UNMODIFIED BarnesHutAlgorithm.hpp.  The class keeps its `Octree octree` and
// `nodesOnStack` members because the header declares them (a maintainer would delete both: the CUDA library owns the
// tree and needs no per-body stack), but nothing reads them.
//   BarnesHutAlgorithm::startSimulation       -> time loop over nb_bh_build + nb_bh_accel / nb_leapfrog_part{1,2}
//   BarnesHutAlgorithm::computeAccelerations(queue, masses, pos x/y/z, acc x/y/z)
//                                             -> nb_op_barnes_hut_accelerations (build + traversal, host arrays)
#include "BarnesHutAlgorithm.hpp"  // the reference's header, unmodified
#include "b200_backend.hpp"

BarnesHutAlgorithm::BarnesHutAlgorithm(double dt, double tEnd, double visualizationStepWidth,
                                       std::string &outputDirectory)
        : nBodyAlgorithm(dt, tEnd, visualizationStepWidth, outputDirectory),
          nodesOnStack_vec(1, 0),                                   // no traversal stack (BarnesHutAlgorithm.cpp:8-15)
          nodesOnStack(nodesOnStack_vec.data(), nodesOnStack_vec.size()) {
    this->description = "Barnes-Hut Algorithm";
}

void BarnesHutAlgorithm::startSimulation(const SimulationData &simulationData) {
    nb_ctx *ctx = b200::open_context(*this);
    // the sequence names of the reference's times.json (BarnesHutAlgorithm.cpp:79-100)
    for (const char *name: {"Total Time", "Octree creation", "Acceleration Kernel Time", "AABB creation",
                            "Compute center of mass", "Build octree to level", "Prepare subtrees",
                            "Sort bodies for subtrees", "Build subtrees"})
        timer.addTimingSequence(name);
    const bool sorted = configuration::barnes_hut_algorithm::sortBodies;
    if (sorted) timer.addTimingSequence("Sort bodies");
    b200::run_time_loop(*this, ctx, simulationData, [&]() {
        b200::check(ctx, nb_bh_build(ctx), "nb_bh_build");          // octree.buildOctree(...)      (:106, :202)
        b200::check(ctx, nb_bh_accel(ctx), "nb_bh_accel");          // computeAccelerations(...)    (:112, :207)
        double ms[NB_T_COUNT];
        b200::check(ctx, nb_get_timers(ctx, ms), "nb_get_timers");
        timer.addTimeToSequence("Octree creation", ms[NB_T_TREE_TOTAL]);
        timer.addTimeToSequence("Acceleration Kernel Time", ms[NB_T_ACCEL]);
        timer.addTimeToSequence("Total Time", ms[NB_T_TREE_TOTAL] + ms[NB_T_ACCEL]);
        timer.addTimeToSequence("AABB creation", ms[NB_T_AABB]);
        timer.addTimeToSequence("Sort bodies for subtrees", ms[NB_T_KEYS_SORT]);
        timer.addTimeToSequence("Build subtrees", ms[NB_T_BUILD]);
        timer.addTimeToSequence("Compute center of mass", ms[NB_T_COM]);
        timer.addTimeToSequence("Build octree to level", 0.0);       // no counterpart in the sort-based build
        timer.addTimeToSequence("Prepare subtrees", 0.0);
        if (sorted) timer.addTimeToSequence("Sort bodies", 0.0);     // the permutation falls out of the radix sort
    });
    nb_destroy(ctx);
}

void BarnesHutAlgorithm::computeAccelerations(queue &, buffer<double> &masses, buffer<double> &currentPositions_x,
                                              buffer<double> &currentPositions_y, buffer<double> &currentPositions_z,
                                              buffer<double> &acceleration_x, buffer<double> &acceleration_y,
                                              buffer<double> &acceleration_z) {
    nb_ctx *ctx = b200::open_context(*this);
    host_accessor<double> M(masses), X(currentPositions_x), Y(currentPositions_y), Z(currentPositions_z),
            AX(acceleration_x), AY(acceleration_y), AZ(acceleration_z);
    const int rc = nb_op_barnes_hut_accelerations(ctx, masses.size(), &M[0], &X[0], &Y[0], &Z[0], &AX[0], &AY[0], &AZ[0]);
    const std::string err = rc == NB_OK ? "" : nb_last_error(ctx);
    nb_destroy(ctx);
    if (rc != NB_OK) throw std::runtime_error("nb_op_barnes_hut_accelerations: " + err);
}

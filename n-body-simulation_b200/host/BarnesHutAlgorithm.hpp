// Barnes-Hut back end (reference src/simulationBackend/BarnesHutAlgorithm.hpp:13-49).  The reference owns an Octree
// member (20 SoA buffers of 16N nodes) and a stackSize*N traversal stack; here the tree lives inside the nb_ctx
// (DFS pre-order node array, no stack) and `octree` is a thin handle exposing the same operations.
#pragma once
#include "nBodyAlgorithm.hpp"

// mirrors BarnesHutOctree's public operations (reference BarnesHutOctree.hpp:102-139) on the context's tree
class BarnesHutOctree {
public:
    explicit BarnesHutOctree(nb_ctx *&ctxRef) : ctx(ctxRef) {}
    // AABB -> keys/sort -> node construction -> centre of mass (-> body order), recording the per-phase timers
    void buildOctree(TimeMeasurement &timer);
    // min xyz, max xyz, edge of the cube around all bodies and the origin
    void computeMinMaxValuesAABB(double out[7]);
    double min_x = 0, min_y = 0, min_z = 0, max_x = 0, max_y = 0, max_z = 0, AABB_EdgeLength = 0;

private:
    nb_ctx *&ctx;
};

class BarnesHutAlgorithm : public nBodyAlgorithm {
public:
    BarnesHutAlgorithm(double dt, double tEnd, double visualizationStepWidth, std::string &outputDirectory);

    void startSimulation(const SimulationData &simulationData) override;
    void computeAccelerations();

    BarnesHutOctree octree;
};

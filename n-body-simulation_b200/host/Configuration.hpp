// Process-wide run configuration, same names and defaults as the reference's `namespace configuration`
// (reference src/utility/Configuration.hpp:12-124, Configuration.cpp:5-33).  No SYCL dependency: the values are
// copied into an nb_config (include/nbody_b200.h) when an algorithm object creates its device context.
#pragma once
#include <cstdint>

#include "nbody_b200.h"

namespace d_type {
typedef unsigned int int_t;
}

namespace configuration {
extern d_type::int_t numberOfBodies;
extern double epsilon2;
extern bool compute_energy;
extern bool use_GPUs;

namespace naive_algorithm {
extern int blockSize;
extern int optimization_stage;
}  // namespace naive_algorithm

namespace barnes_hut_algorithm {
extern double theta;
extern int workGroupSize;
extern d_type::int_t stackSize;
extern d_type::int_t storageSizeParameter;
extern int AABBWorkItemCount;
extern int octreeWorkItemCount;
extern int octreeTopWorkItemCount;
extern int centerOfMassWorkItemCount;
extern int maxBuildLevel;
extern bool sortBodies;
}  // namespace barnes_hut_algorithm

// raw --storage_size_param / --stack_size_param values (the reference only keeps the products with N)
extern int storageSizeParamRaw;
extern int stackSizeParamRaw;
// multi-GPU (new): filled from WORLD_SIZE / RANK / LOCAL_RANK when launched with one process per GPU
extern int worldSize;
extern int rank;
extern int localRank;

void initializeConfigValues(d_type::int_t bodyCount, int storageSizeParam, int stackSizeParam);

void setBlockSize(int blockSize);
void setTheta(double theta);
void setAABBWorkItemCount(int workItemCount);
void setOctreeWorkItemCount(int workItemCount);
void setOctreeTopWorkItemCount(int workItemCount);
void setCenterOfMassWorkItemCount(int workItemCount);
void setMaxBuildLevel(int maxLevel);
void setEnergyComputation(bool computeEnergy);
void setSortBodies(bool sort_bodies);
void setDeviceGPU(bool useGPU);
void setWorkGroupSizeBarnesHut(int workGroupSize);
void setOptimizationStage(int stage);

// snapshot of the globals as the plain-old-data struct the C ABI takes
nb_config toDeviceConfig(double G);
}  // namespace configuration

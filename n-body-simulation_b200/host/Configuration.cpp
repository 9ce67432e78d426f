#include "Configuration.hpp"

#include <cmath>

namespace configuration {

d_type::int_t numberOfBodies = 0;
double epsilon2 = std::pow(10, -22);
bool compute_energy = false;
bool use_GPUs = true;

namespace naive_algorithm {
int blockSize = 64;
int optimization_stage = 2;
}  // namespace naive_algorithm

namespace barnes_hut_algorithm {
d_type::int_t storageSizeParameter = 0;
d_type::int_t stackSize = 0;
int AABBWorkItemCount = 1024;
int octreeWorkItemCount = 640;
int octreeTopWorkItemCount = 1024;
int centerOfMassWorkItemCount = 1024;
double theta = 1.05;
int maxBuildLevel = 7;
bool sortBodies = true;
int workGroupSize = 64;
}  // namespace barnes_hut_algorithm

int storageSizeParamRaw = 16;
int stackSizeParamRaw = 16;
int worldSize = 1;
int rank = 0;
int localRank = 0;

void initializeConfigValues(d_type::int_t bodyCount, int storageSizeParam, int stackSizeParam) {
    namespace bh = barnes_hut_algorithm;
    numberOfBodies = bodyCount;
    storageSizeParamRaw = storageSizeParam;
    stackSizeParamRaw = stackSizeParam;
    bh::storageSizeParameter = storageSizeParam * numberOfBodies;
    // the reference enlarges the per-body stack for small systems; kept for the config echo, the traversal here is
    // stackless
    const d_type::int_t levels = (d_type::int_t) std::ceil(std::log2(bodyCount));
    bh::stackSize = stackSizeParam * levels + (bodyCount < 15000 ? 500 : 0);
}

void setBlockSize(int v) { naive_algorithm::blockSize = v; }
void setTheta(double v) { barnes_hut_algorithm::theta = v; }
void setAABBWorkItemCount(int v) { barnes_hut_algorithm::AABBWorkItemCount = v; }
void setOctreeWorkItemCount(int v) { barnes_hut_algorithm::octreeWorkItemCount = v; }
void setOctreeTopWorkItemCount(int v) { barnes_hut_algorithm::octreeTopWorkItemCount = v; }
void setCenterOfMassWorkItemCount(int v) { barnes_hut_algorithm::centerOfMassWorkItemCount = v; }
void setMaxBuildLevel(int v) { barnes_hut_algorithm::maxBuildLevel = v; }
void setEnergyComputation(bool v) { compute_energy = v; }
void setSortBodies(bool v) { barnes_hut_algorithm::sortBodies = v; }
void setDeviceGPU(bool v) { use_GPUs = v; }
void setWorkGroupSizeBarnesHut(int v) { barnes_hut_algorithm::workGroupSize = v; }
void setOptimizationStage(int v) { naive_algorithm::optimization_stage = v; }

nb_config toDeviceConfig(double G) {
    nb_config c;
    nb_config_default(&c);
    c.device = localRank;
    c.G = G;
    c.epsilon2 = epsilon2;
    c.theta = barnes_hut_algorithm::theta;
    c.block_size = naive_algorithm::blockSize;
    c.opt_stage = naive_algorithm::optimization_stage;
    c.sort_bodies = barnes_hut_algorithm::sortBodies ? 1 : 0;
    c.wg_size_barnes_hut = barnes_hut_algorithm::workGroupSize;
    c.storage_size_param = storageSizeParamRaw;
    c.stack_size_param = stackSizeParamRaw;
    c.num_wi_aabb = barnes_hut_algorithm::AABBWorkItemCount;
    c.num_wi_octree = barnes_hut_algorithm::octreeWorkItemCount;
    c.num_wi_top_octree = barnes_hut_algorithm::octreeTopWorkItemCount;
    c.num_wi_com = barnes_hut_algorithm::centerOfMassWorkItemCount;
    c.max_level_top_octree = barnes_hut_algorithm::maxBuildLevel;
    c.world_size = worldSize;
    c.rank = rank;
    return c;
}

}  // namespace configuration

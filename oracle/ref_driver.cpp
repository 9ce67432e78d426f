// C entry points over the reference's OWN classes -- TEST INFRASTRUCTURE (see oracle/README.md).
//
// This file is linked with the UNMODIFIED reference sources compiled where they lie under /root/reference (g++, through
// the host SYCL subset in oracle/sycl_shim) into oracle/_ref/libnbody_ref.so.  Every function below only marshals
// plain arrays into the reference's sycl::buffer arguments and calls the reference's own operator:
//   ref_naive_accel        -> NaiveAlgorithm::computeAccelerations_opt_{0,1,2}      (NaiveAlgorithm.cpp:262-482)
//   ref_energy             -> nBodyAlgorithm::computeEnergy                        (nBodyAlgorithm.cpp:11-86)
//   ref_tree_*             -> {ParallelOctreeTopDownSubtrees,ParallelOctreeTopDownSynchronized}::buildOctree
//                             (AABB, build, centre of mass, in-order sort)
//   ref_bh_accel           -> BarnesHutAlgorithm::computeAccelerations             (BarnesHutAlgorithm.cpp:280-401)
//   ref_sim_*              -> {Naive,BarnesHut}Algorithm::startSimulation + the snapshot maps + generateParaViewOutput
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs load this library; the product never does.
// every library header the reference's headers pull in comes first, so that the access override below only touches
// the reference's own class definitions
#include <sycl/sycl.hpp>
#include <array>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#define private public      // the naive kernels are private members (NaiveAlgorithm.hpp:22); the driver calls them directly
#define protected public
#include "NaiveAlgorithm.hpp"
#include "BarnesHutAlgorithm.hpp"
#include "ParallelOctreeTopDownSubtrees.hpp"
#include "ParallelOctreeTopDownSynchronized.hpp"
#include "SimulationData.hpp"
#include "Configuration.hpp"
#include "InputParser.hpp"
#include "TimeConverter.hpp"
#undef private
#undef protected

#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

using int_t = d_type::int_t;

namespace {

std::string g_error;

struct CoutSilencer {
    std::streambuf *old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

template<class F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown exception";
        return 1;
    }
}

struct RefTree {
    int builder;
    int_t N;
    std::unique_ptr<BarnesHutOctree> octree;
    std::vector<double> m, x, y, z;
};

struct RefSim {
    std::unique_ptr<nBodyAlgorithm> alg;
    SimulationData data;
    std::string outdir;
};

void canon_rec(const BarnesHutOctree &t, int_t N, size_t S, int_t node, int depth, uint64_t hi, uint64_t lo,
               int_t *o_depth, uint64_t *o_hi, uint64_t *o_lo, int_t *o_kind, int_t *o_body, int_t *o_count,
               double *o_edge, double *o_minx, double *o_miny, double *o_minz, double *o_mass, double *o_comx,
               double *o_comy, double *o_comz, size_t &k) {
    o_depth[k] = depth; o_hi[k] = hi; o_lo[k] = lo;
    const bool leaf = t.nodeIsLeaf_vec[node] != 0;
    o_kind[k] = leaf ? (t.bodyOfNode_vec[node] != N ? 1 : 0) : 2;
    o_body[k] = t.bodyOfNode_vec[node];
    o_count[k] = t.bodyCountNode_vec[node];
    o_edge[k] = t.edgeLengths_vec[node];
    o_minx[k] = t.min_x_values_vec[node]; o_miny[k] = t.min_y_values_vec[node]; o_minz[k] = t.min_z_values_vec[node];
    o_mass[k] = t.sumMasses_vec[node];
    o_comx[k] = t.centerOfMass_x_vec[node]; o_comy[k] = t.centerOfMass_y_vec[node]; o_comz[k] = t.centerOfMass_z_vec[node];
    ++k;
    if (!leaf) {
        for (uint64_t oc = 0; oc < 8; ++oc) {
            uint64_t h2 = hi, l2 = lo;
            if (depth < 21) h2 |= oc << (60 - 3 * depth);
            else if (depth < 42) l2 |= oc << (60 - 3 * (depth - 21));
            canon_rec(t, N, S, t.octants_vec[(size_t) oc * S + node], depth + 1, h2, l2, o_depth, o_hi, o_lo, o_kind,
                      o_body, o_count, o_edge, o_minx, o_miny, o_minz, o_mass, o_comx, o_comy, o_comz, k);
        }
    }
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }
int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// main.cpp:122-159: the same calls the reference's driver makes after parsing its flags (<= 0 keeps the default)
void ref_configure(uint32_t N, int storage_param, int stack_param, double theta, int block_size, int opt_stage,
                   int sort_bodies, int wg_size_barnes_hut, int num_wi_octree, int num_wi_top_octree, int num_wi_AABB,
                   int num_wi_com, int max_level_top_octree, int energy) {
    configuration::initializeConfigValues(N, storage_param > 0 ? storage_param : 16, stack_param > 0 ? stack_param : 16);
    if (theta > 0) configuration::setTheta(theta);
    if (block_size > 0) configuration::setBlockSize(block_size);
    if (opt_stage >= 0) configuration::setOptimizationStage(opt_stage);
    if (sort_bodies >= 0) configuration::setSortBodies(sort_bodies != 0);
    if (wg_size_barnes_hut > 0) configuration::setWorkGroupSizeBarnesHut(wg_size_barnes_hut);
    if (num_wi_octree > 0) configuration::setOctreeWorkItemCount(num_wi_octree);
    if (num_wi_top_octree > 0) configuration::setOctreeTopWorkItemCount(num_wi_top_octree);
    if (num_wi_AABB > 0) configuration::setAABBWorkItemCount(num_wi_AABB);
    if (num_wi_com > 0) configuration::setCenterOfMassWorkItemCount(num_wi_com);
    if (max_level_top_octree > 0) configuration::setMaxBuildLevel(max_level_top_octree);
    if (energy >= 0) configuration::setEnergyComputation(energy != 0);
    configuration::setDeviceGPU(false);
}

uint32_t ref_config_storage_size() { return configuration::barnes_hut_algorithm::storageSizeParameter; }
uint32_t ref_config_stack_size() { return configuration::barnes_hut_algorithm::stackSize; }
double ref_epsilon2() { return configuration::epsilon2; }

double ref_gravitational_constant() {
    std::string dir = ".";
    NaiveAlgorithm a(1, 1, 1, dir);   // nBodyAlgorithm.hpp:49-61 computes G in the constructor
    return a.G;
}

double ref_convert_to_earth_days(const char *text, int *failed) {
    double v = 0;
    *failed = guarded([&] { std::string s(text); v = TimeConverter::convertToEarthDays(s); });
    return v;
}

// ---- naive ---------------------------------------------------------------------------------------------------------------
int ref_naive_accel(int opt_stage, uint32_t N, const double *m, const double *x, const double *y, const double *z,
                    double *ax, double *ay, double *az) {
    return guarded([&] {
        std::string dir = ".";
        NaiveAlgorithm alg(1, 1, 1, dir);
        std::vector<double> M(m, m + N), X(x, x + N), Y(y, y + N), Z(z, z + N);
        sycl::queue q;
        sycl::buffer<double> bm(M.data(), M.size()), bx(X.data(), X.size()), by(Y.data(), Y.size()), bz(Z.data(), Z.size());
        sycl::buffer<double> bax(ax, N), bay(ay, N), baz(az, N);
        if (opt_stage == 2) alg.computeAccelerations_opt_2(q, bm, bx, by, bz, bax, bay, baz);
        else if (opt_stage == 1) alg.computeAccelerations_opt_1(q, bm, bx, by, bz, bax, bay, baz);
        else alg.computeAccelerations_opt_0(q, bm, bx, by, bz, bax, bay, baz);
    });
}

// out = {kinetic, potential, total, virial} of step 0 (nBodyAlgorithm.cpp:77-85)
int ref_energy(uint32_t N, const double *m, const double *x, const double *y, const double *z, const double *vx,
               const double *vy, const double *vz, double out[4]) {
    return guarded([&] {
        std::string dir = ".";
        NaiveAlgorithm alg(1, 1, 1, dir);
        std::vector<double> M(m, m + N), X(x, x + N), Y(y, y + N), Z(z, z + N), VX(vx, vx + N), VY(vy, vy + N), VZ(vz, vz + N);
        sycl::queue q;
        sycl::buffer<double> bm(M.data(), N), bx(X.data(), N), by(Y.data(), N), bz(Z.data(), N), bvx(VX.data(), N),
                bvy(VY.data(), N), bvz(VZ.data(), N);
        alg.computeEnergy(q, bm, 0, bx, by, bz, bvx, bvy, bvz);
        out[0] = alg.kineticEnergy[0];
        out[1] = alg.potentialEnergy[0];
        out[2] = alg.totalEnergy[0];
        out[3] = alg.virialEquilibrium[0];
    });
}

// ---- octree -------------------------------------------------------------------------------------------------------------
// builder 0 = ParallelOctreeTopDownSubtrees (the default), 1 = ParallelOctreeTopDownSynchronized.  ref_configure first.
void *ref_tree_create(int builder) {
    RefTree *t = nullptr;
    int rc = guarded([&] {
        t = new RefTree();
        t->builder = builder;
        t->N = configuration::numberOfBodies;
        if (builder == 1) t->octree.reset(new ParallelOctreeTopDownSynchronized());
        else t->octree.reset(new ParallelOctreeTopDownSubtrees());
    });
    return rc ? nullptr : t;
}

void ref_tree_destroy(void *h) { delete (RefTree *) h; }

int ref_tree_aabb(void *h, const double *x, const double *y, const double *z, double out[7]) {
    RefTree &t = *(RefTree *) h;
    return guarded([&] {
        t.x.assign(x, x + t.N); t.y.assign(y, y + t.N); t.z.assign(z, z + t.N);
        sycl::queue q;
        sycl::buffer<double> bx(t.x.data(), t.N), by(t.y.data(), t.N), bz(t.z.data(), t.N);
        t.octree->computeMinMaxValuesAABB(q, bx, by, bz);
        BarnesHutOctree &o = *t.octree;
        out[0] = o.min_x; out[1] = o.min_y; out[2] = o.min_z; out[3] = o.max_x; out[4] = o.max_y; out[5] = o.max_z;
        out[6] = o.AABB_EdgeLength;
    });
}

int ref_tree_build(void *h, const double *m, const double *x, const double *y, const double *z) {
    RefTree &t = *(RefTree *) h;
    return guarded([&] {
        t.m.assign(m, m + t.N); t.x.assign(x, x + t.N); t.y.assign(y, y + t.N); t.z.assign(z, z + t.N);
        sycl::queue q;
        sycl::buffer<double> bm(t.m.data(), t.N), bx(t.x.data(), t.N), by(t.y.data(), t.N), bz(t.z.data(), t.N);
        TimeMeasurement timer;
        t.octree->buildOctree(q, bx, by, bz, bm, timer);
    });
}

uint32_t ref_tree_num_nodes(void *h) { return ((RefTree *) h)->octree->nextFreeNodeID_vec[0]; }
void ref_tree_get_aabb(void *h, double out[7]) {
    BarnesHutOctree &o = *((RefTree *) h)->octree;
    out[0] = o.min_x; out[1] = o.min_y; out[2] = o.min_z; out[3] = o.max_x; out[4] = o.max_y; out[5] = o.max_z;
    out[6] = o.AABB_EdgeLength;
}
const uint32_t *ref_tree_body_of_node(void *h) { return ((RefTree *) h)->octree->bodyOfNode_vec.data(); }
const uint32_t *ref_tree_body_count(void *h) { return ((RefTree *) h)->octree->bodyCountNode_vec.data(); }
const uint32_t *ref_tree_octants(void *h) { return ((RefTree *) h)->octree->octants_vec.data(); }
const int *ref_tree_is_leaf(void *h) { return ((RefTree *) h)->octree->nodeIsLeaf_vec.data(); }
const double *ref_tree_sum_masses(void *h) { return ((RefTree *) h)->octree->sumMasses_vec.data(); }
const double *ref_tree_com_x(void *h) { return ((RefTree *) h)->octree->centerOfMass_x_vec.data(); }
const double *ref_tree_com_y(void *h) { return ((RefTree *) h)->octree->centerOfMass_y_vec.data(); }
const double *ref_tree_com_z(void *h) { return ((RefTree *) h)->octree->centerOfMass_z_vec.data(); }
const double *ref_tree_edge(void *h) { return ((RefTree *) h)->octree->edgeLengths_vec.data(); }
const uint32_t *ref_tree_sorted_bodies(void *h) { return ((RefTree *) h)->octree->sortedBodiesInOrder_vec.data(); }

// the same canonical (depth, path) record layout as orc_tree_canonical in nbody_oracle.cpp
void ref_tree_canonical(void *h, uint32_t *depth, uint64_t *path_hi, uint64_t *path_lo, uint32_t *kind, uint32_t *body,
                        uint32_t *count, double *edge, double *minx, double *miny, double *minz, double *mass,
                        double *comx, double *comy, double *comz) {
    RefTree &t = *(RefTree *) h;
    size_t k = 0;
    canon_rec(*t.octree, t.N, configuration::barnes_hut_algorithm::storageSizeParameter, 0, 0, 0, 0, depth, path_hi,
              path_lo, kind, body, count, edge, minx, miny, minz, mass, comx, comy, comz, k);
}

// ---- Barnes-Hut accelerations: buildOctree + computeAccelerations exactly as the time loop calls them -----------------
// (BarnesHutAlgorithm.cpp:106-112).  num_nodes receives nextFreeNodeID.
int ref_bh_accel(uint32_t N, const double *m, const double *x, const double *y, const double *z, double *ax, double *ay,
                 double *az, uint32_t *num_nodes) {
    return guarded([&] {
        std::string dir = ".";
        BarnesHutAlgorithm alg(1, 1, 1, dir);
        std::vector<double> M(m, m + N), X(x, x + N), Y(y, y + N), Z(z, z + N);
        sycl::queue q;
        sycl::buffer<double> bm(M.data(), N), bx(X.data(), N), by(Y.data(), N), bz(Z.data(), N);
        sycl::buffer<double> bax(ax, N), bay(ay, N), baz(az, N);
        TimeMeasurement timer;
        alg.octree.buildOctree(q, bx, by, bz, bm, timer);
        alg.computeAccelerations(q, bm, bx, by, bz, bax, bay, baz);
        if (num_nodes) *num_nodes = alg.octree.nextFreeNodeID_vec[0];
    });
}

// ---- whole simulations -------------------------------------------------------------------------------------------------
// algorithm 0 = naive, 1 = BarnesHut.  ref_configure first.  Runs startSimulation (main.cpp:230-237).
void *ref_sim_run(int algorithm, uint32_t N, const double *m, const double *x, const double *y, const double *z,
                  const double *vx, const double *vy, const double *vz, double dt, double t_end, double vs,
                  const char *output_directory) {
    RefSim *s = nullptr;
    int rc = guarded([&] {
        s = new RefSim();
        s->outdir = output_directory ? output_directory : ".";
        SimulationData &d = s->data;
        d.mass.assign(m, m + N);
        d.positions_x.assign(x, x + N); d.positions_y.assign(y, y + N); d.positions_z.assign(z, z + N);
        d.velocities_x.assign(vx, vx + N); d.velocities_y.assign(vy, vy + N); d.velocities_z.assign(vz, vz + N);
        d.names.assign(N, "body"); d.body_classes.assign(N, "AST");
        if (algorithm == 1) s->alg.reset(new BarnesHutAlgorithm(dt, t_end, vs, s->outdir));
        else s->alg.reset(new NaiveAlgorithm(dt, t_end, vs, s->outdir));
        CoutSilencer quiet;
        s->alg->startSimulation(d);
    });
    if (rc) { delete s; return nullptr; }
    return s;
}

// the same, bodies (names and classes included) read by the reference's InputParser from a CSV file
void *ref_sim_run_csv(int algorithm, const char *csv_path, int storage_param, int stack_param, double dt, double t_end,
                      double vs, const char *output_directory) {
    RefSim *s = nullptr;
    int rc = guarded([&] {
        s = new RefSim();
        s->outdir = output_directory ? output_directory : ".";
        std::string path(csv_path);
        InputParser::parse_input(path, s->data);
        configuration::initializeConfigValues((int_t) s->data.mass.size(), storage_param > 0 ? storage_param : 16,
                                              stack_param > 0 ? stack_param : 16);
        if (algorithm == 1) s->alg.reset(new BarnesHutAlgorithm(dt, t_end, vs, s->outdir));
        else s->alg.reset(new NaiveAlgorithm(dt, t_end, vs, s->outdir));
        CoutSilencer quiet;
        s->alg->startSimulation(s->data);
    });
    if (rc) { delete s; return nullptr; }
    return s;
}

void ref_sim_destroy(void *h) { delete (RefSim *) h; }
uint32_t ref_sim_num_bodies(void *h) { return (uint32_t) ((RefSim *) h)->data.mass.size(); }
uint32_t ref_sim_num_snapshots(void *h) { return (uint32_t) ((RefSim *) h)->alg->positions_x.size(); }

// which: 0..2 positions x,y,z; 3..5 velocities; 6 |a|.  Returns 0 when the step exists.
int ref_sim_get(void *h, uint32_t step, int which, double *out) {
    nBodyAlgorithm &a = *((RefSim *) h)->alg;
    std::map<int_t, std::vector<double>> *maps[7] = {&a.positions_x, &a.positions_y, &a.positions_z, &a.velocities_x,
                                                     &a.velocities_y, &a.velocities_z, &a.acceleration};
    auto it = maps[which]->find(step);
    if (it == maps[which]->end()) return 1;
    std::memcpy(out, it->second.data(), it->second.size() * sizeof(double));
    return 0;
}

int ref_sim_energy(void *h, uint32_t step, double out[4]) {
    nBodyAlgorithm &a = *((RefSim *) h)->alg;
    if (!a.kineticEnergy.count(step)) return 1;
    out[0] = a.kineticEnergy[step]; out[1] = a.potentialEnergy[step]; out[2] = a.totalEnergy[step];
    out[3] = a.virialEquilibrium[step];
    return 0;
}

// nBodyAlgorithm::generateParaViewOutput (nBodyAlgorithm.cpp:129-251) into <output_directory>/<ctime>/
int ref_sim_write_output(void *h) {
    RefSim &s = *(RefSim *) h;
    return guarded([&] {
        CoutSilencer quiet;
        s.alg->generateParaViewOutput(s.data);
    });
}

}  // extern "C"

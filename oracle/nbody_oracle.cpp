// ============================================================================
// oracle/nbody_oracle.cpp -- CPU restatement of the reference's gravity hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// path (libnbody_b200.so) never links, loads or calls anything in oracle/.
//
// Parity status: the tree-shape functions below are pinned against the reference's own
// golden vectors (tests/BarnesHutTest.cpp:11-220, restated in tests/test_oracle_golden.py).
// Accelerations / leapfrog / energy are "parity unpinned" at the reference: the reference
// has no test for them and its SYCL toolchain (AdaptiveCpp / DPC++) is absent here, so the
// reference binary cannot be built (every hot-path TU includes <sycl/sycl.hpp>,
// src/utility/Configuration.hpp:6).  For those functions this file follows the reference
// source expression by expression (citations per function) and defines the SYCL runtime
// builtins as  rsqrt(x) := 1.0/std::sqrt(x)  and  sqrt := std::sqrt  (two correctly rounded
// IEEE-754 operations; what AdaptiveCpp's OpenMP back end evaluates on the host).
//
// All citations are relative to /root/reference/.
// Build: see oracle/Makefile (g++ -O3 -fopenmp -ffp-contract=off: no FMA contraction, so
// every expression is evaluated exactly as written, independent of the host ISA).
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#include <omp.h>

typedef unsigned int int_t;  // d_type::int_t, src/utility/Configuration.hpp:8-10

namespace {

inline double ref_rsqrt(double x) { return 1.0 / std::sqrt(x); }

// ---------------------------------------------------------------------------
// Linearised SoA octree, same fields as BarnesHutOctree.hpp:22-95 (octant-major child table:
// child o of node n lives at octants[o*S + n], BarnesHutOctree.hpp:33-46).
// ---------------------------------------------------------------------------
struct Tree {
    int_t N = 0;
    int_t S = 0;  // storageSizeParameter = storage_size_param * N, Configuration.cpp:26
    std::vector<int_t> octants;
    std::vector<double> edge, minx, miny, minz, mass, comx, comy, comz;
    std::vector<int_t> bodyOfNode, bodyCount;
    std::vector<int> isLeaf;
    std::vector<int_t> sorted;
    int_t nextFree = 0;
    double aabb_min[3] = {0, 0, 0}, aabb_max[3] = {0, 0, 0}, aabb_edge = 0;
    int max_depth = 0;
};

// BarnesHutOctree.cpp:45-191 (computeMinMaxValuesAABB).  The per-work-item scratch arrays are
// value-initialised to 0.0 (:58-72), so the box always contains the origin.
void aabb(int_t N, const double *x, const double *y, const double *z, int threadCount, double mn[3], double mx[3],
          double *edge) {
    double min_x = std::numeric_limits<double>::infinity();
    double min_y = std::numeric_limits<double>::infinity();
    double min_z = std::numeric_limits<double>::infinity();
    double max_x = std::numeric_limits<double>::lowest();
    double max_y = std::numeric_limits<double>::lowest();
    double max_z = std::numeric_limits<double>::lowest();

    long bodiesPerThread = (long) std::ceil((double) N / (double) threadCount);
    std::vector<double> lminx(threadCount, 0.0), lminy(threadCount, 0.0), lminz(threadCount, 0.0);
    std::vector<double> lmaxx(threadCount, 0.0), lmaxy(threadCount, 0.0), lmaxz(threadCount, 0.0);
    for (long t = 0; t < threadCount; ++t) {  // the work-items of :91-111
        for (long i = bodiesPerThread * t; i < bodiesPerThread * t + bodiesPerThread; ++i) {
            if (i < (long) N) {
                lminx[t] = std::fmin(lminx[t], x[i]);
                lminy[t] = std::fmin(lminy[t], y[i]);
                lminz[t] = std::fmin(lminz[t], z[i]);
                lmaxx[t] = std::fmax(lmaxx[t], x[i]);
                lmaxy[t] = std::fmax(lmaxy[t], y[i]);
                lmaxz[t] = std::fmax(lmaxz[t], z[i]);
            }
        }
    }
    for (long i = 0; i < threadCount; ++i) {  // host reduction :122-159
        if (i < (long) N) {
            if (lminx[i] < min_x) min_x = lminx[i];
            if (lminy[i] < min_y) min_y = lminy[i];
            if (lminz[i] < min_z) min_z = lminz[i];
            if (lmaxx[i] > max_x) max_x = lmaxx[i];
            if (lmaxy[i] > max_y) max_y = lmaxy[i];
            if (lmaxz[i] > max_z) max_z = lmaxz[i];
        }
    }
    // cube growth :162-190
    double x_length = std::abs(max_x - min_x);
    double y_length = std::abs(max_y - min_y);
    double z_length = std::abs(max_z - min_z);
    double maxEdgeLength = std::max(x_length, std::max(y_length, z_length));
    if (maxEdgeLength == x_length) {
        min_z = min_z - ((maxEdgeLength - z_length) / 2);
        min_y = min_y - ((maxEdgeLength - y_length) / 2);
        max_z = max_z + ((maxEdgeLength - z_length) / 2);
        max_y = max_y + ((maxEdgeLength - y_length) / 2);
    } else if (maxEdgeLength == y_length) {
        min_x = min_x - ((maxEdgeLength - x_length) / 2);
        min_z = min_z - ((maxEdgeLength - z_length) / 2);
        max_x = max_x + ((maxEdgeLength - x_length) / 2);
        max_z = max_z + ((maxEdgeLength - z_length) / 2);
    } else {
        min_x = min_x - ((maxEdgeLength - x_length) / 2);
        min_y = min_y - ((maxEdgeLength - y_length) / 2);
        max_x = max_x + ((maxEdgeLength - x_length) / 2);
        max_y = max_y + ((maxEdgeLength - y_length) / 2);
    }
    mn[0] = min_x; mn[1] = min_y; mn[2] = min_z;
    mx[0] = max_x; mx[1] = max_y; mx[2] = max_z;
    *edge = maxEdgeLength;
}

// Octant of a body inside a node: ParallelOctreeTopDownSubtrees.cpp:400-406
// (= ParallelOctreeTopDownSynchronized.cpp:333-339, BarnesHutOctree.cpp:586-591).
inline int_t octant_of(const Tree &t, int_t node, double px, double py, double pz) {
    double e = t.edge[node];
    bool upperPart = py > t.miny[node] + (e / 2);
    bool rightPart = px > t.minx[node] + (e / 2);
    bool backPart = pz < t.minz[node] + (e / 2);
    return ((int) upperPart) * 4 + ((int) rightPart) * 2 + ((int) backPart) * 1;
}

// Node split: ParallelOctreeTopDownSubtrees.cpp:250-339 (= ...Synchronized.cpp:178-296).
// Children get 8 consecutive IDs for octants [5,7,4,6,1,3,0,2].
bool split(Tree &t, int_t node) {
    if ((uint64_t) t.nextFree + 8 > (uint64_t) t.S) return false;  // the reference would overflow silently
    int_t first = t.nextFree;
    t.nextFree += 8;
    double h = t.edge[node] / 2;
    double px = t.minx[node], py = t.miny[node], pz = t.minz[node];
    for (int_t k = first; k < first + 8; ++k) t.edge[k] = h;
    static const int_t order[8] = {5, 7, 4, 6, 1, 3, 0, 2};
    for (int k = 0; k < 8; ++k) t.octants[(size_t) order[k] * t.S + node] = first + k;
    t.minx[first] = px;         t.miny[first] = py + h;     t.minz[first] = pz;          // upperNW (5)
    t.minx[first + 1] = px + h; t.miny[first + 1] = py + h; t.minz[first + 1] = pz;      // upperNE (7)
    t.minx[first + 2] = px;     t.miny[first + 2] = py + h; t.minz[first + 2] = pz + h;  // upperSW (4)
    t.minx[first + 3] = px + h; t.miny[first + 3] = py + h; t.minz[first + 3] = pz + h;  // upperSE (6)
    t.minx[first + 4] = px;     t.miny[first + 4] = py;     t.minz[first + 4] = pz;      // lowerNW (1)
    t.minx[first + 5] = px + h; t.miny[first + 5] = py;     t.minz[first + 5] = pz;      // lowerNE (3)
    t.minx[first + 6] = px;     t.miny[first + 6] = py;     t.minz[first + 6] = pz + h;  // lowerSW (0)
    t.minx[first + 7] = px + h; t.miny[first + 7] = py;     t.minz[first + 7] = pz + h;  // lowerSE (2)
    for (int_t k = first; k < first + 8; ++k) {
        for (int o = 0; o < 8; ++o) t.octants[(size_t) o * t.S + k] = 0;
        t.bodyOfNode[k] = t.N;
        t.bodyCount[k] = 0;
        t.isLeaf[k] = 1;
        t.comx[k] = t.comy[k] = t.comz[k] = 0;
        t.mass[k] = 0;
    }
    return true;
}

// Sequential execution of the insertion loop of ParallelOctreeTopDownSynchronized.cpp:139-352 with one
// work-item (bodies 0..N-1 in order).  The resulting tree is the canonical one: the
// Subtrees builder (ParallelOctreeTopDownSubtrees.cpp:202-431,588-811) produces the same (depth, path)
// node set, only the node IDs depend on the interleaving.
int build(Tree &t, const double *x, const double *y, const double *z, int max_depth_guard) {
    // root init: ParallelOctreeTopDownSubtrees.cpp:135-163
    t.edge[0] = t.aabb_edge;
    t.minx[0] = t.aabb_min[0]; t.miny[0] = t.aabb_min[1]; t.minz[0] = t.aabb_min[2];
    for (int o = 0; o < 8; ++o) t.octants[(size_t) o * t.S] = 0;
    t.nextFree = 1;
    t.isLeaf[0] = 1;
    t.bodyCount[0] = 0;
    t.mass[0] = 0; t.comx[0] = t.comy[0] = t.comz[0] = 0;
    t.bodyOfNode[0] = t.N;
    t.max_depth = 0;
    for (int_t i = 0; i < t.N; ++i) {
        int_t cur = 0;
        int depth = 0;
        bool inserted = false;
        while (!inserted) {
            if (t.isLeaf[cur] == 1) {
                if (t.bodyOfNode[cur] == t.N) {
                    t.bodyOfNode[cur] = i;
                    inserted = true;
                } else {
                    if (depth >= max_depth_guard) return 2;  // coincident bodies: reference UB
                    int_t old = t.bodyOfNode[cur];
                    if (!split(t, cur)) return 1;
                    int_t o = octant_of(t, cur, x[old], y[old], z[old]);
                    int_t child = t.octants[(size_t) o * t.S + cur];
                    t.bodyOfNode[child] = old;
                    t.bodyOfNode[cur] = t.N;
                    t.isLeaf[cur] = 0;
                }
            } else {
                int_t o = octant_of(t, cur, x[i], y[i], z[i]);
                cur = t.octants[(size_t) o * t.S + cur];
                depth += 1;
                if (depth > t.max_depth) t.max_depth = depth;
            }
        }
    }
    return 0;
}

// prepareCenterOfMass (BarnesHutOctree.cpp:216-226) + the bottom-up sum of computeCenterOfMass_CPU/_GPU
// (BarnesHutOctree.cpp:299-317 / :424-545): children are summed in octant order 0..7, starting from 0.
void center_of_mass(Tree &t, const double *x, const double *y, const double *z, const double *m) {
    int_t n = t.nextFree;
    for (int_t i = 0; i < n; ++i) {
        int_t b = t.bodyOfNode[i];
        if (t.isLeaf[i] == 1 && b != t.N) {
            t.comx[i] = x[b] * m[b];
            t.comy[i] = y[b] * m[b];
            t.comz[i] = z[b] * m[b];
            t.mass[i] = m[b];
            t.bodyCount[i] = 1;
        }
    }
    // children always have larger IDs than their parent (fetch_add allocation), so a reverse sweep is a
    // valid bottom-up order; the arithmetic per node is the reference's.
    for (int_t k = n; k-- > 0;) {
        if (t.isLeaf[k] == 0) {
            double sumMasses = 0, cx = 0, cy = 0, cz = 0;
            int_t bodyCount = 0;
            for (int o = 0; o < 8; ++o) {
                int_t c = t.octants[(size_t) o * t.S + k];
                cx += t.comx[c];
                cy += t.comy[c];
                cz += t.comz[c];
                sumMasses += t.mass[c];
                bodyCount += t.bodyCount[c];
            }
            t.bodyCount[k] = bodyCount;
            t.comx[k] = cx; t.comy[k] = cy; t.comz[k] = cz;
            t.mass[k] = sumMasses;
        }
    }
}

// BarnesHutOctree.cpp:569-610 (sortBodies): rank = bodies in lower-octant siblings along the path.
void sort_bodies(Tree &t, const double *x, const double *y, const double *z) {
    for (int_t i = 0; i < t.N; ++i) {
        int_t insertionIndex = 0, cur = 0;
        while (!t.isLeaf[cur]) {
            int_t o = octant_of(t, cur, x[i], y[i], z[i]);
            for (int_t j = 0; j < o; ++j) insertionIndex += t.bodyCount[t.octants[(size_t) j * t.S + cur]];
            cur = t.octants[(size_t) o * t.S + cur];
        }
        t.sorted[insertionIndex] = i;
    }
}

}  // namespace

extern "C" {

// ---- constants ------------------------------------------------------------
// nBodyAlgorithm.hpp:55-61
double orc_gravitational_constant() {
    double G = 6.67428 * std::pow(10, -11);
    double meter_AU = 1.0 / (1.49597870691 * std::pow(10, 11));
    double second_Days = 1.0 / 86400;
    G = G * (std::pow(meter_AU, 3) / std::pow(second_Days, 2));
    return G;
}
// Configuration.cpp:6
double orc_epsilon2() { return std::pow(10, -22); }

// Configuration.cpp:24-33 (initializeConfigValues)
void orc_init_config(int_t bodyCount, int storageSizeParam, int stackSizeParam, int_t *storageSize, int_t *stackSize) {
    *storageSize = storageSizeParam * bodyCount;
    if (bodyCount < 15000) {
        *stackSize = stackSizeParam * (int_t) std::ceil(std::log2(bodyCount)) + 500;
    } else {
        *stackSize = stackSizeParam * (int_t) std::ceil(std::log2(bodyCount));
    }
}

int orc_max_threads() { return omp_get_max_threads(); }

// ---- naive all-pairs ---------------------------------------------------------
// NaiveAlgorithm.cpp:384-411 (opt_0 loop structure; opt_1 :446-476 and opt_2 :299-353 evaluate the same
// expression in the same j order, so all three stages have one restatement).  Rows [i0, i1) only, so the
// bench can time a bounded sample.  Bodies are processed in blocks of W rows so the compiler can vectorise
// ACROSS rows; each row still accumulates j = 0..N-1 in order, so results are bit-identical to the scalar loop.
void orc_naive_accel_rows(int_t N, const double *m, const double *x, const double *y, const double *z, double eps2,
                          double G, int_t i0, int_t i1, double *ax, double *ay, double *az, int nthreads) {
    const int W = 8;
    long nblk = ((long) i1 - (long) i0 + W - 1) / W;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long b = 0; b < nblk; ++b) {
        double px[W], py[W], pz[W], accx[W], accy[W], accz[W];
        int_t base = i0 + (int_t) (b * W);
        for (int k = 0; k < W; ++k) {
            int_t i = std::min(base + k, i1 - 1);
            px[k] = x[i]; py[k] = y[i]; pz[k] = z[i];
            accx[k] = accy[k] = accz[k] = 0;
        }
        for (int_t j = 0; j < N; ++j) {
            double xj = x[j], yj = y[j], zj = z[j], mj = m[j];
#pragma omp simd
            for (int k = 0; k < W; ++k) {
                double r_x = xj - px[k];
                double r_y = yj - py[k];
                double r_z = zj - pz[k];
                double denominator = (r_x * r_x) + (r_y * r_y) + (r_z * r_z) + eps2;
                denominator = denominator * denominator * denominator;
                denominator = 1.0 / std::sqrt(denominator);
                accx[k] += mj * (r_x * denominator);
                accy[k] += mj * (r_y * denominator);
                accz[k] += mj * (r_z * denominator);
            }
        }
        for (int k = 0; k < W; ++k) {
            int_t i = base + k;
            if (i < i1) {
                ax[i] = accx[k] * G;
                ay[i] = accy[k] * G;
                az[i] = accz[k] * G;
            }
        }
    }
}

void orc_naive_accel(int_t N, const double *m, const double *x, const double *y, const double *z, double eps2, double G,
                     double *ax, double *ay, double *az, int nthreads) {
    orc_naive_accel_rows(N, m, x, y, z, eps2, G, 0, N, ax, ay, az, nthreads);
}

// ---- leapfrog ------------------------------------------------------------------
// NaiveAlgorithm.cpp:154-163 (= BarnesHutAlgorithm.cpp:171-181)
void orc_leapfrog_part1(int_t N, double delta_t, double *x, double *y, double *z, const double *vx, const double *vy,
                        const double *vz, double *vhx, double *vhy, double *vhz, const double *ax, const double *ay,
                        const double *az) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) N; ++i) {
        vhx[i] = vx[i] + ax[i] * (delta_t / 2.0);
        vhy[i] = vy[i] + ay[i] * (delta_t / 2.0);
        vhz[i] = vz[i] + az[i] * (delta_t / 2.0);
        x[i] = x[i] + vhx[i] * delta_t;
        y[i] = y[i] + vhy[i] * delta_t;
        z[i] = z[i] + vhz[i] * delta_t;
    }
}
// NaiveAlgorithm.cpp:217-221 (= BarnesHutAlgorithm.cpp:234-238)
void orc_leapfrog_part2(int_t N, double delta_t, double *vx, double *vy, double *vz, const double *vhx,
                        const double *vhy, const double *vhz, const double *ax, const double *ay, const double *az) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long) N; ++i) {
        vx[i] = vhx[i] + ax[i] * (delta_t / 2.0);
        vy[i] = vhy[i] + ay[i] * (delta_t / 2.0);
        vz[i] = vhz[i] + az[i] * (delta_t / 2.0);
    }
}

// ---- energy ---------------------------------------------------------------------
// nBodyAlgorithm.cpp:48-85.  out = {kinetic, potential, total, virial}.  Optional per-body partials.
void orc_energy(int_t N, double G, const double *M, const double *P_X, const double *P_Y, const double *P_Z,
                const double *V_X, const double *V_Y, const double *V_Z, double out[4], double *E_kin_body,
                double *E_pot_body) {
    std::vector<double> E_KIN(N), E_POT(N);
#pragma omp parallel for schedule(dynamic, 64)
    for (long j = 0; j < (long) N; ++j) {
        double v = V_X[j] * V_X[j] + V_Y[j] * V_Y[j] + V_Z[j] * V_Z[j];
        E_KIN[j] = 0.5 * M[j] * v;
        double e = 0;
        for (long i = 0; i < j; ++i) {
            double r_x = P_X[j] - P_X[i];
            double r_y = P_Y[j] - P_Y[i];
            double r_z = P_Z[j] - P_Z[i];
            double r = std::sqrt(r_x * r_x + r_y * r_y + r_z * r_z);
            e += G * M[i] * M[j] / r;
        }
        E_POT[j] = e;
    }
    double E_kin_result = 0, E_pot_result = 0;
    for (int_t i = 0; i < N; ++i) {  // serial host sum in index order :71-74
        E_kin_result += E_KIN[i];
        E_pot_result += E_POT[i];
    }
    E_pot_result *= -1;
    out[0] = E_kin_result;
    out[1] = E_pot_result;
    out[2] = E_kin_result + E_pot_result;
    out[3] = (2.0 * E_kin_result) / std::abs(E_pot_result);
    if (E_kin_body) std::memcpy(E_kin_body, E_KIN.data(), sizeof(double) * N);
    if (E_pot_body) std::memcpy(E_pot_body, E_POT.data(), sizeof(double) * N);
}

// nBodyAlgorithm.cpp:88-102 (storeAccelerations)
void orc_accel_norm(int_t N, const double *ax, const double *ay, const double *az, double *out) {
    for (int_t i = 0; i < N; ++i) {
        double accelerationNorm = ax[i] * ax[i] + ay[i] * ay[i] + az[i] * az[i];
        out[i] = std::sqrt(accelerationNorm);
    }
}

// nBodyAlgorithm.cpp:104-127 (adjustVelocities): applied to the step-0 OUTPUT velocities only.
void orc_adjust_velocities(int_t N, const double *mass, const double *vx, const double *vy, const double *vz,
                           double *ovx, double *ovy, double *ovz) {
    double sumMasses = 0, sx = 0, sy = 0, sz = 0;
    for (int_t i = 0; i < N; ++i) {
        sumMasses += mass[i];
        sx += mass[i] * vx[i];
        sy += mass[i] * vy[i];
        sz += mass[i] * vz[i];
    }
    double ui_x = sx / sumMasses, ui_y = sy / sumMasses, ui_z = sz / sumMasses;
    for (int_t i = 0; i < N; ++i) {
        ovx[i] = vx[i] - ui_x;
        ovy[i] = vy[i] - ui_y;
        ovz[i] = vz[i] - ui_z;
    }
}

// ---- Barnes-Hut tree -------------------------------------------------------------
void orc_aabb(int_t N, const double *x, const double *y, const double *z, int workItems, double out[7]) {
    double mn[3], mx[3], e;
    aabb(N, x, y, z, workItems, mn, mx, &e);
    out[0] = mn[0]; out[1] = mn[1]; out[2] = mn[2];
    out[3] = mx[0]; out[4] = mx[1]; out[5] = mx[2];
    out[6] = e;
}

void *orc_tree_create(int_t N, int_t storageSize) {
    Tree *t = new Tree();
    t->N = N;
    t->S = storageSize;
    size_t S = storageSize;
    t->octants.assign(8 * S, 0);
    t->edge.assign(S, 0); t->minx.assign(S, 0); t->miny.assign(S, 0); t->minz.assign(S, 0);
    t->mass.assign(S, 0); t->comx.assign(S, 0); t->comy.assign(S, 0); t->comz.assign(S, 0);
    t->bodyOfNode.assign(S, N);
    t->bodyCount.assign(S, 0);
    t->isLeaf.assign(S, 1);
    t->sorted.assign(N, N);
    return t;
}
void orc_tree_destroy(void *h) { delete (Tree *) h; }

// BarnesHutOctree::buildOctree pipeline, ParallelOctreeTopDownSubtrees.cpp:15-93: AABB -> build -> COM -> sort.
// returns 0 ok, 1 node storage overflow, 2 depth guard (coincident bodies)
int orc_tree_build(void *h, const double *x, const double *y, const double *z, const double *m, int aabbWorkItems) {
    Tree &t = *(Tree *) h;
    aabb(t.N, x, y, z, aabbWorkItems, t.aabb_min, t.aabb_max, &t.aabb_edge);
    int rc = build(t, x, y, z, 200);
    if (rc) return rc;
    center_of_mass(t, x, y, z, m);
    sort_bodies(t, x, y, z);
    return 0;
}

int_t orc_tree_num_nodes(void *h) { return ((Tree *) h)->nextFree; }
int orc_tree_max_depth(void *h) { return ((Tree *) h)->max_depth; }
void orc_tree_aabb(void *h, double out[7]) {
    Tree &t = *(Tree *) h;
    for (int k = 0; k < 3; ++k) { out[k] = t.aabb_min[k]; out[3 + k] = t.aabb_max[k]; }
    out[6] = t.aabb_edge;
}
const int_t *orc_tree_body_of_node(void *h) { return ((Tree *) h)->bodyOfNode.data(); }
const int_t *orc_tree_body_count(void *h) { return ((Tree *) h)->bodyCount.data(); }
const int_t *orc_tree_octants(void *h) { return ((Tree *) h)->octants.data(); }
const int *orc_tree_is_leaf(void *h) { return ((Tree *) h)->isLeaf.data(); }
const double *orc_tree_sum_masses(void *h) { return ((Tree *) h)->mass.data(); }
const double *orc_tree_com_x(void *h) { return ((Tree *) h)->comx.data(); }
const double *orc_tree_com_y(void *h) { return ((Tree *) h)->comy.data(); }
const double *orc_tree_com_z(void *h) { return ((Tree *) h)->comz.data(); }
const double *orc_tree_edge(void *h) { return ((Tree *) h)->edge.data(); }
const double *orc_tree_min_x(void *h) { return ((Tree *) h)->minx.data(); }
const double *orc_tree_min_y(void *h) { return ((Tree *) h)->miny.data(); }
const double *orc_tree_min_z(void *h) { return ((Tree *) h)->minz.data(); }
const int_t *orc_tree_sorted_bodies(void *h) { return ((Tree *) h)->sorted.data(); }

// Canonical (ID-independent) node set: one record per node, keyed by (depth, path) where path is the
// sequence of octant codes from the root, packed 3 bits per level, left-aligned in 2 x 63 bits
// (levels 0..20 in path_hi bits 62..0, levels 21..41 in path_lo).  Records are emitted in DFS order with
// children in ascending octant code, i.e. sorted by (path_hi, path_lo, depth).
// kind: 0 = empty leaf, 1 = body leaf, 2 = internal.
struct CanonOut {
    int_t *depth; uint64_t *path_hi; uint64_t *path_lo; int_t *kind; int_t *body; int_t *count;
    double *edge, *minx, *miny, *minz, *mass, *comx, *comy, *comz;
};
static void canon_rec(const Tree &t, int_t node, int depth, uint64_t hi, uint64_t lo, CanonOut &o, size_t &k) {
    o.depth[k] = depth; o.path_hi[k] = hi; o.path_lo[k] = lo;
    o.kind[k] = t.isLeaf[node] ? (t.bodyOfNode[node] != t.N ? 1 : 0) : 2;
    o.body[k] = t.bodyOfNode[node];
    o.count[k] = t.bodyCount[node];
    o.edge[k] = t.edge[node]; o.minx[k] = t.minx[node]; o.miny[k] = t.miny[node]; o.minz[k] = t.minz[node];
    o.mass[k] = t.mass[node]; o.comx[k] = t.comx[node]; o.comy[k] = t.comy[node]; o.comz[k] = t.comz[node];
    ++k;
    if (!t.isLeaf[node]) {
        for (uint64_t oc = 0; oc < 8; ++oc) {
            uint64_t h2 = hi, l2 = lo;
            if (depth < 21) h2 |= oc << (60 - 3 * depth);
            else if (depth < 42) l2 |= oc << (60 - 3 * (depth - 21));
            canon_rec(t, t.octants[(size_t) oc * t.S + node], depth + 1, h2, l2, o, k);
        }
    }
}
void orc_tree_canonical(void *h, int_t *depth, uint64_t *path_hi, uint64_t *path_lo, int_t *kind, int_t *body,
                        int_t *count, double *edge, double *minx, double *miny, double *minz, double *mass,
                        double *comx, double *comy, double *comz) {
    Tree &t = *(Tree *) h;
    CanonOut o{depth, path_hi, path_lo, kind, body, count, edge, minx, miny, minz, mass, comx, comy, comz};
    size_t k = 0;
    canon_rec(t, 0, 0, 0, 0, o, k);
}

// BarnesHutAlgorithm.cpp:319-393 (computeAccelerations).  The reference's per-body slice of the global
// nodesOnStack buffer (:340-344) is a private vector here (same LIFO semantics, no 32-bit index wrap).
// stats (optional, 5 x uint64 per body): pops, non-empty visits (the :349 branch), accepted, opened, max stack depth.
void orc_bh_accel(void *h, const double *POS_X, const double *POS_Y, const double *POS_Z, double THETA, double epsilon_2,
                  double G, int bodiesSorted, double *ACC_X, double *ACC_Y, double *ACC_Z, uint64_t *stats,
                  int nthreads) {
    const Tree &t = *(Tree *) h;
    const int_t N = t.N;
    const size_t S = t.S;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<int_t> stack;
        stack.reserve(1024);
#pragma omp for schedule(dynamic, 256)
        for (long id = 0; id < (long) N; ++id) {
            double acc_x = 0, acc_y = 0, acc_z = 0;
            int_t i = bodiesSorted ? t.sorted[id] : (int_t) id;
            double pos_x_i = POS_X[i], pos_y_i = POS_Y[i], pos_z_i = POS_Z[i];
            uint64_t pops = 0, visits = 0, accepted = 0, opened = 0, maxsp = 1;
            stack.clear();
            stack.push_back(0);
            while (!stack.empty()) {
                int_t current_Node = stack.back();
                stack.pop_back();
                ++pops;
                if (t.mass[current_Node] != 0 && t.bodyOfNode[current_Node] != i) {
                    ++visits;
                    double d_x = (t.comx[current_Node] / t.mass[current_Node]) - pos_x_i;
                    double d_y = (t.comy[current_Node] / t.mass[current_Node]) - pos_y_i;
                    double d_z = (t.comz[current_Node] / t.mass[current_Node]) - pos_z_i;
                    double d = ref_rsqrt(d_x * d_x + d_y * d_y + d_z * d_z);
                    double currentTheta = t.edge[current_Node] * d;
                    if ((currentTheta < THETA) || t.bodyOfNode[current_Node] != N) {
                        ++accepted;
                        double denominator = (d_x * d_x) + (d_y * d_y) + (d_z * d_z) + epsilon_2;
                        denominator = denominator * denominator * denominator;
                        denominator = ref_rsqrt(denominator);
                        acc_x += t.mass[current_Node] * (d_x * denominator);
                        acc_y += t.mass[current_Node] * (d_y * denominator);
                        acc_z += t.mass[current_Node] * (d_z * denominator);
                    } else {
                        ++opened;
                        static const int order[8] = {5, 7, 4, 6, 1, 3, 0, 2};
                        for (int k = 0; k < 8; ++k) stack.push_back(t.octants[(size_t) order[k] * S + current_Node]);
                        if (stack.size() > maxsp) maxsp = stack.size();
                    }
                }
            }
            ACC_X[i] = acc_x * G;
            ACC_Y[i] = acc_y * G;
            ACC_Z[i] = acc_z * G;
            if (stats) {
                stats[5 * (size_t) i + 0] = pops; stats[5 * (size_t) i + 1] = visits;
                stats[5 * (size_t) i + 2] = accepted; stats[5 * (size_t) i + 3] = opened;
                stats[5 * (size_t) i + 4] = maxsp;
            }
        }
    }
}

// ---- subtree helpers of the default builder (golden vectors tests/BarnesHutTest.cpp:129-220) -----------
// prepareSubtrees, ParallelOctreeTopDownSubtrees.cpp:436-476
void orc_prepare_subtrees(int_t N, const int_t *subtreeOfBody, int_t nodeCount, int_t *bodyCountSubtree,
                          int_t *subtrees, int_t *subtreeCount) {
    for (int_t i = 0; i < nodeCount; ++i) bodyCountSubtree[i] = 0;
    for (int_t i = 0; i < N; ++i) bodyCountSubtree[subtreeOfBody[i]] += 1;
    int_t nextIndex = 0;
    for (int_t i = 1; i < nodeCount; ++i) {
        if (bodyCountSubtree[i] > 0) {
            subtrees[nextIndex] = i;
            nextIndex += 1;
        }
    }
    *subtreeCount = nextIndex;
}
// sortBodiesForSubtrees, ParallelOctreeTopDownSubtrees.cpp:478-534 (serial: slots are taken in body order)
void orc_sort_bodies_for_subtrees(int_t N, const int_t *subtreeOfBody, const int_t *bodyCountSubtree,
                                  const int_t *subtrees, int_t numberOfSubtrees, int_t *startIndex,
                                  int_t *sortedBodies) {
    std::vector<int_t> next(numberOfSubtrees, 0);
    for (int_t i = 0; i < numberOfSubtrees; ++i) {
        int_t firstIndex = 0;
        for (int j = (int) i - 1; j >= 0; j--) firstIndex += bodyCountSubtree[subtrees[j]];
        startIndex[i] = firstIndex;
        next[i] = firstIndex;
    }
    for (int_t i = 0; i < N; ++i) {
        int subtreeIndex = -1;
        for (int_t j = 0; j < numberOfSubtrees; ++j) {
            if (subtrees[j] == subtreeOfBody[i]) { subtreeIndex = (int) j; break; }
        }
        if (subtreeIndex != -1) sortedBodies[next[subtreeIndex]++] = i;
    }
}

// ---- whole simulation loop --------------------------------------------------------
// NaiveAlgorithm.cpp:15-260 / BarnesHutAlgorithm.cpp:18-278 (startSimulation).  algorithm: 0 naive, 1 BarnesHut.
// Snapshots ("visualised steps") are written to snap_* (capacity max_snap); step 0 velocities are the
// ADJUSTED ones (NaiveAlgorithm.cpp:31-36) while the integrator starts from the unadjusted ones (:43-45).
// returns 0 ok, 1/2 tree errors, 3 snapshot capacity exceeded.
int orc_simulate(int algorithm, int_t N, const double *mass, const double *x0, const double *y0, const double *z0,
                 const double *vx0, const double *vy0, const double *vz0, double dt, double t_end, double vs,
                 double theta, int compute_energy, int sort_bodies_flag, int storageSizeParam, int aabbWorkItems,
                 int nthreads, int_t max_snap, double *snap_px, double *snap_py, double *snap_pz, double *snap_vx,
                 double *snap_vy, double *snap_vz, double *snap_anorm, double *snap_energy, int_t *n_snap,
                 uint64_t *n_steps) {
    const double G = orc_gravitational_constant();
    const double eps2 = orc_epsilon2();
    std::vector<double> x(x0, x0 + N), y(y0, y0 + N), z(z0, z0 + N);
    std::vector<double> vx(vx0, vx0 + N), vy(vy0, vy0 + N), vz(vz0, vz0 + N);
    std::vector<double> ax(N), ay(N), az(N), vhx(N), vhy(N), vhz(N);
    Tree *tree = nullptr;
    if (algorithm == 1) tree = (Tree *) orc_tree_create(N, (int_t) storageSizeParam * N);
    int rc = 0;
    auto accel = [&]() -> int {
        if (algorithm == 0) {
            orc_naive_accel(N, mass, x.data(), y.data(), z.data(), eps2, G, ax.data(), ay.data(), az.data(), nthreads);
            return 0;
        }
        int r = orc_tree_build(tree, x.data(), y.data(), z.data(), mass, aabbWorkItems);
        if (r) return r;
        orc_bh_accel(tree, x.data(), y.data(), z.data(), theta, eps2, G, sort_bodies_flag, ax.data(), ay.data(),
                     az.data(), nullptr, nthreads);
        return 0;
    };
    auto snap3 = [&](double *dst, int_t step, const std::vector<double> &src) {
        std::memcpy(dst + (size_t) step * N, src.data(), sizeof(double) * N);
    };
    if (max_snap < 1) { rc = 3; goto done; }
    {
        // step 0 output: positions as read, velocities adjusted
        snap3(snap_px, 0, x); snap3(snap_py, 0, y); snap3(snap_pz, 0, z);
        orc_adjust_velocities(N, mass, vx0, vy0, vz0, snap_vx, snap_vy, snap_vz);
        double time = 0.0, timeSinceLastVisualization = 0.0;
        int_t currentStep = 0;
        uint64_t steps = 0;
        if ((rc = accel())) goto done;
        if (compute_energy)
            orc_energy(N, G, mass, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(), snap_energy, nullptr,
                       nullptr);
        orc_accel_norm(N, ax.data(), ay.data(), az.data(), snap_anorm);
        time += dt;
        timeSinceLastVisualization += dt;
        currentStep += 1;
        while (time <= t_end + 0.000001) {
            bool visualizeCurrentStep = (std::abs(timeSinceLastVisualization - vs) < 0.000001);
            if (visualizeCurrentStep && currentStep >= max_snap) { rc = 3; goto done; }
            orc_leapfrog_part1(N, dt, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(), vhx.data(),
                               vhy.data(), vhz.data(), ax.data(), ay.data(), az.data());
            if (visualizeCurrentStep) {
                snap3(snap_px, currentStep, x); snap3(snap_py, currentStep, y); snap3(snap_pz, currentStep, z);
            }
            if ((rc = accel())) goto done;
            orc_leapfrog_part2(N, dt, vx.data(), vy.data(), vz.data(), vhx.data(), vhy.data(), vhz.data(), ax.data(),
                               ay.data(), az.data());
            ++steps;
            if (visualizeCurrentStep) {
                orc_accel_norm(N, ax.data(), ay.data(), az.data(), snap_anorm + (size_t) currentStep * N);
                snap3(snap_vx, currentStep, vx); snap3(snap_vy, currentStep, vy); snap3(snap_vz, currentStep, vz);
                if (compute_energy)
                    orc_energy(N, G, mass, x.data(), y.data(), z.data(), vx.data(), vy.data(), vz.data(),
                               snap_energy + 4 * (size_t) currentStep, nullptr, nullptr);
                currentStep += 1;
                timeSinceLastVisualization = 0.0;
            }
            time += dt;
            timeSinceLastVisualization += dt;
        }
        *n_snap = currentStep;
        *n_steps = steps;
    }
done:
    if (tree) orc_tree_destroy(tree);
    return rc;
}

}  // extern "C"

// ---- analysis helper (not part of the reference): size of the UNION of the node sets visited by a group of bodies ----
// Used to study warp-cooperative traversal efficiency.  Returns the union size; *sum_visits = sum of per-body visits.
extern "C" uint64_t orc_bh_group_union(void *h, const double *POS_X, const double *POS_Y, const double *POS_Z, double THETA,
                                       const int_t *group, int_t gsize, uint64_t *sum_visits) {
    const Tree &t = *(Tree *) h;
    const int_t N = t.N;
    const size_t S = t.S;
    std::vector<int_t> all;
    std::vector<int_t> stack;
    uint64_t visits = 0;
    for (int_t g = 0; g < gsize; ++g) {
        int_t i = group[g];
        stack.clear();
        stack.push_back(0);
        while (!stack.empty()) {
            int_t n = stack.back();
            stack.pop_back();
            if (t.mass[n] != 0 && t.bodyOfNode[n] != i) {
                ++visits;
                all.push_back(n);
                double d_x = (t.comx[n] / t.mass[n]) - POS_X[i], d_y = (t.comy[n] / t.mass[n]) - POS_Y[i],
                       d_z = (t.comz[n] / t.mass[n]) - POS_Z[i];
                double d = ref_rsqrt(d_x * d_x + d_y * d_y + d_z * d_z);
                if (!((t.edge[n] * d < THETA) || t.bodyOfNode[n] != N)) {
                    static const int order[8] = {5, 7, 4, 6, 1, 3, 0, 2};
                    for (int k = 0; k < 8; ++k) stack.push_back(t.octants[(size_t) order[k] * S + n]);
                }
            }
        }
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    *sum_visits = visits;
    return all.size();
}

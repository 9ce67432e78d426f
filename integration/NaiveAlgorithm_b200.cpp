// Drop-in replacement for the reference's src/simulationBackend/NaiveAlgorithm.cpp.
//
// Compiled against the reference's UNMODIFIED NaiveAlgorithm.hpp: same class, same member functions, same
// signatures.  The SYCL kernels are gone; every member forwards to the C ABI of libnbody_b200.so.
//   NaiveAlgorithm::startSimulation              -> nb_set_bodies + time loop over nb_naive_accel / nb_leapfrog_part{1,2}
//   NaiveAlgorithm::computeAccelerations_opt_{0,1,2}(queue, masses, pos x/y/z, acc x/y/z)
//                                                -> nb_op_naive_accelerations (host arrays in, host arrays out)
#include "NaiveAlgorithm.hpp"      // the reference's header, unmodified
#include "b200_backend.hpp"

NaiveAlgorithm::NaiveAlgorithm(double dt, double tEnd, double visualizationStepWidth, std::string &outputDirectory)
        : nBodyAlgorithm(dt, tEnd, visualizationStepWidth, outputDirectory) {
    this->description = "Naive Algorithm";
}

void NaiveAlgorithm::startSimulation(const SimulationData &simulationData) {
    nb_ctx *ctx = b200::open_context(*this);
    timer.addTimingSequence("Acceleration Kernel Time");
    b200::run_time_loop(*this, ctx, simulationData, [&]() {
        b200::check(ctx, nb_naive_accel(ctx), "nb_naive_accel");
        double ms[NB_T_COUNT];
        b200::check(ctx, nb_get_timers(ctx, ms), "nb_get_timers");
        timer.addTimeToSequence("Acceleration Kernel Time", ms[NB_T_ACCEL]);
    });
    nb_destroy(ctx);
}

namespace {
// the operator form: the buffers' host memory goes in, the acceleration buffers' host memory comes out
void accelerations(nBodyAlgorithm &alg, buffer<double> &masses, buffer<double> &px, buffer<double> &py,
                   buffer<double> &pz, buffer<double> &ax, buffer<double> &ay, buffer<double> &az) {
    nb_ctx *ctx = b200::open_context(alg);
    host_accessor<double> M(masses), X(px), Y(py), Z(pz), AX(ax), AY(ay), AZ(az);
    const int rc = nb_op_naive_accelerations(ctx, masses.size(), &M[0], &X[0], &Y[0], &Z[0], &AX[0], &AY[0], &AZ[0]);
    const std::string err = rc == NB_OK ? "" : nb_last_error(ctx);
    nb_destroy(ctx);
    if (rc != NB_OK) throw std::runtime_error("nb_op_naive_accelerations: " + err);
}
}  // namespace

void NaiveAlgorithm::computeAccelerations_opt_2(queue &, buffer<double> &masses, buffer<double> &currentPositions_x,
                                                buffer<double> &currentPositions_y, buffer<double> &currentPositions_z,
                                                buffer<double> &acceleration_x, buffer<double> &acceleration_y,
                                                buffer<double> &acceleration_z) {
    accelerations(*this, masses, currentPositions_x, currentPositions_y, currentPositions_z, acceleration_x,
                  acceleration_y, acceleration_z);
}

void NaiveAlgorithm::computeAccelerations_opt_1(queue &, buffer<double> &masses, buffer<double> &currentPositions_x,
                                                buffer<double> &currentPositions_y, buffer<double> &currentPositions_z,
                                                buffer<double> &acceleration_x, buffer<double> &acceleration_y,
                                                buffer<double> &acceleration_z) {
    accelerations(*this, masses, currentPositions_x, currentPositions_y, currentPositions_z, acceleration_x,
                  acceleration_y, acceleration_z);
}

void NaiveAlgorithm::computeAccelerations_opt_0(queue &, buffer<double> &masses, buffer<double> &currentPositions_x,
                                                buffer<double> &currentPositions_y, buffer<double> &currentPositions_z,
                                                buffer<double> &acceleration_x, buffer<double> &acceleration_y,
                                                buffer<double> &acceleration_z) {
    accelerations(*this, masses, currentPositions_x, currentPositions_y, currentPositions_z, acceleration_x,
                  acceleration_y, acceleration_z);
}

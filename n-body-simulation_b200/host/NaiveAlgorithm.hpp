// Naive all-pairs back end (reference src/simulationBackend/NaiveAlgorithm.hpp:13-54).  The three
// computeAccelerations_opt_N variants of the reference differ only in how a SYCL device is fed; on B200 one kernel
// (csrc/naive.cu) serves every --opt_stage, so they all forward to computeAccelerations().
#pragma once
#include "nBodyAlgorithm.hpp"

class NaiveAlgorithm : public nBodyAlgorithm {
public:
    NaiveAlgorithm(double dt, double tEnd, double visualizationStepWidth, std::string &outputDirectory);

    void startSimulation(const SimulationData &simulationData) override;

    // accelerations of the bodies currently on the device (nb_naive_accel)
    void computeAccelerations();
    void computeAccelerations_opt_0() { computeAccelerations(); }
    void computeAccelerations_opt_1() { computeAccelerations(); }
    void computeAccelerations_opt_2() { computeAccelerations(); }
};

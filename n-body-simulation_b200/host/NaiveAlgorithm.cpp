#include "NaiveAlgorithm.hpp"

NaiveAlgorithm::NaiveAlgorithm(double dt, double tEnd, double visualizationStepWidth, std::string &outputDirectory)
    : nBodyAlgorithm(dt, tEnd, visualizationStepWidth, outputDirectory) {
    description = "Naive Algorithm";
}

void NaiveAlgorithm::computeAccelerations() { check(nb_naive_accel(ctx), "nb_naive_accel"); }

void NaiveAlgorithm::startSimulation(const SimulationData &simulationData) {
    openDevice(simulationData);
    timer.addTimingSequence("Acceleration Kernel Time");
    batchAlgorithm = 0;
    recordForceTimers = [this](const double *ms) { timer.addTimeToSequence("Acceleration Kernel Time", ms[NB_T_ACCEL]); };
    runTimeLoop(simulationData, [this]() {
        computeAccelerations();
        double ms[NB_T_COUNT];
        check(nb_get_timers(ctx, ms), "nb_get_timers");
        recordForceTimers(ms);
    });
}

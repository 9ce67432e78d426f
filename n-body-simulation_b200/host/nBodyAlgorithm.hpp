// Base class of the two simulation back ends: same public surface as the reference's nBodyAlgorithm
// (reference src/simulationBackend/nBodyAlgorithm.hpp:19-187) -- description, dt / t_end / visualisation step, G,
// the per-step snapshot maps, energy maps, generateParaViewOutput / outputLastState / adjustVelocities -- but the
// sycl::queue + sycl::buffer arguments are gone: device state lives in an nb_ctx (include/nbody_b200.h) owned by the
// object, and computeEnergy / storeAccelerations read it through the C ABI.
#pragma once
#include <cmath>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "Configuration.hpp"
#include "SimulationData.hpp"
#include "TimeMeasurement.hpp"
#include "nbody_b200.h"

class nBodyAlgorithm {
public:
    std::string description;  // "Naive Algorithm" / "Barnes-Hut Algorithm"
    std::string outputDirectory;
    double dt;
    double t_end;
    double visualizationStepWidth;
    double G;

    TimeMeasurement timer;

    // visualised step -> per-body values
    std::map<d_type::int_t, std::vector<double>> positions_x, positions_y, positions_z;
    std::map<d_type::int_t, std::vector<double>> velocities_x, velocities_y, velocities_z;
    std::map<d_type::int_t, std::vector<double>> acceleration;  // |a|
    std::map<d_type::int_t, double> kineticEnergy, potentialEnergy, totalEnergy, virialEquilibrium;

    nBodyAlgorithm(double dt, double t_end, double visualizationStepWidth, std::string &outputDirectory);
    virtual ~nBodyAlgorithm();

    virtual void startSimulation(const SimulationData &simulationData) = 0;

    void generateParaViewOutput(const SimulationData &simulationData);
    void outputLastState(const std::string &path);

    // Streaming mode (new, `--stream_output=true`): every visualised step is written as soon as it is
    // complete and its vectors are released, so host memory stays O(N) instead of O(N * snapshots).  The files are
    // byte-identical to the ones generateParaViewOutput writes.
    void enableStreaming(const SimulationData &simulationData);

    // Binary snapshots (new, `--vtp_format=binary`): the same arrays as the ASCII .vtp, written as VTK
    // XML "appended raw" blocks at full fp64 precision; the ASCII writer stays the byte-compatible default.
    void setBinaryOutput(bool on) { binaryOutput = on; }

    // Checkpoints (new, `--checkpoint=<path>` [`--checkpoint_every_vs=true`]): masses, positions and the
    // integrator's velocities as a binary state file (host/StateFile.hpp) after the last step and, optionally, after
    // every visualised step.  `--file=<checkpoint>` resumes; the continued trajectory is bit-identical.
    void setCheckpoint(const std::string &path, bool everyVisualizedStep) {
        checkpointPath = path;
        checkpointEveryVisualizedStep = everyVisualizedStep;
    }

    // energies of the current device state, stored under currentStep (reference nBodyAlgorithm.cpp:11-86)
    void computeEnergy(d_type::int_t currentStep);
    // |a| of the current device accelerations (reference nBodyAlgorithm.cpp:88-102)
    void storeAccelerations(d_type::int_t currentStep);
    // removes the mass-weighted mean velocity from the step-0 OUTPUT velocities only (reference :104-127)
    void adjustVelocities(const SimulationData &simulationData);

    // where generateParaViewOutput wrote its files (empty before the call)
    std::string lastOutputPath;

private:
    const SimulationData *streamData = nullptr;      // non-null in streaming mode
    d_type::int_t stepsStreamed = 0;                  // .vtp files already written
    std::vector<double> lastPos_x, lastPos_y, lastPos_z;
    bool binaryOutput = false;
    std::string checkpointPath;
    bool checkpointEveryVisualizedStep = false;
    void writeStepFileBinary(d_type::int_t step, const SimulationData &simulationData);
    // state = positions/velocities given (a visualised step's host copies) or, when null, read back from the device
    void writeCheckpoint(const SimulationData &simulationData, double time, const std::vector<double> *pos[3],
                         const std::vector<double> *vel[3]);
    void prepareOutputDirectory();
    void writeStepFile(d_type::int_t step, const SimulationData &simulationData);
    void streamStep(d_type::int_t step);

public:
    bool isOutputRank() const { return configuration::rank == 0; }

protected:
    nb_ctx *ctx = nullptr;

    // creates the device context from the configuration globals, joins the NCCL communicator when launched with one
    // process per GPU, uploads the bodies (unadjusted velocities) and registers the device name with the timer
    void openDevice(const SimulationData &simulationData);
    void check(int status, const char *what);
    // the time loop shared by both back ends (reference NaiveAlgorithm.cpp:82-259 = BarnesHutAlgorithm.cpp:102-277);
    // `forces` evaluates the accelerations of the current positions and records its own timing sequences
    void runTimeLoop(const SimulationData &simulationData, const std::function<void()> &forces);
    // set by the back end before runTimeLoop: the nb_advance algorithm id of `forces` (0 naive, 1 Barnes-Hut; -1 = never
    // batch) and the part of `forces` that appends one entry to each of its timing sequences from a timer array
    int batchAlgorithm = -1;
    std::function<void(const double *)> recordForceTimers;
};

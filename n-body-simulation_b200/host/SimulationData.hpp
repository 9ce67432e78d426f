// Host SoA container of the bodies read from the input file (reference src/simulationData/SimulationData.hpp:13-27):
// index in every vector == body id.
#pragma once
#include <string>
#include <vector>

struct SimulationData {
    std::vector<std::string> names;
    std::vector<std::string> body_classes;
    std::vector<double> mass;
    std::vector<double> positions_x, positions_y, positions_z;
    std::vector<double> velocities_x, velocities_y, velocities_z;
    // simulated time of this state in earth days: 0 for a CSV, the checkpoint's time for a binary state file
    // (host/StateFile.hpp); names / body_classes may be empty for binary input
    double start_time = 0.0;
};

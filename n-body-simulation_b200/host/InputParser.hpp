// CSV reader (reference src/simulationData/InputParser.hpp:17-23, InputParser.cpp:7-51).
// Format: id,name,class,mass[kg],pos_x,pos_y,pos_z[AU],vel_x,vel_y,vel_z[AU/day]; first line is a header.
// New: a file starting with the StateFile magic is read as a binary SoA state instead (host/StateFile.hpp).
#pragma once
#include <string>
#include <vector>

#include "SimulationData.hpp"

class InputParser {
public:
    static void parse_input(std::string &path, SimulationData &simulationData);
    // split on ',' keeping empty fields (tests/InputParserTest.cpp:4-32)
    static std::vector<std::string> splitString(std::string string);
};

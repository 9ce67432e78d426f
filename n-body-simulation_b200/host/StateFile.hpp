// Binary body-state file (new).  The reference reads bodies with a serial getline + stod CSV parser
// (reference src/simulationData/InputParser.cpp:15-31) and its only state dump, lastState.csv, holds positions alone
// (reference src/simulationBackend/nBodyAlgorithm.cpp:396-411), so a run can neither start from 10^7+ bodies in
// reasonable time nor be resumed.  This format holds everything the integrator needs, SoA like the device arrays:
//
//   offset  0  char[8]  magic "NBSTATE1"
//           8  u32      version (1)
//          12  u32      flags (bit 0: a name/class table follows the arrays)
//          16  u64      number of bodies N
//          24  f64      simulated time of the state in earth days
//          32  u8[32]   reserved, zero
//          64  f64[N] x 7   mass, pos_x, pos_y, pos_z, vel_x, vel_y, vel_z   (little endian, body id order)
//   then, if flags bit 0: for every body "name\0class\0"
//
// `--file=<state file>` is recognised by the magic (any extension); `--checkpoint=<path>` writes one at the end of a
// run.  Velocities are the integrator's own (not the mean-adjusted step-0 output copy), so a resumed run continues
// the leapfrog sequence bit for bit.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "SimulationData.hpp"

namespace StateFile {

constexpr char kMagic[9] = "NBSTATE1";
constexpr std::uint32_t kVersion = 1;
constexpr std::size_t kHeaderBytes = 64;

bool isStateFile(const std::string &path);
// throws std::invalid_argument on a missing, truncated or unsupported file
void read(const std::string &path, SimulationData &data, double *time = nullptr);
// names/classes are stored when data.names is non-empty; the seven arrays must have equal length
void write(const std::string &path, const SimulationData &data, double time);
// same from seven caller-owned arrays (mass, pos_x..z, vel_x..z) of n doubles, without copying them
void writeArrays(const std::string &path, std::size_t n, const double *const arrays[7],
                 const std::vector<std::string> &names, const std::vector<std::string> &classes, double time);

}  // namespace StateFile

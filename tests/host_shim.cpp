// Test-only C shim over the host-side C++ classes (InputParser, StateFile, TimeConverter, Configuration) so the Python tests can
// restate the reference's googletest cases (tests/InputParserTest.cpp, tests/TimeConverterTest.cpp).  g++ only, no CUDA.
#include <cstring>
#include <stdexcept>
#include <string>

#include "Configuration.hpp"
#include "InputParser.hpp"
#include "StateFile.hpp"
#include "TimeConverter.hpp"

extern "C" {

// fields joined with '\x1f'; returns the number of fields
int shim_split(const char *line, char *out, int out_len) {
    auto v = InputParser::splitString(line);
    std::string joined;
    for (size_t i = 0; i < v.size(); ++i) {
        if (i) joined += '\x1f';
        joined += v[i];
    }
    std::strncpy(out, joined.c_str(), out_len - 1);
    out[out_len - 1] = 0;
    return (int) v.size();
}

// 0 ok, 1 std::invalid_argument
int shim_convert_time(const char *s, double *out) {
    std::string t = s;
    try {
        *out = TimeConverter::convertToEarthDays(t);
    } catch (const std::invalid_argument &) {
        return 1;
    }
    return 0;
}

// parses a CSV and copies up to cap bodies; returns the body count (or -1 on error)
int shim_parse_csv(const char *path, int cap, double *mass, double *px, double *py, double *pz, double *vx, double *vy,
                   double *vz, char *names, int names_len) {
    SimulationData d;
    std::string p = path;
    try {
        InputParser::parse_input(p, d);
    } catch (const std::exception &) {
        return -1;
    }
    std::string joined;
    for (size_t i = 0; i < d.mass.size() && (int) i < cap; ++i) {
        mass[i] = d.mass[i];
        px[i] = d.positions_x[i]; py[i] = d.positions_y[i]; pz[i] = d.positions_z[i];
        vx[i] = d.velocities_x[i]; vy[i] = d.velocities_y[i]; vz[i] = d.velocities_z[i];
        if (i < d.names.size()) joined += d.names[i] + "|" + d.body_classes[i] + "\x1f";
    }
    std::strncpy(names, joined.c_str(), names_len - 1);
    names[names_len - 1] = 0;
    return (int) d.mass.size();
}

// CSV or state file -> state file (exercises StateFile::write and the name table); returns the body count or -1
int shim_convert_to_state(const char *in_path, const char *out_path, double time) {
    SimulationData d;
    std::string p = in_path;
    try {
        InputParser::parse_input(p, d);
        StateFile::write(out_path, d, time);
    } catch (const std::exception &) {
        return -1;
    }
    return (int) d.mass.size();
}

// start_time of a parsed input (0 for CSV); -1 on error
double shim_start_time(const char *path) {
    SimulationData d;
    std::string p = path;
    try {
        InputParser::parse_input(p, d);
    } catch (const std::exception &) {
        return -1.0;
    }
    return d.start_time;
}

void shim_init_config(unsigned n, int storage, int stack, unsigned *storage_size, unsigned *stack_size) {
    configuration::initializeConfigValues(n, storage, stack);
    *storage_size = configuration::barnes_hut_algorithm::storageSizeParameter;
    *stack_size = configuration::barnes_hut_algorithm::stackSize;
}

}  // extern "C"

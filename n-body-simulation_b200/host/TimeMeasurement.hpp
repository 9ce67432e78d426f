// Named timing sequences exported as times.json (reference src/utility/TimeMeasurement.hpp:9-29, .cpp:7-68).
// Same keys and layout as the reference so existing post-processing keeps working; values come from CUDA events.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "Configuration.hpp"

class TimeMeasurement {
    std::map<std::string, std::vector<double>> times;

public:
    std::string algorithmType;
    d_type::int_t bodyCount = 0;
    std::string device;

    void addTimingSequence(const std::string &name);
    void addTimeToSequence(const std::string &sequenceName, double time);
    void exportJSON(const std::string &path);
    void setProperties(std::string &algorithm, d_type::int_t &bodyCount, std::string &device);
    const std::map<std::string, std::vector<double>> &sequences() const { return times; }
};

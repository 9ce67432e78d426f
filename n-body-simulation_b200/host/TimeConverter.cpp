#include "TimeConverter.hpp"

#include <cstdlib>
#include <stdexcept>

double TimeConverter::convertToEarthDays(std::string &time) {
    static const char *kFormat = "Time values have to be of the format <double><h|d|m|y>";
    if (time.size() < 2) throw std::invalid_argument(kFormat);
    double factor;
    switch (time.back()) {
        case 'h': factor = 1.0 / 24; break;   // hours
        case 'd': factor = 1.0; break;        // days
        case 'm': factor = 30.4167; break;    // months
        case 'y': factor = 365.25; break;     // years
        default: throw std::invalid_argument(kFormat);
    }
    const std::string number = time.substr(0, time.size() - 1);
    std::size_t used = 0;
    double value;
    try {
        value = std::stod(number, &used);
    } catch (const std::exception &) {
        throw std::invalid_argument(kFormat);
    }
    if (used != number.size()) throw std::invalid_argument(kFormat);  // e.g. "1y 5m"
    return value * factor;
}
